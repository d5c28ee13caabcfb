#!/bin/bash
# round 2, call o (2 GPUs): bench.py under torchrun -- hypothesis-sharded C2 with the fused exchange + lanes, C3 (256 hypotheses / 2), C4 / C5 sharded by time
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
echo "bench rc=$?"; tail -c 1200 gpurun_out/r2o_bench_n2.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2o_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=%d value %.3e ms %.4f e2e %.3e (%.1f us/step) launches %d" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["gpu_launches"]))
    for k, v in d.get("configs", {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f" % v["roofline"]["frac"])
except Exception as e:
    print("parse failed", e)
PY
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2o_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
for k in ("C4", "C5"):
    print(k, d["configs"][k].get("exchange_modes"))
PY
