#!/bin/bash
# round 2, final call B (8 GPUs): bench.py under torchrun at N = 8 (C2 headline with the fused exchange, e2e with 8 event streams, C3 256 hypotheses, C4 / C5 sharded by time: three exchanges)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 > gpurun_out/r3z_bench_n8.json 2> gpurun_out/r3z_bench_n8.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r3z_bench_n8.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3z_bench_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N=%d value %.3e ms %.4f e2e %.3e (%.1f us/step) launches %d" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["gpu_launches"]))
for k, v in d.get("configs", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f" % v["roofline"]["frac"])
PY
