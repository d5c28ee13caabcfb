#!/bin/bash
# call 8 (2 GPUs): sharded upload e2e at N=2 (vs replicated), all GPU tests incl. the C4 full-size properties
mkdir -p gpurun_out
make -C oracle -s
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for u in sharded replicated; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 300 --warmup 10 --upload $u > gpurun_out/bench_n2_$u.json 2> gpurun_out/bench_n2_$u.err
  tail -c 400 gpurun_out/bench_n2_$u.err
done
python - <<PY
import json
for f in ("n2_sharded", "n2_replicated"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f e2e %.3e e2e_ms %.4f h2d %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
