#!/bin/bash
# call 11 (8 GPUs): final bench line at N=8 (fused exchange, sharded upload, pipelined e2e)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 300 --warmup 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 300 gpurun_out/bench_n8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("n8 value %.3e ms %.4f warm %.4f e2e %.3e e2e_ms %.4f h2d %d" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]))
PY
