#!/bin/bash
# call 9 (4 GPUs): sharded-upload e2e at N=4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 300 --warmup 10 --upload sharded > gpurun_out/bench_n4_sharded.json 2> gpurun_out/bench_n4_sharded.err
tail -c 400 gpurun_out/bench_n4_sharded.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n4_sharded.json").read().strip().splitlines()[-1])
print("n4_sharded value %.3e ms %.4f e2e %.3e e2e_ms %.4f h2d %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]))
PY
