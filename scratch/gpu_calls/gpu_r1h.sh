#!/bin/bash
# call 7 (8 GPUs): FE scaling bench with the fused exchange; back-end windows C4 / C5 sharded by time over 8 ranks
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 300 --warmup 10 --collective p2p > gpurun_out/bench_n8_p2p.json 2> gpurun_out/bench_n8_p2p.err
tail -c 300 gpurun_out/bench_n8_p2p.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 scratch/be_sharded_bench.py C4 1 2>&1 | grep -E "^\{|rror" | tee gpurun_out/be_sharded_c4_n8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scratch/be_sharded_bench.py C5 5 2>&1 | grep -E "^\{|rror" | tee gpurun_out/be_sharded_c5_n8.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n8_p2p.json").read().strip().splitlines()[-1])
print("n8_p2p value %.3e ms %.4f warm %.4f lat_us %.1f e2e %.3e" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"], d["e2e"]["value"]))
PY
