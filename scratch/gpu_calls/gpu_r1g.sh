#!/bin/bash
# call 6 (N GPUs): scaling bench with the fused in-kernel exchange (and NCCL for comparison)
N=$1
mkdir -p gpurun_out
for c in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 300 --warmup 10 --collective $c > gpurun_out/bench_n${N}_$c.json 2> gpurun_out/bench_n${N}_$c.err
  tail -c 300 gpurun_out/bench_n${N}_$c.err
done
python - <<PY
import json
for f in ("n${N}_p2p", "n${N}_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f warm %.4f lat_us %.1f e2e %.3e frac %.3f kern_us %.1f" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
    except Exception as e:
        print(f, "failed", e)
PY
