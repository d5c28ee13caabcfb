#!/bin/bash
# call 3 (2 GPUs): fused in-kernel exchange over NVLink vs NCCL all-gather, + N=1 line with the reverted kernel
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 | tee gpurun_out/topo.txt
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for c in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 --collective $c > gpurun_out/bench_n2_$c.json 2> gpurun_out/bench_n2_$c.err
  tail -c 600 gpurun_out/bench_n2_$c.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 10 --collective p2p --depth 1 > gpurun_out/bench_n2_p2p_d1.json 2> gpurun_out/bench_n2_p2p_d1.err
python - <<'PY'
import json
for f in ("n1", "n2_p2p", "n2_nccl", "n2_p2p_d1"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f warm %.4f lat_us %.1f e2e %.3e frac %.3f kern_us %.1f" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
    except Exception as e:
        print(f, "failed", e)
PY
