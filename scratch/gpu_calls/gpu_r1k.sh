#!/bin/bash
# call 10 (2 GPUs): pipelined e2e loop at N=1 and N=2
mkdir -p gpurun_out
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 300 gpurun_out/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 300 --warmup 10 --collective nccl --upload replicated > gpurun_out/bench_n2_nccl_repl.json 2> gpurun_out/bench_n2_nccl_repl.err
tail -c 300 gpurun_out/bench_n2_nccl_repl.err
python - <<PY
import json
for f in ("n1", "n2", "n2_nccl_repl"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f e2e %.3e e2e_ms %.4f h2d %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
