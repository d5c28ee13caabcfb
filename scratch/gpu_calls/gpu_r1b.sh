#!/bin/bash
# round-1 re-entry call 1 (1 GPU): full GPU tests, bench at depth 2/1, f32 vs f64 gather A/B, phase stamps
mkdir -p gpurun_out
make -C oracle -s
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.err
CMAXB_FE_GATHER_F64=1 timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1_g64.json 2> gpurun_out/bench_n1_g64.err
timeout 120 python scratch/phase.py 2>&1 | tee gpurun_out/phase_f32.txt
CMAXB_FE_GATHER_F64=1 timeout 120 python scratch/phase.py 2>&1 | tee gpurun_out/phase_f64.txt
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_g64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f warm %.4f lat_us %.1f e2e %.3e frac %.3f kern_us %.1f" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
    except Exception as e:
        print(f, "failed", e)
PY
