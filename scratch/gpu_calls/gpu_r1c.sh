#!/bin/bash
# call 2 (1 GPU): tests after the mu-free adjoint restructure, bench + phase stamps, compile-time variants A/B
mkdir -p gpurun_out
make -C oracle -s
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.err
timeout 120 python scratch/phase.py 2>&1 | grep C2 | tee gpurun_out/phase_base.txt
for v in occ4 ev8 ev2 g4 occ4ev2; do
  export CMAXB_LIB_PATH=$PWD/scratch/variants/libcmax_b200_$v.so
  timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  timeout 120 python scratch/phase.py 2>&1 | grep "C2" | head -4 | tee gpurun_out/phase_$v.txt
  unset CMAXB_LIB_PATH
done
python - <<'PY'
import json
for f in ("n1", "occ4", "ev8", "ev2", "g4", "occ4ev2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f warm %.4f lat_us %.1f e2e %.3e frac %.3f kern_us %.1f" % (d["value"], d["ms_per_step"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
    except Exception as e:
        print(f, "failed", e)
PY
