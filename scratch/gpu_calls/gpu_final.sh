#!/bin/bash
# final validation (1 GPU): what the driver runs at round end
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_final.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("final value %.3e ms %.4f e2e %.3e frac %.3f launches %d cpu %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["cpu_baseline"]["value"]))
PY
