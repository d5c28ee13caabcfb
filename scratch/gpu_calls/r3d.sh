#!/bin/bash
# round 2, call 3d (1 GPU): gather pass reading the scatter's per-event results (vote cache) -- FE tests, phase stamps and bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_full_size.py tests/test_optim.py tests/test_gsl_adapter.py -m gpu -x -q 2>&1 | tail -4
for v in 1 0; do
  echo "== CMAXB_FE_VOTE_CACHE=$v"
  CMAXB_FE_VOTE_CACHE=$v timeout 300 python scratch/phase.py 2>&1 | grep "C2" | sed -n 1,3p
  CMAXB_FE_VOTE_CACHE=$v timeout 600 python bench.py --skip-configs > gpurun_out/r3d_$v.json 2> gpurun_out/r3d_$v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3d_$v.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e ms %.4f e2e %.3e (%.1f us/step) frac %.3f l2_warm %.4f latency %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["roofline"]["frac"], d["l2_warm"]["ms_per_step"], d["latency"]["us_per_eval"]))
PY
done
