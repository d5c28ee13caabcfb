#!/bin/bash
# round 2, call 3h (1 GPU): TMA A/B on the final kernel (phase stamps + bench value), C2 packet solve time (plain / fused trials)
mkdir -p gpurun_out
for v in 1 0; do
  echo "== CMAXB_FE_TMA=$v"
  CMAXB_FE_TMA=$v timeout 300 python scratch/phase.py 2>&1 | grep "C2" | sed -n 1,5p
  CMAXB_FE_TMA=$v timeout 600 python bench.py --skip-configs --steps 200 > gpurun_out/r3h_tma$v.json 2> gpurun_out/r3h_tma$v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3h_tma$v.json").read().strip().splitlines() if l.startswith("{")][-1])
print("bench value %.3e (%.1f us/step) l2_warm %.1f us latency %.1f us value_only %.1f us" % (d["value"], d["ms_per_step"] * 1e3, d["l2_warm"]["ms_per_step"] * 1e3, d["latency"]["us_per_eval"], d["value_only"]["ms_per_step"] * 1e3))
PY
done
echo "== CMAXB_FE_NO_BINNING=1 (TMA image tiles kept, events in arrival order, LUT from global memory)"
CMAXB_FE_NO_BINNING=1 timeout 300 python scratch/phase.py 2>&1 | grep "C2" | sed -n 1,3p
echo "== solve"
timeout 300 python scratch/solve_time.py 2>&1 | tail -5
