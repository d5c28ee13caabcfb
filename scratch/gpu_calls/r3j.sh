#!/bin/bash
# round 2, call 3j (1 GPU): bench value with the non-TMA fallback kernel (A/B for profiles/r02_tma_ab.txt)
mkdir -p gpurun_out
for v in 0 1; do
CMAXB_FE_TMA=$v timeout 600 python bench.py --skip-configs --steps 200 > gpurun_out/r3j_tma$v.json 2> gpurun_out/r3j_tma$v.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3j_tma$v.json").read().strip().splitlines() if l.startswith("{")][-1])
print("CMAXB_FE_TMA=$v bench value %.3e (%.1f us/step) l2_warm %.1f us latency %.1f us value_only %.1f us e2e %.3e kernel (profiler) %.1f us" % (d["value"], d["ms_per_step"] * 1e3, d["l2_warm"]["ms_per_step"] * 1e3, d["latency"]["us_per_eval"], d["value_only"]["ms_per_step"] * 1e3, d["e2e"]["value"], d["roofline"].get("kernel_us", 0) or 0))
PY
done
