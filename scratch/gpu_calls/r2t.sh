#!/bin/bash
# round 2, call t (1 GPU): where the e2e step goes -- host time per call class, depth, uploads + preparation only
mkdir -p gpurun_out
for v in "CMAXB_E2E_TRACE=1" "CMAXB_E2E_TRACE=1 CMAXB_E2E_DEPTH=6" "CMAXB_E2E_TRACE=1 CMAXB_E2E_SKIP=eval" "CMAXB_E2E_TRACE=1 CMAXB_FE_LANES=1"; do
  echo "== $v"
  env $v timeout 600 python bench.py --skip-configs --steps 100 > gpurun_out/r2t.json 2> gpurun_out/r2t.err
  grep "e2e host" gpurun_out/r2t.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2t.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e e2e %.3e (%.1f us/step)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3))
PY
done
