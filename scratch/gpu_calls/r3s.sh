#!/bin/bash
# round 2, call 3s (1 GPU): e2e against the number of evaluations in flight (packet slots 6)
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --skip-configs --steps 100 > gpurun_out/r3s.json 2> gpurun_out/r3s.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3s.json").read().strip().splitlines() if l.startswith("{")][-1])
print("%-60s value %.3e (%.1f us)  e2e %.3e (%.1f us/step)" % ("$1", d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3))
PY
}
run "CMAXB_E2E_DEPTH=3"
run "CMAXB_E2E_DEPTH=4"
run "CMAXB_E2E_DEPTH=5"
run "CMAXB_E2E_DEPTH=6"
run "CMAXB_E2E_DEPTH=5 CMAXB_FE_LANES=4"
