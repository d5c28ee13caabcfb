#!/bin/bash
# round 2, call 3r (1 GPU): e2e against lanes / grid fraction / depth (is the e2e step bound by SM slots shared between preparation and evaluations?)
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --skip-configs --steps 100 > gpurun_out/r3r.json 2> gpurun_out/r3r.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3r.json").read().strip().splitlines() if l.startswith("{")][-1])
print("%-60s value %.3e (%.1f us)  e2e %.3e (%.1f us/step)" % ("$1", d["value"], d["ms_per_step"] * 1e3, d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3))
PY
}
run "CMAXB_X=0"
run "CMAXB_FE_LANES=2"
run "CMAXB_FE_LANES=4 CMAXB_FE_GRID_FRACTION=0.34"
run "CMAXB_FE_LANES=3 CMAXB_FE_GRID_FRACTION=0.34"
run "CMAXB_E2E_DEPTH=4"
run "CMAXB_E2E_DEPTH=2"
