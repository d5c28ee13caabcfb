#!/bin/bash
# round 2, call 3p (1 GPU): final check of the committed state -- whole GPU suite, smoke(), default bench (all keys), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r3p_bench_n1.json 2> gpurun_out/r3p_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3p_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print(sorted(d.keys()))
print("value %.3e ms %.4f e2e %.3e (%.1f us/step) frac %.3f launches %d clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
print({k: (round(v["fg_ms"], 3) if "fg_ms" in v else round(v["latency_fg_us"], 1)) for k, v in d["configs"].items()})
PY
timeout 600 python bench.py --impl reference > gpurun_out/r3p_ref.json 2> gpurun_out/r3p_ref.err; echo "ref rc=$?"; head -c 400 gpurun_out/r3p_ref.json
