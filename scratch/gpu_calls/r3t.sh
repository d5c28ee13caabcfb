#!/bin/bash
# round 2, call 3t (1 GPU): default bench after the e2e pipeline depth change (5 evaluations in flight over 6 packet slots)
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r3t_bench_n1.json 2> gpurun_out/r3t_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3t_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e ms %.4f e2e %.3e (%.1f us/step) frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["roofline"]["frac"], d["gpu_launches"]))
print({k: (round(v["fg_ms"], 3) if "fg_ms" in v else round(v["latency_fg_us"], 1)) for k, v in d["configs"].items()})
PY
timeout 600 python -m pytest tests/test_bench_contract.py -q 2>&1 | tail -1
