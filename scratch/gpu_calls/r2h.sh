#!/bin/bash
# round 2, call h: full GPU suite incl. full-size parity; BE 24-byte cache + graph replay; e2e with copy stream; bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2h_pytest.log
PROBE_TAG=graph timeout 600 python scratch/be_small_probe.py c4 2>&1 | tail -8 | tee gpurun_out/r2h_be.txt
PROBE_TAG=nograph CMAXB_BE_GRAPH=0 timeout 600 python scratch/be_small_probe.py c4 2>&1 | tail -8 | tee -a gpurun_out/r2h_be.txt
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2h_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
    print("value %.3e ms %.4f e2e %.3e (h2d %.0f, %.1f us/step) frac %.3f thr_frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["ms_per_step"] * 1e3, d["roofline"]["frac"], d["roofline"]["throughput_frac"], d["gpu_launches"]))
    for k, v in d.get("configs", {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f" % v["roofline"]["frac"])
except Exception as e:
    print("parse failed", e)
PY
