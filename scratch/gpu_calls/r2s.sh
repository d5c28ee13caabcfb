#!/bin/bash
# round 2, call s (1 GPU): two-launch packet binning -- FE / stream / pipeline tests, then the default bench (e2e is the number to watch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_stream_device.py tests/test_gpu_firstparty.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --skip-c5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2s_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2s_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e ms %.4f e2e %.3e (%.1f us/step, h2d %d) frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["e2e"]["h2d_bytes_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
for k in ("C1", "C3", "C4"):
    v = d["configs"][k]; print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 20 --warmup 3 --skip-configs > gpurun_out/r2s_ncu.log 2>&1
grep -c fe_bin gpurun_out/r2s_launches.csv
