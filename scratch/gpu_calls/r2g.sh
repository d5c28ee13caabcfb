#!/bin/bash
# round 2, call g: new bench.py (N = 1) + ncu capture / launch list of the committed fused kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2g_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
    print("value %.3e ms %.4f e2e %.3e (h2d %.0f) frac %.3f thr_frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["roofline"]["frac"], d["roofline"]["throughput_frac"], d["gpu_launches"]))
    for k in ("l2_warm", "latency", "value_only", "flushed_serial", "latency_pipelined_1"):
        print(k, d.get(k))
    for k, v in d.get("configs", {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f adj %.3f" % (v["roofline"]["frac"], v["roofline"]["adjoint_min_frac"]))
    print("cpu", d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_eval_fused -s 12 -c 2 -o gpurun_out/r2g_fused -f python scratch/prof_fe.py > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 5 --warmup 3 --skip-configs > gpurun_out/r2g_bench_ncu.log 2>&1
tail -3 gpurun_out/r2g_launches.csv
