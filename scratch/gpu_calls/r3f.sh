#!/bin/bash
# round 2, final call A (1 GPU): default bench, reference arm, ncu launch list of the bench command, ncu --set full of the packet preparation kernels
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r3f_bench_n1.json 2> gpurun_out/r3f_bench_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r3f_bench_n1.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r3f_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e ms %.4f e2e %.3e (%.1f us/step, h2d %.0f) frac %.3f thr_frac %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3, d["e2e"]["h2d_bytes_per_step"], d["roofline"]["frac"], d["roofline"]["throughput_frac"], d["gpu_launches"]))
for k in ("l2_warm", "latency", "value_only"):
    print(k, d.get(k))
for k, v in d.get("configs", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("roofline", "workload")}, "frac %.3f" % v["roofline"]["frac"])
print("cpu", d["cpu_baseline"], d["clocks"])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3f_bench_reference.json 2> gpurun_out/r3f_bench_reference.err
echo "reference rc=$?"; tail -c 300 gpurun_out/r3f_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 20 --warmup 3 --skip-configs > gpurun_out/r3f_bench_ncu.log 2>&1
tail -2 gpurun_out/r3f_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_bin -s 4 -c 2 -o gpurun_out/r3f_prep -f python scratch/prof_fe.py > gpurun_out/r3f_ncu.log 2>&1
tail -2 gpurun_out/r3f_ncu.log
