#!/bin/bash
# call 5 (1 GPU): full GPU tests, bench, ncu launch list + full capture of the fused kernel, all configs, e2e timing
mkdir -p gpurun_out
make -C oracle -s
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_eval_megakernel -s 6 -c 2 -o gpurun_out/prof_fe_fused -f python bench.py --steps 5 --warmup 3 >> gpurun_out/ncu_bench.log 2>&1
timeout 200 python scratch/e2e_time.py 2>&1 | tail -3 | tee gpurun_out/e2e_time.txt
timeout 600 python scratch/all_configs.py 2>&1 | tee gpurun_out/all_configs.jsonl | cut -c1-300
ls -la gpurun_out | head -30
