#!/bin/bash
# one GPU round: tests, bench, launch list + one full ncu capture of the top kernel
set -x
mkdir -p gpurun_out
make -C oracle -s
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 10 --grad-mode dense > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
python bench.py --steps 300 --warmup 10 --grad-mode adjoint > gpurun_out/bench_adjoint.json 2> gpurun_out/bench_adjoint.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 600 gpurun_out/bench_dense.err gpurun_out/bench_adjoint.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fe_eval_megakernel -s 6 -c 2 -o gpurun_out/prof_fe_fused -f python bench.py --steps 5 --warmup 3 >> gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"be_scatter|be_gather" -s 2 -c 4 -o gpurun_out/prof_be -f python scratch/prof_be.py >> gpurun_out/ncu_bench.log 2>&1
python scratch/prof_be.py 1.0 | tee gpurun_out/be_c4.log
ls -la gpurun_out
