#!/bin/bash
mkdir -p gpurun_out
make -C oracle -s
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 10 > gpurun_out/bench_adjoint.json 2> gpurun_out/bench_adjoint.err
python bench.py --steps 300 --warmup 10 --grad-mode dense > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
tail -c 400 gpurun_out/bench_adjoint.err
python scratch/fe_probe.py 2>&1 | grep -E "eval want|kernels|mismatch" | tail -8 | tee gpurun_out/fe_probe.log
