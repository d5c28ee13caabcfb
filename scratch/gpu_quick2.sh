#!/bin/bash
mkdir -p gpurun_out
make -C oracle -s
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python scratch/phase.py 2>&1 | grep -E "^C[12]" | tee gpurun_out/phase.log
python scratch/fe_probe.py 2>&1 | grep -E "eval want|kernels|mismatch" | tee gpurun_out/fe_probe.log
