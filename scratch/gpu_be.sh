#!/bin/bash
mkdir -p gpurun_out
make -C oracle -s
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python scratch/prof_be.py 1.0 | tee gpurun_out/be_c4.log
python scratch/prof_be.py 0.1 | tee -a gpurun_out/be_c4.log
