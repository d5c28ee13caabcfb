#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_firstparty.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_firstparty.log
