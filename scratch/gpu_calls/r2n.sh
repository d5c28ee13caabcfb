#!/bin/bash
# round 2, call n (1 GPU): peer-memory sparse exchange of the back end -- single-rank and two-process (IPC on one GPU) parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -30
