#!/bin/bash
# round 2, call k (1 GPU): row-band sharded BE image phases -- emulated-rank test + the two-process gloo test; BE kernel times of C5's pieces
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -15
