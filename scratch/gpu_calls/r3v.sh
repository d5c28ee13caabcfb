#!/bin/bash
# round 2, call 3v (1 GPU): back-end gather with the loads of 2 x 32 events hoisted (64 registers) -- BE tests, C4 timing
timeout 600 python -m pytest tests/test_gpu_be.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scratch/prof_be.py 1.0 2>&1 | tail -2
