#!/bin/bash
# round 2, call 3e (2 GPUs): multi-GPU test of the back-end exchanges; new FE batch-size tests
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fe.py -m gpu -q 2>&1 | tail -5
