#!/bin/bash
# round 2, call 3l (1 GPU): back-end gather with interleaved loads, pose kernel capped at 80 registers -- BE tests, C4 / C5 / small windows
timeout 600 python -m pytest tests/test_gpu_be.py tests/test_gpu_sharded.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/prof_be.py 1.0 2>&1 | tail -2
PROBE_TAG=graph timeout 300 python scratch/be_small_probe.py 2>&1 | tail -4
timeout 300 python scratch/be_band_probe.py C5 1 2>&1 | grep -E "PLAIN|plain" | head -3
