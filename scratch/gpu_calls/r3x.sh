#!/bin/bash
# round 2, call 3x (1 GPU): 32 x 32 tiles in the stand-alone image kernels -- whole GPU suite, C4 classes, small windows
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python scratch/prof_be.py 1.0 2>&1 | tail -2
PROBE_TAG=graph timeout 300 python scratch/be_small_probe.py 2>&1 | tail -4
