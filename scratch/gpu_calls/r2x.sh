#!/bin/bash
# round 2, call x (1 GPU): back-end launch diet (accumulators cleared by the pose kernel, results written to mapped memory) -- all GPU tests, small-window probe
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
PROBE_TAG=graph timeout 300 python scratch/be_small_probe.py C4 2>&1 | tail -6
