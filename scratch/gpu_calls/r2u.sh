#!/bin/bash
# round 2, call u (1 GPU): CUDA-event timeline of the e2e loop (lanes = 1)
timeout 300 python scratch/e2e_timeline.py 2>&1 | tail -20
