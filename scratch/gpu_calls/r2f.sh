#!/bin/bash
# round 2, call f: full GPU suite on the cleaned-up fused kernel (no records, atomic accumulators) + stamps + lanes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
echo "== phase stamps: default" | tee -a gpurun_out/r2f_phase.txt
timeout 300 python scratch/phase.py 2>&1 | tail -16 | tee -a gpurun_out/r2f_phase.txt
PROBE_TAG=default timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -10 | tee -a gpurun_out/r2f_lanes.txt
