#!/bin/bash
# round 2, call a: parity of the new fused kernel + phase stamps + A/B (gather records, TMA staging) + lane throughput
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv | tail -1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
for v in "default" "CMAXB_FE_CACHE=0" "CMAXB_FE_TMA=0" "CMAXB_FE_CACHE=0 CMAXB_FE_TMA=0"; do
  echo "== phase stamps: $v" | tee -a gpurun_out/r2a_phase.txt
  if [ "$v" = "default" ]; then timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2a_phase.txt
  else env $v timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2a_phase.txt; fi
done
PROBE_TAG=default timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -14 | tee gpurun_out/r2a_lanes.txt
PROBE_TAG=nocache CMAXB_FE_CACHE=0 timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -14 | tee -a gpurun_out/r2a_lanes.txt
PROBE_TAG=frac0.34 CMAXB_FE_GRID_FRACTION=0.34 timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -14 | tee -a gpurun_out/r2a_lanes.txt
