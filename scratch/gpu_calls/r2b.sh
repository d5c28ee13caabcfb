#!/bin/bash
# round 2, call b: fused kernel with TMA-staged LUT tiles, pipelined event loops, register-tiled image phase, tagged result words
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.log
for v in "default" "CMAXB_FE_REC=1" "CMAXB_FE_TMA=0"; do
  echo "== phase stamps: $v" | tee -a gpurun_out/r2b_phase.txt
  if [ "$v" = "default" ]; then timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2b_phase.txt
  else env $v timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2b_phase.txt; fi
done
PROBE_TAG=rec0 timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -14 | tee gpurun_out/r2b_lanes.txt
PROBE_TAG=rec1 CMAXB_FE_REC=1 timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -14 | tee -a gpurun_out/r2b_lanes.txt
PROBE_TAG=rec0_frac0.34 CMAXB_FE_GRID_FRACTION=0.34 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -14 | tee -a gpurun_out/r2b_lanes.txt
PROBE_TAG=rec1_frac0.34 CMAXB_FE_REC=1 CMAXB_FE_GRID_FRACTION=0.34 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -14 | tee -a gpurun_out/r2b_lanes.txt
PROBE_TAG=rec0_frac0.67 CMAXB_FE_GRID_FRACTION=0.67 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -14 | tee -a gpurun_out/r2b_lanes.txt
