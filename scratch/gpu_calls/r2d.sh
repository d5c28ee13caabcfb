#!/bin/bash
# round 2, call d: lock-step batched warp (ILP) + unroll variants; device-resident event store test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_firstparty.py tests/test_gpu_stream_device.py tests/test_optim.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2d_pytest.log
for v in default u1 u3 u4 u4c2; do
  echo "== phase stamps: $v" | tee -a gpurun_out/r2d_phase.txt
  if [ "$v" = "default" ]; then timeout 300 python scratch/phase.py 2>&1 | tail -16 | tee -a gpurun_out/r2d_phase.txt
  else CMAXB_LIB_PATH=scratch/variants/libcmax_b200_$v.so timeout 300 python scratch/phase.py 2>&1 | tail -16 | tee -a gpurun_out/r2d_phase.txt; fi
done
echo "== phase stamps: default REC=1" | tee -a gpurun_out/r2d_phase.txt
CMAXB_FE_REC=1 timeout 300 python scratch/phase.py 2>&1 | tail -8 | tee -a gpurun_out/r2d_phase.txt
for v in default u3 u4; do
  if [ "$v" = "default" ]; then PROBE_TAG=$v timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -4 | tee -a gpurun_out/r2d_lanes.txt
  else PROBE_TAG=$v CMAXB_LIB_PATH=scratch/variants/libcmax_b200_$v.so timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -4 | tee -a gpurun_out/r2d_lanes.txt; fi
done
PROBE_TAG=default_rec1 CMAXB_FE_REC=1 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -4 | tee -a gpurun_out/r2d_lanes.txt
