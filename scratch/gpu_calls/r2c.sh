#!/bin/bash
# round 2, call c: pipelined gather; ncu source-level capture of the fused kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_firstparty.py tests/test_gpu_sharded.py tests/test_gsl_adapter.py tests/test_optim.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.log
for v in "default" "CMAXB_FE_REC=1"; do
  echo "== phase stamps: $v" | tee -a gpurun_out/r2c_phase.txt
  if [ "$v" = "default" ]; then timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2c_phase.txt
  else env $v timeout 300 python scratch/phase.py 2>&1 | tail -14 | tee -a gpurun_out/r2c_phase.txt; fi
done
PROBE_TAG=rec0 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -6 | tee gpurun_out/r2c_lanes.txt
PROBE_TAG=rec1 CMAXB_FE_REC=1 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -6 | tee -a gpurun_out/r2c_lanes.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_eval_fused -s 12 -c 1 -o gpurun_out/r2c_fused -f python scratch/prof_fe.py > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
