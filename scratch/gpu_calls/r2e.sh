#!/bin/bash
# round 2, call e: border term in registers, 17-tap adjoint image phase (S1, S2 by adjointness)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_firstparty.py tests/test_gpu_stream_device.py tests/test_optim.py tests/test_gsl_adapter.py tests/test_end_to_end.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2e_pytest.log
echo "== phase stamps: default" | tee -a gpurun_out/r2e_phase.txt
timeout 300 python scratch/phase.py 2>&1 | tail -16 | tee -a gpurun_out/r2e_phase.txt
echo "== phase stamps: default REC=1" | tee -a gpurun_out/r2e_phase.txt
CMAXB_FE_REC=1 timeout 300 python scratch/phase.py 2>&1 | tail -16 | tee -a gpurun_out/r2e_phase.txt
PROBE_TAG=rec0 timeout 600 python scratch/fe_lanes_probe.py 2>&1 | tail -10 | tee -a gpurun_out/r2e_lanes.txt
PROBE_TAG=rec1 CMAXB_FE_REC=1 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -4 | tee -a gpurun_out/r2e_lanes.txt
PROBE_TAG=rec0_frac0.34 CMAXB_FE_GRID_FRACTION=0.34 timeout 600 python scratch/fe_lanes_probe.py short 2>&1 | tail -4 | tee -a gpurun_out/r2e_lanes.txt
