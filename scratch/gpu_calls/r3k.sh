#!/bin/bash
# round 2, call 3k (1 GPU): ncu --set full of the back-end kernels at full C4 size (1e7 events), graph replay off so that the kernels are visible by name
mkdir -p gpurun_out
CMAXB_BE_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"be_scatter_kernel|be_gather_kernel|be_pose_kernel|be_grad_reduce|blur_reduce|adjoint_blur" -s 24 -c 6 -o gpurun_out/r3k_be -f python scratch/prof_be.py 1.0 > gpurun_out/r3k_ncu.log 2>&1
tail -3 gpurun_out/r3k_ncu.log
ls -la gpurun_out/r3k_be.ncu-rep
