#!/bin/bash
mkdir -p gpurun_out
ncu --set full --cache-control none --clock-control none --import-source on -k regex:"blur_reduce|adjoint_blur|fe_gather|fe_scatter" -s 16 -c 4 -o gpurun_out/prof_fe_adjoint -f python scratch/prof_fe.py > gpurun_out/ncu_prof.log 2>&1
tail -3 gpurun_out/ncu_prof.log
