#!/bin/bash
# round 2, call 3a (1 GPU): run-to-run variation of the FE solve, current library vs the one with the previous binning kernels
echo "== current"; timeout 300 python scratch/fe_solve_repeat.py 2>&1 | tail -12
echo "== previous binning (3 kernels, memsets)"; CMAXB_LIB_PATH=scratch/variants/libcmax_b200_oldbin.so timeout 300 python scratch/fe_solve_repeat.py 2>&1 | tail -12
