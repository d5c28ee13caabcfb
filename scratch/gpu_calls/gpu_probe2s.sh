#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python scratch/two_stream_probe.py; CMAXB_FE_GRID_FRACTION=0.5 timeout 100 python scratch/two_stream_probe.py; CMAXB_FE_GRID_FRACTION=0.34 timeout 100 python scratch/two_stream_probe.py ) 2>&1 | grep fraction | tee gpurun_out/two_stream_probe.txt
