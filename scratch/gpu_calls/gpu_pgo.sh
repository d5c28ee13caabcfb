#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 200 python -m pytest tests/test_pgo.py -q -m gpu 2>&1 | tail -3; done | tee gpurun_out/pytest_pgo.log
