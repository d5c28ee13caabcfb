#!/bin/bash
# round 2, call z (1 GPU): the optimiser-outcome tests 15 times each (run-to-run variation of the device cost)
for i in $(seq 1 15); do timeout 300 python -m pytest tests/test_optim.py tests/test_gsl_adapter.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | head -6; done
