#!/bin/bash
# round 2, call y (1 GPU): the failing optimiser-outcome test in detail (3 repetitions), then the whole GPU suite without -x
for i in 1 2 3; do timeout 300 python -m pytest tests/test_optim.py -m gpu -q -k test_library_fe_solve 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -12; done
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
