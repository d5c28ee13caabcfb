#!/bin/bash
# round 2, call 3b (1 GPU): the whole GPU suite three times (flake hunt), then smoke()
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
