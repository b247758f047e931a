#!/bin/bash
# round 2, call X (1 GPU): final check of the committed tree: smoke + GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > $O/x_tests.log 2>&1; echo "exit=$?" >> $O/x_tests.log; tail -4 $O/x_tests.log | cut -c1-250
