#!/bin/bash
# round 2, call B (1 GPU): the whole GPU suite on the cleaned-up code (new: forced multi-pass, super-ranges, 1 M-read SHA-256
# fixtures, un-skipped tests), then one 100 M-read bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > $O/b_tests.log 2>&1; echo "exit=$?" >> $O/b_tests.log; tail -15 $O/b_tests.log
step "bench 100M"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/b_bench100.json 2> $O/b_bench100.err; tail -3 $O/b_bench100.err; cat $O/b_bench100.json | head -c 3000
step "done"
