#!/bin/bash
# round 2, call I/J (1 GPU): level-A instruction diet, then warp-autonomous tiles: suite + benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/j_tests.log 2>&1; echo "exit=$?" >> $O/j_tests.log; tail -8 $O/j_tests.log | cut -c1-250
step "bench 100 M"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/j_bench_100M.json 2> $O/j_bench_100M.err; echo "exit=$?"; cut -c1-1800 $O/j_bench_100M.json
step "bench 20 M"
timeout 300 python bench.py --reads 20000000 --steps 5 --warmup 3 --no-cpu-baseline > $O/j_bench_20M.json 2> $O/j_bench_20M.err; echo "exit=$?"; cut -c1-1200 $O/j_bench_20M.json
step "done"
