#!/bin/bash
# round 2, call V (1 GPU): contiguous refinement back to one load per thread, two-word key_bits kept: suite + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > $O/v_tests.log 2>&1; echo "exit=$?" >> $O/v_tests.log; tail -4 $O/v_tests.log | cut -c1-250
step "bench 100 M"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/v_bench_100M.json 2> $O/v_bench_100M.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/v_bench_100M.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
step "done"
