#!/bin/bash
# round 2, call U (1 GPU): two-word key_bits fast path (refinement, local sort): suite + bench; coverage pre-filter at 20 M reads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > $O/u_tests.log 2>&1; echo "exit=$?" >> $O/u_tests.log; tail -4 $O/u_tests.log | cut -c1-250
step "bench 100 M"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/u_bench_100M.json 2> $O/u_bench_100M.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/u_bench_100M.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
step "coverage pre-filter + whole path, 20 M reads, threshold 3"
timeout 900 python scripts/bench_graph.py --reads 20000000 --cov-threshold 3 > $O/u_graph_cov_20M.json 2> $O/u_graph_cov_20M.err; echo "exit=$?"; cat $O/u_graph_cov_20M.json | cut -c1-900; tail -2 $O/u_graph_cov_20M.err
step "done"
