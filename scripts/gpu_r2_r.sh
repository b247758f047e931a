#!/bin/bash
# round 2, call R (1 GPU): sample-sort link records, device edge table: suite + whole path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/r_tests.log 2>&1; echo "exit=$?" >> $O/r_tests.log; tail -4 $O/r_tests.log | cut -c1-250
step "whole path reads -> GFA, 20 M reads, traced"
SGPU_TRACE=1 timeout 900 python scripts/bench_graph.py --reads 20000000 --edge-index > $O/r_graph_20M.json 2> $O/r_graph_20M.err; echo "exit=$?"; cat $O/r_graph_20M.json; grep "sgpu g" $O/r_graph_20M.err | tail -12
step "whole path reads -> GFA, 40 M reads"
timeout 900 python scripts/bench_graph.py --reads 40000000 --edge-index > $O/r_graph_40M.json 2> $O/r_graph_40M.err; echo "exit=$?"; cat $O/r_graph_40M.json; tail -2 $O/r_graph_40M.err
step "done"
