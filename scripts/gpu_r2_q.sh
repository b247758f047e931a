#!/bin/bash
# round 2, call Q (1 GPU): whole-path timing after the parallel file write / vertex sort, traced; suite (the 1 M-read SHA-256 fixtures cover the threaded GFA path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/q_tests.log 2>&1; echo "exit=$?" >> $O/q_tests.log; tail -4 $O/q_tests.log | cut -c1-250
step "whole path reads -> GFA, 20 M reads, traced"
SGPU_TRACE=1 timeout 900 python scripts/bench_graph.py --reads 20000000 --edge-index > $O/q_graph_20M.json 2> $O/q_graph_20M.err; echo "exit=$?"; cat $O/q_graph_20M.json; grep "sgpu g" $O/q_graph_20M.err | tail -12
step "whole path reads -> GFA, 40 M reads"
timeout 900 python scripts/bench_graph.py --reads 40000000 --edge-index > $O/q_graph_40M.json 2> $O/q_graph_40M.err; echo "exit=$?"; cat $O/q_graph_40M.json; tail -2 $O/q_graph_40M.err
step "done"
