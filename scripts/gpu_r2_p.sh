#!/bin/bash
# round 2, call P (1 GPU): coverage pre-filter (8f-3) parity, host-parallel GFA writer / edge-index packing, whole-path timing with the graph phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "coverage pre-filter tests first"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k coverage_prefilter 2>&1 | tail -15 | cut -c1-250
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/p_tests.log 2>&1; echo "exit=$?" >> $O/p_tests.log; tail -6 $O/p_tests.log | cut -c1-250
step "whole path reads -> GFA, 20 M reads, traced"
SGPU_TRACE=1 timeout 900 python scripts/bench_graph.py --reads 20000000 --edge-index > $O/p_graph_20M.json 2> $O/p_graph_20M.err; echo "exit=$?"; cat $O/p_graph_20M.json; grep "sgpu graph" $O/p_graph_20M.err | tail -12
step "done"
