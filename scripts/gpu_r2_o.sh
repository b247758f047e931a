#!/bin/bash
# round 2, call O (1 GPU): the lines the driver produces -- default bench (with cpu_baseline), reference arm -- and the whole-path reads -> GFA measurement
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "host"
nproc; free -g | head -2; df -h /tmp | tail -1
step "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
step "default bench"
timeout 900 python bench.py > $O/o_bench_default.json 2> $O/o_bench_default.err; echo "exit=$?"; tail -c 3000 $O/o_bench_default.json; tail -3 $O/o_bench_default.err
step "reference arm (--steps 3 --warmup 1)"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/o_bench_reference.json 2> $O/o_bench_reference.err; echo "exit=$?"; tail -c 2500 $O/o_bench_reference.json; tail -3 $O/o_bench_reference.err
step "whole path reads -> GFA, 20 M reads"
timeout 600 python scripts/bench_graph.py --reads 20000000 --edge-index > $O/o_graph_20M.json 2> $O/o_graph_20M.err; echo "exit=$?"; cat $O/o_graph_20M.json; tail -3 $O/o_graph_20M.err
step "whole path reads -> GFA, 40 M reads"
timeout 600 python scripts/bench_graph.py --reads 40000000 > $O/o_graph_40M.json 2> $O/o_graph_40M.err; echo "exit=$?"; cat $O/o_graph_40M.json; tail -3 $O/o_graph_40M.err
step "done"
