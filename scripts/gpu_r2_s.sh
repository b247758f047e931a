#!/bin/bash
# round 2, call S (N GPUs): final lines. N=1: suite, whole path, default bench. N>1: distributed parity (N=2) + bench.py --gpus N as the driver launches it
N=${1:-1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
if [ "$N" = "1" ]; then
  step "gpu suite"
  timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/s1_tests.log 2>&1; echo "exit=$?" >> $O/s1_tests.log; tail -4 $O/s1_tests.log | cut -c1-250
  step "whole path reads -> GFA, 20 M reads, traced"
  SGPU_TRACE=1 timeout 900 python scripts/bench_graph.py --reads 20000000 --edge-index > $O/s1_graph_20M.json 2> $O/s1_graph_20M.err; echo "exit=$?"; cat $O/s1_graph_20M.json; grep "sgpu g" $O/s1_graph_20M.err | tail -12
  step "whole path reads -> GFA, 40 M reads"
  timeout 900 python scripts/bench_graph.py --reads 40000000 --edge-index > $O/s1_graph_40M.json 2> $O/s1_graph_40M.err; echo "exit=$?"; cat $O/s1_graph_40M.json; tail -2 $O/s1_graph_40M.err
  step "default bench"
  timeout 900 python bench.py > $O/s1_bench_default.json 2> $O/s1_bench_default.err; echo "exit=$?"; tail -c 1500 $O/s1_bench_default.json
else
  if [ "$N" = "2" ]; then
    step "2-GPU parity (tests/test_distributed.py)"
    timeout 600 python -m pytest tests/test_distributed.py -q -m gpu --timeout 500 -x > $O/s${N}_tests.log 2>&1; echo "exit=$?" >> $O/s${N}_tests.log; tail -5 $O/s${N}_tests.log
  fi
  step "bench N=$N, 100 M reads per GPU, default steps/warmup"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > $O/s${N}_bench100.json 2> $O/s${N}_bench100.err
  echo "rc=$?"; tail -4 $O/s${N}_bench100.err; tail -c 2600 $O/s${N}_bench100.json
fi
step "done"
