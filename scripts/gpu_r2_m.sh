#!/bin/bash
# round 2, call M (1 GPU): store of a window delayed by one window (atomic latency behind the next roll): suite, bench, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/m_tests.log 2>&1; echo "exit=$?" >> $O/m_tests.log; tail -4 $O/m_tests.log | cut -c1-250
step "bench 100 M"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/m_bench_100M.json 2> $O/m_bench_100M.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/m_bench_100M.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
step "ncu --set full at 10 M reads (arena 40 GB): count / scatter"
SGPU_ARENA_GB=40 timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"levelA_count_roll_k|levelA_scatter_roll_k" -c 2 -o $O/m_full_10M python bench.py --reads 10000000 --steps 1 --warmup 0 --no-cpu-baseline > $O/m_ncu_full.log 2>&1; echo "exit=$?"
ls -la $O/m_full_10M.ncu-rep
step "done"
