#!/bin/bash
# round 2, call K (1 GPU): warp tiles with the one-round-trip staging copy: bench, then ncu --set full of the new level-A kernels and the refinement at 10 M reads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "bench 100 M"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/k_bench_100M.json 2> $O/k_bench_100M.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/k_bench_100M.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
step "ncu --set full at 10 M reads (arena 40 GB): count / scatter x2 / refine"
SGPU_ARENA_GB=40 timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"levelA_count_roll_k|levelA_scatter_roll_k|refine_k" -c 4 -o $O/k_full_10M python bench.py --reads 10000000 --steps 1 --warmup 0 --no-cpu-baseline > $O/k_ncu_full.log 2>&1; echo "exit=$?"
ls -la $O/k_full_10M.ncu-rep
step "done"
