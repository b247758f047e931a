#!/bin/bash
# round 2, call N (1 GPU): CTAs per SM of the level-A grid (tail balance) at 100 M reads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
for c in 3 4 6; do
step "bench 100 M, SGPU_A_CTAS=$c"
SGPU_A_CTAS=$c timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/n_bench_100M_c$c.json 2> $O/n_bench_100M_c$c.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/n_bench_100M_c$c.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
done
step "parity at 4 CTAs per SM"
SGPU_A_CTAS=4 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x 2>&1 | tail -3
step "done"
