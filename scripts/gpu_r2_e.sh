#!/bin/bash
# round 2, call E (1 GPU): fourth-generation local sort (local_sort4_k) parity + A/B against local_sort3_k
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
export SGPU_A_SORTED=0
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
ph() { grep -o '"value": [0-9.]*' $1 | head -2 | tr '\n' ' '; grep -o '"phases_ms_per_step": {[^}]*}' $1; }
step "gpu suite with local_sort4_k"
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > $O/e_tests.log 2>&1; echo "exit=$?" >> $O/e_tests.log; tail -25 $O/e_tests.log
step "sanitizer: smoke"
timeout 200 compute-sanitizer --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/e_sanitizer.log 2>&1; echo "exit=$?" >> $O/e_sanitizer.log; tail -4 $O/e_sanitizer.log
for S in 1 0; do
  step "bench 20M sort4=$S"
  SGPU_SORT4=$S timeout 150 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-cpu-baseline > $O/e_bench20_sort4_$S.json 2> $O/e_bench20_sort4_$S.err; ph $O/e_bench20_sort4_$S.json; tail -2 $O/e_bench20_sort4_$S.err
done
step "bench 100M sort4=1"
SGPU_SORT4=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/e_bench100_sort4_1.json 2> $O/e_bench100_sort4_1.err; ph $O/e_bench100_sort4_1.json; tail -2 $O/e_bench100_sort4_1.err
step "done"
