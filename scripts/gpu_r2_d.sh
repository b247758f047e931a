#!/bin/bash
# round 2, call D (1 GPU): the third-generation partition kernel (levelA_scatter_sorted_k) and the smem sort primitive
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
ph() { grep -o '"value": [0-9.]*' $1 | head -2 | tr '\n' ' '; grep -o '"phases_ms_per_step": {[^}]*}' $1; }
step "sort primitive + arithmetic"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k device_arithmetic --timeout 250 > $O/d_prim.log 2>&1; echo "exit=$?" >> $O/d_prim.log; tail -12 $O/d_prim.log
step "gpu suite (sorted partition kernel is the default)"
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > $O/d_tests.log 2>&1; echo "exit=$?" >> $O/d_tests.log; tail -25 $O/d_tests.log
step "sanitizer: smoke"
timeout 200 compute-sanitizer --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_sanitizer.log 2>&1; echo "exit=$?" >> $O/d_sanitizer.log; tail -4 $O/d_sanitizer.log
step "sanitizer racecheck: smoke"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_racecheck.log 2>&1; echo "exit=$?" >> $O/d_racecheck.log; tail -6 $O/d_racecheck.log
for S in 1 0; do
  step "bench 20M sorted=$S"
  SGPU_A_SORTED=$S timeout 150 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-cpu-baseline > $O/d_bench20_sorted$S.json 2> $O/d_bench20_sorted$S.err; ph $O/d_bench20_sorted$S.json; tail -2 $O/d_bench20_sorted$S.err
done
for S in 1 0; do
  step "bench 100M sorted=$S"
  SGPU_A_SORTED=$S timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/d_bench100_sorted$S.json 2> $O/d_bench100_sorted$S.err; ph $O/d_bench100_sorted$S.json; tail -2 $O/d_bench100_sorted$S.err
done
step "done"
