#!/bin/bash
# One GPU-box session for the rolling level-A kernels: parity first, then timings old vs new, then profiles.
# Every step is bounded; later steps are skipped once the session budget is used up.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
BUDGET=${BUDGET:-460}
left() { echo $(( BUDGET - ( $(date +%s) - T0 ) )); }
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
export PYTHONUNBUFFERED=1

step "parity, rolling kernels forced"
SGPU_ROLL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 > $O/c1_tests_roll.log 2>&1
echo "exit=$?" >> $O/c1_tests_roll.log
tail -4 $O/c1_tests_roll.log

step "parity, rolling kernels without the id array"
SGPU_ROLL=1 SGPU_NO_IDS=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 100 -k "oracle_random or medium or ragged" > $O/c1_tests_roll_noids.log 2>&1
echo "exit=$?" >> $O/c1_tests_roll_noids.log
tail -3 $O/c1_tests_roll_noids.log

for roll in 1 0; do
  [ $(left) -gt 60 ] || break
  step "bench 20M roll=$roll"
  SGPU_ROLL=$roll timeout 150 python bench.py --reads 20000000 --no-cpu-baseline > $O/c1_bench20_roll$roll.json 2> $O/c1_bench20_roll$roll.err
  tail -c 1500 $O/c1_bench20_roll$roll.json | grep -o '"phases_ms_per_step": {[^}]*}'
done

if [ $(left) -gt 120 ]; then
  step "bench 100M roll=1"
  SGPU_ROLL=1 timeout 240 python bench.py --no-cpu-baseline > $O/c1_bench100_roll1.json 2> $O/c1_bench100_roll1.err
  grep -o '"value": [0-9.]*\|"phases_ms_per_step": {[^}]*}' $O/c1_bench100_roll1.json | head -3
fi

if [ $(left) -gt 80 ]; then
  step "ncu launch list 20M roll=1"
  SGPU_ROLL=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c1_launches_20M_roll1.csv \
      python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/c1_ncu_launch.log 2>&1
fi

if [ $(left) -gt 100 ]; then
  step "ncu full, rolling kernels, 5M reads"
  SGPU_ROLL=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:roll_k -c 2 -o $O/c1_roll_5M -f \
      python bench.py --reads 5000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/c1_ncu_full.log 2>&1
fi

if [ $(left) -gt 70 ]; then
  step "compute-sanitizer smoke, rolling"
  SGPU_ROLL=1 timeout 120 compute-sanitizer --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/c1_sanitizer.log 2>&1
  echo "exit=$?" >> $O/c1_sanitizer.log
  tail -3 $O/c1_sanitizer.log
fi
step "done"
