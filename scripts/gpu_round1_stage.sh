#!/bin/bash
# GPU session 2: CTA-major staging of the level-A output + gathering first refinement round, partition sub-ranges.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
BUDGET=${BUDGET:-330}
left() { echo $(( BUDGET - ( $(date +%s) - T0 ) )); }
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
export PYTHONUNBUFFERED=1
ph() { grep -o '"value": [0-9.]*' $1 | head -2 | tr '\n' ' '; grep -o '"phases_ms_per_step": {[^}]*}' $1; }

step "parity, staging forced"
SGPU_STAGE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 > $O/c2_tests_stage.log 2>&1
echo "exit=$?" >> $O/c2_tests_stage.log; tail -3 $O/c2_tests_stage.log
step "parity, staging + 3 sub-ranges"
SGPU_STAGE=1 SGPU_A_SUB=3 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 > $O/c2_tests_stage_sub3.log 2>&1
echo "exit=$?" >> $O/c2_tests_stage_sub3.log; tail -3 $O/c2_tests_stage_sub3.log
step "sanitizer smoke, staging + 2 sub-ranges"
SGPU_STAGE=1 SGPU_A_SUB=2 timeout 100 compute-sanitizer --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/c2_sanitizer.log 2>&1
echo "exit=$?" >> $O/c2_sanitizer.log; tail -3 $O/c2_sanitizer.log

for cfg in "1 1" "1 2" "1 4" "0 4"; do
  set -- $cfg
  [ $(left) -gt 40 ] || break
  step "bench 20M stage=$1 sub=$2"
  SGPU_STAGE=$1 SGPU_A_SUB=$2 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/c2_bench20_stage$1_sub$2.json 2> $O/c2_bench20_stage$1_sub$2.err
  ph $O/c2_bench20_stage$1_sub$2.json
done
for cfg in "1 1 4096" "1 2 4096" "1 1 8192"; do
  set -- $cfg
  [ $(left) -gt 60 ] || break
  step "bench 100M stage=$1 sub=$2 pamax=$3"
  SGPU_STAGE=$1 SGPU_A_SUB=$2 SGPU_PA_MAX=$3 timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/c2_bench100_stage$1_sub$2_pa$3.json 2> $O/c2_bench100_stage$1_sub$2_pa$3.err
  ph $O/c2_bench100_stage$1_sub$2_pa$3.json
done
step "done"
