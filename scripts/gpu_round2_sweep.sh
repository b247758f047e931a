#!/bin/bash
# First GPU session of round 2: correctness of the opt-in variants prepared at the end of round 1, then the parameter sweep that
# DESIGN.md 6.1 motivates (<= 512 open streams per CTA at every level; sector pairing). Every step is bounded.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
BUDGET=${BUDGET:-540}
left() { echo $(( BUDGET - ( $(date +%s) - T0 ) )); }
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
export PYTHONUNBUFFERED=1
ph() { grep -o '"value": [0-9.]*' $1 | head -2 | tr '\n' ' '; grep -o '"phases_ms_per_step": {[^}]*}' $1; }

#step "microbenchmark: streams per CTA x store width x layout"
#[ -x scripts/microbench/_build/scatter_bench ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/_build/scatter_bench scripts/microbench/scatter_bench.cu
#timeout 90 scripts/microbench/_build/scatter_bench -g 8 > $O/s1_microbench.txt 2>&1; tail -12 $O/s1_microbench.txt
step "gpu suite, defaults"
timeout 200 python -m pytest tests -q -m gpu --timeout 150 > $O/s1_tests.log 2>&1; echo "exit=$?" >> $O/s1_tests.log; tail -3 $O/s1_tests.log
step "first GPU run of the gbuilder-style adapter tool"
SGPU_RUN_NEW=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "long_reads or EarlyPairedInfo" --timeout 180 > $O/s1_tests_long.log 2>&1; echo "exit=$?" >> $O/s1_tests_long.log; tail -3 $O/s1_tests_long.log
SGPU_RUN_NEW=1 timeout 300 python -m pytest tests/test_integration_tool.py -q -m gpu --timeout 280 > $O/s1_tests_gbuilder.log 2>&1; echo "exit=$?" >> $O/s1_tests_gbuilder.log; tail -5 $O/s1_tests_gbuilder.log
step "parity, sector pairing"
SGPU_PAIR=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_pair.log 2>&1; echo "exit=$?" >> $O/s1_tests_pair.log; tail -3 $O/s1_tests_pair.log
step "parity, sector pairing in the refinement too (+ prefetch)"
SGPU_PAIR=1 SGPU_PAIR_REFINE=1 SGPU_PREFETCH=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_pair_refine.log 2>&1; echo "exit=$?" >> $O/s1_tests_pair_refine.log; tail -3 $O/s1_tests_pair_refine.log
step "parity, three-level split (RMAX=7, PA_MAX=1280)"
SGPU_RMAX=7 SGPU_PA_MAX=1280 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_rmax7.log 2>&1; echo "exit=$?" >> $O/s1_tests_rmax7.log; tail -3 $O/s1_tests_rmax7.log
step "parity, staging for every source (k-mers from (k+1)-mers, all-windows mode)"
SGPU_STAGE_ALL=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_stage_all.log 2>&1; echo "exit=$?" >> $O/s1_tests_stage_all.log; tail -3 $O/s1_tests_stage_all.log
step "parity, 1024 local-sort bins"
SGPU_BINBITS=10 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_bins10.log 2>&1; echo "exit=$?" >> $O/s1_tests_bins10.log; tail -3 $O/s1_tests_bins10.log
step "parity, 1024-record local-sort segments, 1024 bins"
SGPU_SORT_CAP=1024 SGPU_BINBITS=10 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 150 > $O/s1_tests_cap1024.log 2>&1; echo "exit=$?" >> $O/s1_tests_cap1024.log; tail -3 $O/s1_tests_cap1024.log
step "sanitizer smoke, pairing"
SGPU_PAIR=1 timeout 100 compute-sanitizer --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/s1_sanitizer_pair.log 2>&1; echo "exit=$?" >> $O/s1_sanitizer_pair.log; tail -2 $O/s1_sanitizer_pair.log

# 20 M reads: pair x sub-ranges, rmax
for cfg in "0 4 11 4096" "1 4 11 4096" "1 2 11 4096" "1 3 11 4096" "0 4 9 4096" "0 4 7 1280" "1 4 7 1280"; do
  set -- $cfg
  [ $(left) -gt 40 ] || break
  step "bench 20M pair=$1 sub=$2 rmax=$3 pamax=$4"
  f=$O/s1_bench20_pair$1_sub$2_rmax$3_pa$4.json
  SGPU_PAIR=$1 SGPU_A_SUB=$2 SGPU_RMAX=$3 SGPU_PA_MAX=$4 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $f 2> ${f%.json}.err
  ph $f
done
[ $(left) -gt 40 ] && { step "bench 20M pair + L2 prefetch of the next id row"; SGPU_PAIR=1 SGPU_PREFETCH=1 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/s1_bench20_pair_prefetch.json 2> $O/s1_bench20_pair_prefetch.err; ph $O/s1_bench20_pair_prefetch.json; }
[ $(left) -gt 40 ] && { step "bench 20M binbits=10"; SGPU_BINBITS=10 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/s1_bench20_bins10.json 2> $O/s1_bench20_bins10.err; ph $O/s1_bench20_bins10.json; }
[ $(left) -gt 40 ] && { step "bench 20M binbits=10 target=7/8"; SGPU_BINBITS=10 SGPU_TARGET_8THS=7 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/s1_bench20_bins10_t7.json 2> $O/s1_bench20_bins10_t7.err; ph $O/s1_bench20_bins10_t7.json; }
[ $(left) -gt 40 ] && { step "bench 20M sort cap=1024 binbits=10"; SGPU_SORT_CAP=1024 SGPU_BINBITS=10 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/s1_bench20_cap1024_bins10.json 2> $O/s1_bench20_cap1024_bins10.err; ph $O/s1_bench20_cap1024_bins10.json; }
[ $(left) -gt 40 ] && { step "bench 20M sort cap=1024 binbits=11"; SGPU_SORT_CAP=1024 timeout 100 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline > $O/s1_bench20_cap1024.json 2> $O/s1_bench20_cap1024.err; ph $O/s1_bench20_cap1024.json; }
# 100 M reads
for cfg in "0 11 4096" "1 11 4096" "2 11 4096" "0 7 1280" "1 7 1280" "0 9 4096" "0 8 2560"; do
  set -- $cfg          # pair: 0 = off, 1 = level A, 2 = level A + refinement
  [ $(left) -gt 60 ] || break
  step "bench 100M pair=$1 rmax=$2 pamax=$3"
  f=$O/s1_bench100_pair$1_rmax$2_pa$3.json
  SGPU_PAIR=$(( $1 > 0 )) SGPU_PAIR_REFINE=$(( $1 > 1 )) SGPU_RMAX=$2 SGPU_PA_MAX=$3 timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $f 2> ${f%.json}.err
  ph $f
done
step "done"
