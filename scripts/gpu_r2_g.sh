#!/bin/bash
# round 2, call G (1 GPU): new rows (A/T clipper, EdgeIndex refill) on the GPU, then the ncu evidence for the roofline block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > $O/g_tests.log 2>&1; echo "exit=$?" >> $O/g_tests.log; tail -25 $O/g_tests.log
step "sanitizer: A/T clipper + edge index on a small case"
timeout 300 compute-sanitizer --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rna_k21 or loops_k21_B10_eigraph" --timeout 250 > $O/g_sanitizer.log 2>&1; echo "exit=$?" >> $O/g_sanitizer.log; tail -4 $O/g_sanitizer.log
step "ncu launch list (100 M reads, arena capped so that ncu keeps memory for its replay buffers)"
SGPU_ARENA_GB=100 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sg:: -c 300 --csv --log-file $O/g_launches_100M.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/g_ncu_launch.log 2>&1; echo "exit=$?"; tail -2 $O/g_ncu_launch.log | cut -c1-300
step "ncu --set full: first pass's count / scatter / refine / sort / compact (100 M reads)"
SGPU_ARENA_GB=100 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"levelA_count_roll_k|levelA_scatter_roll_k|refine_k|local_sort3_k|compact_k" -c 5 -o $O/g_full_100M python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/g_ncu_full.log 2>&1; echo "exit=$?"; tail -3 $O/g_ncu_full.log | cut -c1-300; ls -la $O/g_full_100M.ncu-rep
step "done"
