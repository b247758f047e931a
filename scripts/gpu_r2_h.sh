#!/bin/bash
# round 2, call H (1 GPU): GPU suite after the A/T clipper fix (+ packer, clients), then ncu evidence with the id array on
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
step "gpu suite"
timeout 700 python -m pytest tests -q -m gpu --timeout 200 -x > $O/h_tests.log 2>&1; echo "exit=$?" >> $O/h_tests.log; tail -25 $O/h_tests.log
step "ncu launch list, 100 M reads, arena 100 GB (ids on)"
SGPU_ARENA_GB=100 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sg:: -c 600 --csv --log-file $O/h_launches_100M.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/h_ncu_launch.log 2>&1; echo "exit=$?"
step "ncu reduced sections at 100 M reads: scatter / refine / sort of the first pass"
SGPU_ARENA_GB=100 timeout 420 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --clock-control none --kernel-name-base demangled -k regex:"levelA_scatter_roll_k|refine_k|local_sort3_k" -c 4 -o $O/h_sections_100M python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/h_ncu_sections.log 2>&1; echo "exit=$?"; grep -v "^{" $O/h_ncu_sections.log | tail -8 | cut -c1-200
step "ncu --set full at 10 M reads (arena 40 GB): count / scatter / refine / sort / compact"
SGPU_ARENA_GB=40 timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"levelA_count_roll_k|levelA_scatter_roll_k|refine_k|local_sort3_k|compact_k" -c 6 -o $O/h_full_10M python bench.py --reads 10000000 --steps 1 --warmup 0 --no-cpu-baseline > $O/h_ncu_full.log 2>&1; echo "exit=$?"; grep -v "^{" $O/h_ncu_full.log | tail -8 | cut -c1-200
ls -la $O/*.ncu-rep
step "done"
