#!/bin/bash
# GPU session 3 (last of round 1): the whole GPU suite on the final defaults (incl. early tip clipper + reference-side adapter tool),
# the ncu launch list of our kernels, the default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
export PYTHONUNBUFFERED=1
step "gpu suite"
timeout 200 python -m pytest tests -q -m gpu --timeout 150 > $O/c3_tests.log 2>&1
echo "exit=$?" >> $O/c3_tests.log; tail -15 $O/c3_tests.log
step "ncu launch list, 20M, our kernels"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sg:: -c 600 --csv --log-file $O/c3_launches_20M.csv \
    python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/c3_ncu_launch.log 2>&1
tail -2 $O/c3_ncu_launch.log | cut -c1-300
step "bench default (100M)"
timeout 200 python bench.py > $O/c3_bench100.json 2> $O/c3_bench100.err
cut -c1-400 $O/c3_bench100.json
step "done"
