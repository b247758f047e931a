#!/bin/bash
# round 2, call Z (1 GPU): ncu launch list of the final kernels (bench.py at 20 M reads, one pass): per-launch durations for profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 140 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:sg:: -c 700 --csv --log-file gpurun_out/z_launches_20M.csv python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/z_ncu_launch.log 2>&1; echo "exit=$?"
tail -c 600 gpurun_out/z_ncu_launch.log
