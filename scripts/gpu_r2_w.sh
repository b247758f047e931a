#!/bin/bash
# round 2, call W (1 GPU): refinement at two CTAs per SM (32 registers, two loads in flight per warp): bench only
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/w_bench_100M.json 2> $O/w_bench_100M.err; echo "exit=$?"; python -c "
import json;d=json.loads(open('$O/w_bench_100M.json').read().strip().splitlines()[-1]);print(round(d['ms_per_step'],1),{k:round(v,1) for k,v in d['phases_ms_per_step'].items()})"
