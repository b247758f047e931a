#!/bin/bash
# last seconds of the round-1 GPU budget: one ncu --set full capture of the staged partition kernel + the gathering refinement
# round (small input; the arena is capped so that ncu has device memory left for its save/restore buffers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SGPU_ARENA_GB=40 SGPU_A_SUB=1 timeout 60 ncu --set full --clock-control none --import-source on -k regex:'levelA_scatter_roll_k|refine_k' -c 2 -f -o gpurun_out/c4_staged_4M \
    python bench.py --reads 4000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4_ncu.log | cut -c1-200
