#!/usr/bin/env bash
# ncu capture of the two-paths-per-warp experiments (TB_SCAN_PAIR=1 divergent half-warps, =2 lockstep)
set -u
mkdir -p gpurun_out
MODE=${1:-2}
TB_SCAN_PAIR=$MODE timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'scan_lock_kernel|scan_pair_kernel' -s 3 -c 1 \
   -o gpurun_out/prof_pair$MODE -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_full_pair.log 2>&1
tail -2 gpurun_out/ncu_full_pair.log
