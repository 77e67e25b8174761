#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'scan_pair_kernel' -s 3 -c 1 \
   -o gpurun_out/prof_pair -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_full_pair.log 2>&1
tail -2 gpurun_out/ncu_full_pair.log
