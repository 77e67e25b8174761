#!/usr/bin/env bash
# Round-2 final evidence, part B: ncu full of the large-batch forward pass and of the cfg-3 scan
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'forward_threads' -s 2 -c 1 \
   -o gpurun_out/prof_r02_fwd -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none --batch 32768 > gpurun_out/ncu_full_fwd.log 2>&1
tail -1 gpurun_out/ncu_full_fwd.log
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 1 -c 1 \
   -o gpurun_out/prof_r02_cfg3scan -f python scripts/cfg3_probe.py > gpurun_out/ncu_full_cfg3.log 2>&1
tail -1 gpurun_out/ncu_full_cfg3.log
ls -la gpurun_out | tail -6
