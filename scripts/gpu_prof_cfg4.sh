#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_robust' -s 1 -c 1 \
   -o gpurun_out/prof_r02_cfg4 -f python scripts/cfg4_probe.py > gpurun_out/ncu_full_cfg4.log 2>&1
tail -1 gpurun_out/ncu_full_cfg4.log
