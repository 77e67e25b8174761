#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none \
   -k regex:'second_order|coeff_velacc|scan_kernel|init_bounds|xbound' -c 20 --csv --log-file gpurun_out/launches_cfg3.csv \
   python scripts/cfg3_probe.py > gpurun_out/cfg3_under_ncu.log 2>&1
tail -2 gpurun_out/cfg3_under_ncu.log
