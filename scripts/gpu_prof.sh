#!/usr/bin/env bash
# ncu full capture of K1/K2 (one launch each after warm-up) + launch list.  usage: gpu_prof.sh <tag>
set -u
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|coeff_velacc' -s 6 -c 2 \
   -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -8
