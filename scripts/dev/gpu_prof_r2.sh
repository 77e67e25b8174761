#!/usr/bin/env bash
# Round-2 evidence: ncu launch list of the bench command + ncu full capture of the dominant kernels + sanitizer.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/bench_under_ncu_$TAG.log 2>&1
grep -c . gpurun_out/launches_$TAG.csv
echo "== ncu full (fused scan, xbound, records scan, coeff_velacc)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|coeff_velacc|xbound_velocity' -s 8 -c 6 \
   -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/ | tail -6
