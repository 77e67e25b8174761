#!/usr/bin/env bash
# Round-2 final evidence, part A: bench (both arms), ncu launch list, ncu full of the fused scan
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench.err > gpurun_out/bench_ref.json; cut -c1-200 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>>gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-200 gpurun_out/bench.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_r02.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches_r02.csv
echo "== ncu full: fused scan"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 3 -c 1 \
   -o gpurun_out/prof_r02_scan -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_full_scan.log 2>&1
tail -1 gpurun_out/ncu_full_scan.log
ls -la gpurun_out | tail -8
