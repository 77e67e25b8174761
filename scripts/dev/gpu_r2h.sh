#!/usr/bin/env bash
# Round-2 checkpoint call: parity tests, bench (both arms), ncu launch list, ncu full capture of the scan kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 4 -c 2 \
   -o gpurun_out/prof_r02h -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
