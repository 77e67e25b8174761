#!/usr/bin/env bash
# Round-2 final evidence on one B200: parity tests, smoke, bench (both arms), ncu launch list, ncu full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench.err > gpurun_out/bench_ref.json; cut -c1-200 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>>gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_r02.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches_r02.csv
echo "== ncu full: fused scan (full kernel), then the split kernels of the e2e arm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 3 -c 1 \
   -o gpurun_out/prof_r02_scan -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_full_scan.log 2>&1
tail -1 gpurun_out/ncu_full_scan.log
echo "== ncu full: large-batch forward pass (one thread per path) and the backward-only warp kernel at 32768 paths"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'forward_threads|scan_kernel' -s 4 -c 2 \
   -o gpurun_out/prof_r02_fwd -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none --batch 32768 > gpurun_out/ncu_full_fwd.log 2>&1
tail -1 gpurun_out/ncu_full_fwd.log
echo "== ncu full: cfg 3 kernels"
B=4096 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'second_order|coeff_velacc|scan_kernel' -s 3 -c 3 \
   -o gpurun_out/prof_r02_cfg3 -f python scripts/cfg3_probe.py > gpurun_out/ncu_full_cfg3.log 2>&1
tail -1 gpurun_out/ncu_full_cfg3.log
ls -la gpurun_out | tail -12
