#!/usr/bin/env bash
# round 2, call A: parity of the sub-warp scan kernel + A/B timing against the round-1 kernel
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (no -x)"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_r2a.txt
echo "== bench new"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_r2a_new.json
tail -3 gpurun_out/bench.err
echo "== bench v1"; TB_SCAN_IMPL=v1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_r2a_v1.json
echo "== bench new B=65536"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 65536 2>gpurun_out/bench.err | tee gpurun_out/bench_r2a_new_64k.json
echo "== bench v1 B=65536"; TB_SCAN_IMPL=v1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 65536 2>gpurun_out/bench.err | tee gpurun_out/bench_r2a_v1_64k.json
