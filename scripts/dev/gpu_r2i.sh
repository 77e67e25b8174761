#!/usr/bin/env bash
# parity tests (all, no -x) + bench with the configs block
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
