#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2e.txt
bash scripts/gpu_quick.sh r2e
