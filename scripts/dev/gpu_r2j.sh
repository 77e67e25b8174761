#!/usr/bin/env bash
# parity tests (all) + K2 occupancy A/B on the fused scan
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
for occ in 32 28; do
  echo "== bench TB_SCAN_FUSED_OCC=$occ"
  TB_SCAN_FUSED_OCC=$occ timeout 600 python bench.py --steps 20 --warmup 5 --configs none --no-cpu-baseline 2>>gpurun_out/bench.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernels_ms'], d['e2e']['value'], d['records_path']['kernels_ms'])"
done
