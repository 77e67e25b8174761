#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (TB_SCAN_PAIR=2)"; TB_SCAN_PAIR=2 timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
for pair in 2 0; do
  echo "== bench TB_SCAN_PAIR=$pair"
  TB_SCAN_PAIR=$pair timeout 600 python bench.py --steps 20 --warmup 5 --configs 5 --no-cpu-baseline 2>>gpurun_out/bench.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernels_ms'], 'e2e', d['e2e']['value'], 'fast', d['opt_in_fast_lower_bound']['value'], 'cfg5', d['configs'].get('cfg5',{}).get('paths_per_s'))"
done
tail -3 gpurun_out/bench.err
