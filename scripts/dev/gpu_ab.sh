#!/usr/bin/env bash
# A/B timing helper: parity tests, then bench with both scan occupancy variants.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
for occ in dense free; do
  echo "== TB_SCAN_OCC=$occ"
  TB_SCAN_OCC=$occ timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['kernels_ms'], 'k1 frac %.3f' % d['roofline_k1']['frac'])
    else: print(l.rstrip())
"
done
