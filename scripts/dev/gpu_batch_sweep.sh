#!/usr/bin/env bash
for b in 1024 2048 4096 8192 16384 65536; do
  for occ in dense free; do
  TB_SCAN_OCC=$occ timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('B=$b occ=$occ value %.0f e2e %.0f K1 %.3f ms K2 %.3f ms  K2-only %.0f paths/s' % (d['value'], d['e2e']['value'], d['kernels_ms']['K1_coeff'], d['kernels_ms']['K2_scan'], $b/d['kernels_ms']['K2_scan']*1e3))
"
  done
done
