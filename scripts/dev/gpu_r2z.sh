#!/usr/bin/env bash
set -u
for b in 32768 65536; do for minb in 8192 0; do
  TB_SCAN_FWD_THREADS_MIN=$minb timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --configs none --batch $b 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$b min=$minb', d['value'], d['kernels_ms']['K2_scan_velacc'])"
done; done
