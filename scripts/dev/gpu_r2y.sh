#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
for minb in 8192 0; do
  echo "== bench TB_SCAN_FWD_THREADS_MIN=$minb"
  TB_SCAN_FWD_THREADS_MIN=$minb timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --configs 5 2>>gpurun_out/bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernels_ms'], 'cfg5', d['configs']['cfg5']['paths_per_s'], d['configs']['cfg5']['ms_per_step'])"
done
echo "== batch sweep: 8192 / 16384 paths with and without"
for b in 8192 16384; do for minb in 8192 0; do
  TB_SCAN_FWD_THREADS_MIN=$minb timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --configs none --batch $b 2>>gpurun_out/bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$b min=$minb', d['value'], d['kernels_ms']['K2_scan_velacc'])"
done; done
tail -3 gpurun_out/bench.err
