#!/usr/bin/env bash
# A/B of the two scan kernels at the cfg-3 shape (nC = 50) and cfg-2 shape with 6 DOF (nC = 26)
set -u
mkdir -p gpurun_out
for impl in v2 v1; do
  echo "== cfg3 scale 1/16 impl $impl"
  TB_SCAN_IMPL=$impl timeout 600 python scripts/bench_configs.py cfg3 --scale 0.0625 2>&1 | tail -1
done
for impl in v2 v1; do
  echo "== 6-dof vel+acc impl $impl"
  TB_SCAN_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dof 6 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['kernels_ms'])"
done
