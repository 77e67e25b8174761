#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for occ in 0 16 20 24; do
  echo "== TB_SCAN_RPL2_OCC=$occ"
  TB_SCAN_RPL2_OCC=$occ timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs 3 2>>gpurun_out/bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); v=d['configs']['cfg3']; print(v.get('paths_per_s'),v.get('ms_per_step'),v.get('split'))"
done
echo "== pytest cfg3 + parity with 20"; TB_SCAN_RPL2_OCC=20 timeout 900 python -m pytest tests/test_cfg3_torque.py tests/test_gpu_parity.py tests/test_gpu_scale.py -q -m gpu -x 2>&1 | tail -2
