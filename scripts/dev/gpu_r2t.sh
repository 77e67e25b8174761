#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --configs 3,4 2>gpurun_out/bench.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],d['kernels_ms']); print({k:(v.get('paths_per_s'),v.get('ms_per_step'),v.get('split')) for k,v in d['configs'].items() if isinstance(v,dict)})"
tail -3 gpurun_out/bench.err
