#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_frows_batch.py -q -m gpu -x 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],d['kernels_ms']); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['host_sync_every_step']); print({k:(v.get('paths_per_s'),v.get('ms_per_step')) for k,v in d['configs'].items() if isinstance(v,dict)}); print(d['cpu_baseline'])"
tail -3 gpurun_out/bench.err
