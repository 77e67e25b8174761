#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench.err > gpurun_out/bench_ref.json; cut -c1-160 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>>gpurun_out/bench.err > gpurun_out/bench.json; python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],d['kernels_ms']); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['host_sync_every_step']['value']); print({k:(v.get('paths_per_s'),v.get('ms_per_step')) for k,v in d['configs'].items() if isinstance(v,dict)}); print(d['cpu_baseline'])"
