#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 5 --warmup 3 --configs 5 2>gpurun_out/bench2.err | tee gpurun_out/bench_n2.json | python -c "
import sys,json; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step']); v=d['configs']['cfg5']; print(v.get('paths_per_s'),v.get('ms_per_step'),v.get('kernels_ms_ranks_min_max'),v.get('gather_exposed_ms'), v.get('status'))"
tail -2 gpurun_out/bench2.err
