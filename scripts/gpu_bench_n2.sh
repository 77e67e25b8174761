#!/usr/bin/env bash
# 2-GPU call: NCCL parity tests + bench under torchrun
set -u
mkdir -p gpurun_out
echo "== pytest multigpu"; timeout 900 python -m pytest tests/test_multigpu_nccl.py -q -m gpu -x 2>&1 | tail -5
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench2.err | tee gpurun_out/bench_n2.json | python -c "
import sys,json; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('value',d['value'],'ms',d['ms_per_step'],d['kernels_ms']); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['host_sync_every_step']); print({k:(v.get('paths_per_s'),v.get('ms_per_step'),v.get('kernels_ms_ranks_min_max'),v.get('gather_exposed_ms')) for k,v in d['configs'].items() if isinstance(v,dict)})"
tail -5 gpurun_out/bench2.err
