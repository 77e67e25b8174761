#!/usr/bin/env bash
# N-GPU scaling check: bench under torchrun for N = $1
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | python -c "
import sys,json; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('value',d['value'],'ms',d['ms_per_step'],d['kernels_ms']); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['host_sync_every_step']); print({k:(v.get('paths_per_s'),v.get('ms_per_step'),v.get('kernels_ms_ranks_min_max'),v.get('gather_exposed_ms')) for k,v in d['configs'].items() if isinstance(v,dict)}); print(d['clocks'])"
tail -3 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>>gpurun_out/bench_n$N.err | tee gpurun_out/bench_ref_n$N.json | cut -c1-300
