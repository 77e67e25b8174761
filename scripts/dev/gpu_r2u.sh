#!/usr/bin/env bash
# N=2: exposed gather of cfg 5 with different residency caps of the fused scan (32 = no cap)
set -u
mkdir -p gpurun_out
for cap in 32 24 16; do
  echo "== TB_SHARD_RESIDENT=$cap"
  TB_SHARD_RESIDENT=$cap timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$((cap % 10)) bench.py --gpus 2 --steps 5 --warmup 3 --configs 5 2>>gpurun_out/bench2.err | python -c "
import sys,json; d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); v=d['configs']['cfg5']; print(v.get('paths_per_s'),v.get('ms_per_step'),v.get('kernels_ms_ranks_min_max'),v.get('gather_exposed_ms'))"
done
tail -3 gpurun_out/bench2.err
