#!/usr/bin/env bash
set -u
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --configs none 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['host_sync_every_step']['ms_per_step'], d['records_path']['kernels_ms'])"
done
