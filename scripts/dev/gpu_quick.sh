#!/usr/bin/env bash
# quick A/B: bench.py (no CPU arm) with optional env; usage: gpu_quick.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.3fM e2e %.3fM (sync %.3fM) ms/step %.4f kernels %s fast %.3fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['host_sync_every_step']['value']/1e6, d['ms_per_step'], d['kernels_ms'], d['opt_in_fast_lower_bound']['value']/1e6))"
tail -2 gpurun_out/bench_$TAG.err
