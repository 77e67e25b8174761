#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-120 gpurun_out/bench.json
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err > gpurun_out/bench_ref.json; cut -c1-160 gpurun_out/bench_ref.json
