#!/usr/bin/env bash
# parity tests, then bench with the default build and with other register caps of the scan kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$1 value %.0f e2e %.0f fast %.0f' % (d['value'], d['e2e']['value'], d['opt_in_fast_lower_bound']['value']), d['kernels_ms'])
"; }
run "default(32 warps/SM)"
for w in ${WARPS:-28}; do
  (cd toppra_b200/csrc && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -DTB_SCAN_WARPS_PER_SM=$w -c tb_scan.cu -o tb_scan.o 2>/dev/null && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libtoppra_b200.so tb_api.o tb_spline.o tb_coeff.o tb_scan.o tb_robust.o tb_param.o)
  run "warps/SM=$w"
done
