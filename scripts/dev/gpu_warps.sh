#!/usr/bin/env bash
# A/B: resident warps per SM (register cap) in the scan kernel, 1 warp per CTA
for w in 24 28 32; do
  (cd toppra_b200/csrc && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -DTB_SCAN_WARPS_PER_SM=$w -c tb_scan.cu -o tb_scan.o 2>/dev/null && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libtoppra_b200.so tb_api.o tb_spline.o tb_coeff.o tb_scan.o tb_robust.o tb_param.o)
  for b in 4096 65536; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $b | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('warps/SM=$w B=$b value %.0f  K2 %.3f ms' % (d['value'], d['kernels_ms']['K2_scan']))"
  done
done
