#!/usr/bin/env bash
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC"
link() { (cd toppra_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libtoppra_b200.so tb_api.o tb_spline.o tb_coeff.o tb_scan.o tb_robust.o tb_param.o); }
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $2 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$1 B=$2 value %.0f K1 %.4f K2 %.3f ms' % (d['value'], d['kernels_ms']['K1_coeff'], d['kernels_ms']['K2_scan']))"; }
for cfg in "64 16" "64 32" "96 32" "128 16" "192 32" "96 16" "128 24"; do
  set -- $cfg
  (cd toppra_b200/csrc && $NV -DTB_COEFF_THREADS=$1 -DTB_COEFF_CH=$2 -c tb_coeff.cu -o tb_coeff.o 2>/dev/null); link
  run "K1 threads=$1 CH=$2" 4096
done
