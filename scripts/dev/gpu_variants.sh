#!/usr/bin/env bash
# Timing of compile-time variants of the scan kernel (built on the GPU box).  usage: gpu_variants.sh "<flags>" "<flags>" ...
set -u
mkdir -p gpurun_out; : > gpurun_out/variants.txt
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC"
for flags in "$@"; do
  (cd toppra_b200/csrc && $NV $flags -c tb_scan.cu -o tb_scan.o 2>/dev/null && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libtoppra_b200.so tb_api.o tb_spline.o tb_coeff.o tb_scan.o tb_robust.o tb_param.o)
  timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('[$flags] value %.0f e2e %.0f fast %.0f K2 %.4f ms' % (d['value'], d['e2e']['value'], d['opt_in_fast_lower_bound']['value'], d['kernels_ms']['K2_scan']))
        open('gpurun_out/variants.txt','a').write('$flags|%.5f\n' % d['kernels_ms']['K2_scan'])
"
done
# rebuild the fastest variant and run the parity tests with it
best=$(sort -t'|' -k2 -n gpurun_out/variants.txt | head -1 | cut -d'|' -f1)
echo "== best: [$best]"
(cd toppra_b200/csrc && $NV $best -c tb_scan.cu -o tb_scan.o 2>/dev/null && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libtoppra_b200.so tb_api.o tb_spline.o tb_coeff.o tb_scan.o tb_robust.o tb_param.o)
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
