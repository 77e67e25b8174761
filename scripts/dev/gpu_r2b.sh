#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests')
import toppra_b200 as ta
from problems import make_batch_fast
B, G = 4096, 200
ss, way, vlim, alim = make_batch_fast(B, seed=1234)
grid = np.linspace(0, 1, G)
path = ta.BatchSplineInterpolator(ss, way)
inst = ta.BatchTOPPRA([ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)], path, grid)
res = inst.compute_parameterization(0.0, 0.0, counters=True)
c = res.counters.cpu().numpy()
print("counters mean per path: lp2d %.1f lp1d %.1f resolves %.1f retries %.3f" % tuple(c.mean(0)))
r = c[:, 2].reshape(-1, 4)
print("re-solves per path: min %d max %d; per-warp max-sum/mean ratio %.3f" % (c[:, 2].min(), c[:, 2].max(), r.max(1).mean() / r.mean()))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 6 -c 1 \
   -o gpurun_out/prof_r2b -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r2b.log 2>&1
tail -3 gpurun_out/ncu_full_r2b.log
ls -la gpurun_out | tail -4
