#!/usr/bin/env bash
# compute-sanitizer over a small end-to-end solve (K0, K1, K2, K2r, K3, LP shims): memcheck + racecheck + synccheck.
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import toppra_b200 as ta
from problems import make_batch
B, G = 24, 40
ss, way, vlim, alim = make_batch(B, 1000)
grid = np.linspace(0, 1, G)
path = ta.BatchSplineInterpolator(ss, way)
vel, acc = ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)
inst = ta.BatchTOPPRA([vel, acc], path, grid)
res = inst.compute_parameterization(0.0, 0.0)
host = inst.solve_to_host(0.0, 0.0)
X = inst.compute_feasible_sets()
rob = ta.BatchTOPPRA([vel, ta.constraint.RobustLinearConstraint(acc, [1e-3, 5e-2, 9e-3], 1)], path, grid).compute_parameterization(0, 0)
bp = ta.BatchParametrizeConstAccel(path, grid, res.sd)
q = bp(np.linspace(0, 1.0, 16), 1)
r = ta.engine.lp2d_batch(np.random.randn(8, 3), np.random.randn(8, 40), np.random.randn(8, 40), -np.random.rand(8, 40),
                         np.tile([-1.0, -1.0], (8, 1)), np.tile([1.0, 1.0], (8, 1)))
print("ok", int((res.status != 0).sum()), int((rob.status != 0).sum()), float(bp.durations.min()))
PY
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|error|Error" | head -8
done 2>&1 | tee gpurun_out/sanitizer_r01.txt
