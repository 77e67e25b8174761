"""cfg 4 (robust) on 4096 paths: one BatchTOPPRA solve (for ncu captures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import toppra_b200 as ta
from problems import make_batch_fast
B, G, dof = int(os.environ.get("B", 4096)), 200, 7
ss, way, vlim, alim = make_batch_fast(B, seed=3000, dof=dof)
cons = [ta.constraint.JointVelocityConstraint(vlim),
        ta.constraint.RobustLinearConstraint(ta.constraint.JointAccelerationConstraint(alim), [1e-3, 5e-2, 9e-3], 1)]
path = ta.BatchSplineInterpolator(ss, way)
for _ in range(3):
    res = ta.BatchTOPPRA(cons, path, np.linspace(0, 1, G)).compute_parameterization(0.0, 0.0)
torch.cuda.synchronize()
print("ok", int((res.status != 0).sum()))
