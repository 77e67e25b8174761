"""cfg 3 (device inverse-dynamics model) on 16384 paths: the kernels of one BatchTOPPRA solve (for ncu launch lists)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import toppra_b200 as ta
from problems import make_batch_fast
B, G, dof = int(os.environ.get("B", 16384)), 500, 6
ss, way, vlim, alim = make_batch_fast(B, seed=2000, dof=dof)
tl = 40 + np.random.RandomState(7).rand(B, dof) * 10
cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim),
        ta.constraint.SecondOrderConstraint.joint_torque_constraint(None, np.stack((-tl, tl), axis=-1), np.zeros(dof),
                                                                    device_model=("coupled_cosine", [2.0, 0.3, 0.1, 4.9]))]
path = ta.BatchSplineInterpolator(ss, way)
for _ in range(2):
    res = ta.BatchTOPPRA(cons, path, np.linspace(0, 1, G)).compute_parameterization(0.0, 0.0)
torch.cuda.synchronize()
print("ok", int((res.status != 0).sum()))
