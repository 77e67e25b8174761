"""Quick GPU sanity run: parity vs the C oracle on a small batch + a first timing.  (dev helper)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toppra_b200 as ta
from toppra_b200 import engine
from oracle import oracle as orc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from problems import make_batch, make_batch_fast

def run(B, G, vel_active, sd0=0.0, sd1=0.0):
    ss, way, vlim, alim = make_batch(B, 1000, vel_active=vel_active)
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    c_gpu = path.d_ppoly.cpu().numpy()
    c_orc = np.stack([orc.cubic_spline_fit(ss, way[b]) for b in range(B)])
    print("fit bit-exact:", np.array_equal(c_gpu, c_orc), np.abs(c_gpu - c_orc).max())
    inst = ta.BatchTOPPRA([ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)], path, grid)
    res = inst.compute_parameterization(sd0, sd1, counters=True)
    h = res.to_host()
    cnt = res.counters.cpu().numpy()
    o = orc.solve_velacc_batch(c_orc, np.tile(ss, (B, 1)), grid, vlim, alim, True, np.full(B, sd0), np.full(B, sd1), nthreads=8)
    R = inst.R
    rec = inst.records.cpu().numpy()
    o1 = orc.solve_velacc(c_orc[0], ss, grid, vlim[0], alim[0], True, sd0, sd1, want_rows=True)
    print("rows bit-exact:", np.array_equal(rec[0, :, :3 * R].reshape(G, 3, R), o1["rows"]), "xbound:", np.array_equal(rec[0, :, 3 * R:3 * R + 2], o1["xbound"]))
    for k, ok in (("K", "K"), ("sd", "sd"), ("sdd", "u")):
        a, b = h[k], o[ok]
        eq = np.array_equal(a, b, equal_nan=True)
        print(k, "bit-exact:", eq, "maxdiff", np.nanmax(np.abs(a - b)))
    print("status equal:", np.array_equal(h["status"], o["status"]), np.bincount(h["status"]), "counters mean", cnt.mean(0), "oracle0", o1["counters"])

run(64, 200, False)
run(64, 200, True)
run(32, 100, True, 0.1, 0.1)
run(8, 50, False, 5.0, 0.0)

# timing, cfg 2
B, G = 4096, 200
ss, way, vlim, alim = make_batch_fast(B)
grid = np.linspace(0, 1, G)
d_way = torch.as_tensor(way).cuda(); d_ss = torch.as_tensor(ss).cuda()
cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)]
def step():
    path = ta.BatchSplineInterpolator(d_ss, d_way)
    inst = ta.BatchTOPPRA(cons, path, grid)
    return inst.compute_parameterization(0.0, 0.0)
for _ in range(3): r = step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10): r = step()
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print("B=%d G=%d: %.3f ms/step -> %.0f paths/s; status hist %s" % (B, G, ms, B / ms * 1e3, np.bincount(r.status.cpu().numpy())))
# per-kernel
path = ta.BatchSplineInterpolator(d_ss, d_way); inst = ta.BatchTOPPRA(cons, path, grid)
for name, fn in (("K0 fit", lambda: ta.BatchSplineInterpolator(d_ss, d_way)), ("K1 coeff", inst.setup), ("K2 scan", lambda: inst.compute_parameterization(0.0, 0.0))):
    fn(); torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10): fn()
    ev[1].record(); torch.cuda.synchronize()
    print("  %s: %.3f ms" % (name, ev[0].elapsed_time(ev[1]) / 10))
