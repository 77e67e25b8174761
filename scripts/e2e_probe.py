"""Where does the e2e step spend host time?  (development probe, not a bench)"""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import toppra_b200 as ta
from problems import make_batch_fast
B, G = 4096, 200
ss, way, vlim, alim = make_batch_fast(B, seed=1234)
grid = np.linspace(0, 1, G)
dev = torch.device("cuda", 0)
h_way = torch.as_tensor(way).pin_memory(); h_vlim = torch.as_tensor(vlim).pin_memory(); h_alim = torch.as_tensor(alim).pin_memory()
h_ss = torch.as_tensor(ss).pin_memory(); h_grid = torch.as_tensor(grid).pin_memory()
h_out = None
T = {}
def tick(name, t0):
    torch.cuda.synchronize() if SYNC_EACH else None
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(sync):
    global h_out
    t0 = time.perf_counter(); path = ta.BatchSplineInterpolator(h_ss, h_way, device=dev); tick("path", t0)
    t0 = time.perf_counter()
    pc_vel = ta.constraint.JointVelocityConstraint(vlim); pc_acc = ta.constraint.JointAccelerationConstraint(alim)
    pc_vel._d_cache[str(dev)] = h_vlim.to(dev, non_blocking=True); pc_acc._d_cache[str(dev)] = h_alim.to(dev, non_blocking=True)
    tick("cons", t0)
    t0 = time.perf_counter(); inst = ta.BatchTOPPRA([pc_vel, pc_acc], path, h_grid); tick("inst", t0)
    t0 = time.perf_counter(); h_out = inst.solve_to_host(0.0, 0.0, pinned=h_out, sync=sync); tick("solve", t0)
for SYNC_EACH in (False, True):
    for sync in (False, True):
        for _ in range(5): step(sync)
        torch.cuda.synchronize(); T.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(50): step(sync)
        e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
        print("probe-sync-each-call=%s sync=%s: wall %.3f ms/step, gpu %.3f ms/step, host parts (ms/step): %s" % (
            SYNC_EACH, sync, wall / 50 * 1e3, e0.elapsed_time(e1) / 50, {k: round(v / 50 * 1e3, 3) for k, v in T.items()}))

# ---- bench-like loop variants: per-step events, L2 flush between steps, NVML sampler thread ----
SYNC_EACH = False
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
import threading
def run(label, use_flush, use_events, sampler):
    stop = {"f": False}
    th = None
    if sampler:
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
        def poll():
            while not stop["f"]:
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h); time.sleep(0.02)
        th = threading.Thread(target=poll, daemon=True); th.start()
    for _ in range(5):
        step(False)
        if use_flush: flush.zero_()
    torch.cuda.synchronize()
    pairs = []
    t0 = time.perf_counter()
    E0, E1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    E0.record()
    for _ in range(20):
        if use_flush: flush.zero_()
        if use_events:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.record()
        step(False)
        if use_events:
            e.record(); pairs.append((s, e))
    E1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    stop["f"] = True
    if th: th.join()
    per = sum(s.elapsed_time(e) for s, e in pairs) / 20 if pairs else float("nan")
    print("%-40s wall %.3f ms/step  gpu-total %.3f  sum-of-step-events %.3f" % (label, wall / 20 * 1e3, E0.elapsed_time(E1) / 20, per))
run("plain", False, False, False)
run("events", False, True, False)
run("flush", True, False, False)
run("flush+events", True, True, False)
run("flush+events+sampler", True, True, True)
