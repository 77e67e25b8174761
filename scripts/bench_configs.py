#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (3: torque, 4: robust, 5: 1M paths) on this repo's GPU path.
Not the driver's bench (bench.py measures configs[1]); prints one JSON line per config.
  python scripts/bench_configs.py [cfg3] [cfg4] [cfg5] [--scale 0.25]      (torchrun for cfg5 on several GPUs)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import toppra_b200 as ta  # noqa: E402
from problems import inv_dyn_torch, make_batch_fast  # noqa: E402


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    names = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg3", "cfg4", "cfg5"]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    for name in names:
        if name == "cfg3":   # batch 65536 6-DOF paths, 500 gridpoints, vel+acc + SecondOrder (torque)
            B, G, dof = int(65536 * scale) // world, 500, 6
            ss, way, vlim, alim = make_batch_fast(B, seed=2000 + rank, dof=dof)
            rng = np.random.RandomState(7 + rank)
            tl = 40 + rng.rand(B, dof) * 10
            taulim = np.stack((-tl, tl), axis=-1)
            cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim),
                    ta.constraint.SecondOrderConstraint.joint_torque_constraint(inv_dyn_torch, taulim, np.zeros(dof),
                                                                                batched=True)]
        elif name == "cfg4":  # robust TOPP-RA, batch 4096, 7-DOF, 200 gridpoints
            B, G, dof = int(4096 * scale), 200, 7
            ss, way, vlim, alim = make_batch_fast(B, seed=3000 + rank, dof=dof)
            cons = [ta.constraint.JointVelocityConstraint(vlim),
                    ta.constraint.RobustLinearConstraint(ta.constraint.JointAccelerationConstraint(alim),
                                                         [1e-3, 5e-2, 9e-3], 1)]
        else:                 # cfg5: 1M-path batch sharded across the GPUs, 7-DOF, 200 gridpoints
            B, G, dof = int((1 << 20) * scale) // world, 200, 7
            ss, way, vlim, alim = make_batch_fast(B, seed=1000 + rank, dof=dof)
            cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)]
        grid = np.linspace(0, 1, G)
        d_way = torch.as_tensor(way).cuda()
        d_ss = torch.as_tensor(ss).cuda()
        d_grid = torch.as_tensor(grid).cuda()
        for c in cons:   # limits resident on the device, like the inputs
            if hasattr(c, "device_limits"):
                c.device_limits(d_way.device)

        def step():
            path = ta.BatchSplineInterpolator(d_ss, d_way)
            inst = ta.BatchTOPPRA(cons, path, d_grid)
            res = inst.compute_parameterization(0.0, 0.0)
            return inst, res

        ms, (inst, res) = timed(step, steps=3 if B * G > 5e7 else 5)
        split = {}
        if inst.chunk_size() >= B:   # single chunk: time K0 / record build (K1 + user inv_dyn) / scan separately
            split["fit_ms"], path0 = timed(lambda: ta.BatchSplineInterpolator(d_ss, d_way), steps=3)
            inst0 = ta.BatchTOPPRA(cons, path0, d_grid)
            split["records_ms"], _ = timed(inst0.setup, steps=3)
            split["scan_ms"], _ = timed(lambda: inst0.compute_parameterization(0.0, 0.0), steps=3)
            del inst0, path0
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        hist = torch.bincount(res.status, minlength=5).cpu().tolist()
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "paths_per_gpu": B, "gridpoints": G, "dof": dof, "rows": inst.R,
                              "chunk_paths": inst.chunk_size(), "ms_per_step": float(t.item()),
                              "paths_per_s": B * world / float(t.item()) * 1e3, "status_hist_rank0": hist, "split": split}),
                  flush=True)
        del inst, res, d_way
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
