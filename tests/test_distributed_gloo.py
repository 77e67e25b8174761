"""CPU: the N > 1 path (contiguous sharding + result gather) with world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from toppra_b200.distributed import gather_results, shard_range, shard_sizes


def test_shard_range_covers_batch():
    for B in (1, 2, 7, 4096, 4097, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(B, world)) - min(shard_sizes(B, world)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, G, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    # stand-in for the per-rank kernel outputs: values that encode the global path index
    idx = torch.arange(lo, hi, dtype=torch.float64)
    local = dict(sd=idx[:, None] + torch.arange(G, dtype=torch.float64)[None, :] / 1000.0,
                 K=torch.stack((idx, idx + 0.5), dim=1)[:, None, :].expand(hi - lo, G, 2).contiguous(),
                 status=(torch.arange(lo, hi) % 4).to(torch.int32))
    full = gather_results(local, B)
    ok = (full["sd"].shape == (B, G) and full["K"].shape == (B, G, 2) and full["status"].shape == (B,)
          and torch.equal(full["sd"][:, 0], torch.arange(B, dtype=torch.float64))
          and torch.equal(full["K"][:, 3, 1], torch.arange(B, dtype=torch.float64) + 0.5)
          and torch.equal(full["status"], (torch.arange(B) % 4).to(torch.int32)))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 9, 1])
def test_gather_world2_gloo(B):
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), B, 5, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _solve_worker(rank, world, port, ret):
    """solve_sharded end to end on the oracle-backed CPU engine double (tests/cpu_engine.py): every rank solves its
    contiguous shard, gloo gathers; the result must be the reference's golden batch on every rank."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_engine
    cpu_engine.install_plain()
    from toppra_b200.distributed import solve_sharded
    g = np.load(os.path.join(here, "golden", "cfg2_seeds1000.npz"))
    B = 7  # odd: shards of 4 and 3 paths
    full = solve_sharded(g["ss"], g["way"][:B], g["grid"], g["vlim"][:B], g["alim"][:B], 0.0, 0.0)
    ok = all(np.array_equal(full[k].numpy(), g[k if k != "sdd" else "sdd"][:B]) for k in ("K", "sd", "sdd"))
    ok = ok and np.array_equal(full["status"].numpy(), g["status"][:B]) and full["K"].shape == (B, len(g["grid"]), 2)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_solve_sharded_world2_gloo_on_cpu_double():
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_solve_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _solve_worker_tiny(rank, world, port, ret):
    """More ranks than paths: the rank with an empty shard must not hang the gather."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_engine
    cpu_engine.install_plain()
    from toppra_b200.distributed import solve_sharded
    g = np.load(os.path.join(here, "golden", "cfg2_seeds1000.npz"))
    full = solve_sharded(g["ss"], g["way"][:1], g["grid"], g["vlim"][:1], g["alim"][:1], 0.0, 0.0, device="cpu")
    ok = np.array_equal(full["K"].numpy(), g["K"][:1]) and np.array_equal(full["sd"].numpy(), g["sd"][:1])
    ret[rank] = bool(ok and full["status"].shape == (1,))
    dist.destroy_process_group()


def test_solve_sharded_more_ranks_than_paths():
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_solve_worker_tiny, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _sharded_solver_worker(rank, world, port, ret):
    """distributed.ShardedSolver (BASELINE cfg 5's entry point: equal shards, chunked solve, per-chunk all-gather into global
    path order) on the CPU engine double over gloo: every rank must hold the reference's golden batch in path order."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_engine
    cpu_engine.install_plain()
    from toppra_b200.distributed import ShardedSolver, shard_range
    g = np.load(os.path.join(here, "golden", "cfg2_seeds1000.npz"))
    B, G = 8, len(g["grid"])
    lo, hi = shard_range(B, rank, world)
    solver = ShardedSolver(B, G, "cpu", nchunks=3, gather=True)          # 4 paths per rank in chunks of 1, 1, 2
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))              # noqa: E731
    full = solver.solve(t(g["ss"]), t(g["way"][lo:hi]), t(g["grid"]), t(g["vlim"][lo:hi]), t(g["alim"][lo:hi]))
    ok = all(np.array_equal(full[k].numpy(), g[k][:B]) for k in ("K", "sd", "sdd"))
    ok = ok and np.array_equal(full["status"].numpy(), g["status"][:B]) and solver.nchunks == 3
    local = ShardedSolver(B, G, "cpu", nchunks=2, gather=False).solve(t(g["ss"]), t(g["way"][lo:hi]), t(g["grid"]),
                                                                        t(g["vlim"][lo:hi]), t(g["alim"][lo:hi]))
    ok = ok and np.array_equal(local["sd"].numpy(), g["sd"][lo:hi])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_solver_world2_gloo_on_cpu_double():
    world = 2
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_sharded_solver_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
