"""GPU (-m gpu), needs >= 2 GPUs (skipped otherwise): the N > 1 path on hardware.  `toppra_b200.distributed.solve_sharded`
over NCCL — every rank solves its contiguous shard on its own GPU, the results are all-gathered — must return, on every
rank, exactly the arrays of the single-GPU solve of the whole batch (BASELINE cfg 5 is this with 2^20 paths)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, G, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch
    import torch.distributed as dist
    from problems import make_batch_fast
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import toppra_b200 as ta
    from toppra_b200.distributed import solve_sharded
    ss, way, vlim, alim = make_batch_fast(B, seed=4242)
    vlim[: B // 8] *= 0.03          # some velocity-active paths
    s0 = np.where(np.arange(B) % 11 == 0, 30.0, 0.0)   # some inadmissible starts -> FailUncontrollable
    grid = np.linspace(0, 1, G)
    full = solve_sharded(ss, way, grid, vlim, alim, s0, 0.0)
    if rank == 0:
        one = ta.batch.solve_batch(ss, way, grid, vlim, alim, s0, 0.0)
        ok = all(torch.equal(full[k], getattr(one, k)) or
                 bool(((full[k] == getattr(one, k)) | (full[k].isnan() & getattr(one, k).isnan())).all())
                 for k in ("K", "sd", "sdd"))
        ok = ok and torch.equal(full["status"], one.status) and int((one.status == 3).sum()) >= B // 11
        ret[0] = bool(ok)
    else:
        ret[rank] = bool(full["K"].shape == (B, G, 2) and full["status"].shape == (B,))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [1001, 4096])
def test_solve_sharded_nccl_equals_single_gpu(B):
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 4)
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), B, 120, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
