"""Multi-GPU plumbing: paths are independent, so a batch is sharded contiguously over the ranks with no
data-path collective; NCCL (or gloo in the CPU tests) is used once, to gather the results
(SURVEY.md §8e).  One process per GPU (torchrun), `torch.distributed` already initialised by the caller."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous shard [lo, hi) of B paths for `rank` of `world` (the first B % world ranks get one more)."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_results(local, B_total, group=None):
    """All-gather per-path result tensors (dict name -> tensor whose dim 0 is the local shard) into full-batch
    tensors on every rank.  Shards may differ in size by one path (padded for the collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(B_total, world)
    nmax = max(sizes)
    out = {}
    for name, t in local.items():
        pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if t.shape[0]:
            pad[: t.shape[0]] = t
        full = torch.empty((world * nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, pad, group=group)
        parts = [full[r * nmax: r * nmax + sizes[r]] for r in range(world)]
        out[name] = torch.cat(parts, dim=0)
    return out


def solve_sharded(ss_waypoints, waypoints, gridpoints, vlim, alim, sd_start=0.0, sd_end=0.0, gather=True,
                  group=None, device=None):
    """Every rank passes the FULL batch description (numpy); each solves its shard on its GPU and, with
    gather=True, every rank returns the full-batch K, sd, sdd, status tensors."""
    import torch.distributed as dist
    from .batch import solve_batch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = waypoints.shape[0]
    lo, hi = shard_range(B, rank, world)

    def sh(x, nd):  # per-path arrays are sharded, shared ones passed through
        x = np.asarray(x)
        return x[lo:hi] if x.ndim == nd else x

    if hi > lo:
        res = solve_batch(sh(ss_waypoints, 2), waypoints[lo:hi], sh(gridpoints, 2), sh(vlim, 3), sh(alim, 3),
                          sh(sd_start, 1) if np.ndim(sd_start) else sd_start,
                          sh(sd_end, 1) if np.ndim(sd_end) else sd_end, device=device)
        local = dict(K=res.K, sd=res.sd, sdd=res.sdd, status=res.status)
    else:
        # more ranks than paths: this rank has nothing to solve but still takes part in the gather
        from . import engine
        torch = engine.torch_mod()
        dev = engine.default_device(device)
        G = np.asarray(gridpoints).shape[-1]
        local = dict(K=torch.empty((0, G, 2), dtype=torch.float64, device=dev),
                     sd=torch.empty((0, G), dtype=torch.float64, device=dev),
                     sdd=torch.empty((0, max(G - 1, 0)), dtype=torch.float64, device=dev),
                     status=torch.empty((0,), dtype=torch.int32, device=dev))
    return gather_results(local, B, group) if gather else local
