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


class ShardedSolver(object):
    """BASELINE config 5: a batch of `B_total` vel+acc problems sharded contiguously over the ranks; every rank owns the
    inputs of ITS shard as device tensors.  The shard is solved in `nchunks` chunks; the all-gather of chunk c (K, sd,
    sdd, status; NCCL) is issued on a side stream and overlaps the scan of chunk c+1 ("overlap by
    gathering per chunk on a side stream"), so the collective leaves the critical path except for the last chunk.

    The gathered tensors are preallocated once and laid out in GLOBAL path order ([B_total, ...] on every rank, or only
    the local shard with gather=False).  Shards must be equal (B_total % world == 0): the bench's and the tests' case;
    `solve_sharded` covers ragged shards.  Without an initialised process group this is a plain chunked solve."""

    def __init__(self, B_total, G, device, nchunks=4, gather=True, group=None, validate=False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        if B_total % self.world:
            raise ValueError("ShardedSolver needs equal shards: B_total %d %% world %d != 0" % (B_total, self.world))
        self.B, self.G, self.dev = B_total, G, device
        self.shard = B_total // self.world
        self.nchunks = max(1, min(int(nchunks), self.shard))
        self.bounds = [(self.shard * c) // self.nchunks for c in range(self.nchunks + 1)]
        self.gather = bool(gather) and self.world > 1
        n_out = B_total if self.gather else self.shard
        f64 = dict(dtype=torch.float64, device=device)
        self.out = dict(K=torch.empty((n_out, G, 2), **f64), sd=torch.empty((n_out, G), **f64),
                        sdd=torch.empty((n_out, G - 1), **f64),
                        status=torch.empty((n_out,), dtype=torch.int32, device=device))
        # the overlap needs CUDA streams; on a CPU device (the gloo tests of the host logic) the gathers run inline
        self.cuda = torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device) if (self.gather and self.cuda) else None
        cmax = max(b - a for a, b in zip(self.bounds[:-1], self.bounds[1:]))
        # NCCL writes [world][chunk] blocks; two scratch sets so that chunk c+1's gather never waits for chunk c's unpack
        self.scratch = [dict((k, torch.empty((self.world * cmax,) + tuple(v.shape[1:]), dtype=v.dtype, device=device))
                             for k, v in self.out.items()) for _ in range(2)] if self.gather else None
        self.kernel_events = []
        self.validate = bool(validate)   # input checks on CUDA tensors cost a device synchronisation per chunk

    def solve(self, d_ss, d_way, d_grid, d_vlim, d_alim, sd_start=0.0, sd_end=0.0, record_events=False):
        """d_way [shard, n, dof], d_vlim/d_alim [shard, dof, 2] (or shared [dof, 2]), d_ss/d_grid shared 1-D: this rank's
        shard.  Returns the dict of (gathered) result tensors; they are complete on the current stream."""
        torch, dist = self.torch, self.dist
        from .batch import BatchTOPPRA
        from .constraint import JointAccelerationConstraint, JointVelocityConstraint
        from .interpolator import BatchSplineInterpolator
        main = torch.cuda.current_stream(self.dev) if self.cuda else None
        record_events = record_events and self.cuda
        for c in range(self.nchunks):
            lo, hi = self.bounds[c], self.bounds[c + 1]
            n = hi - lo
            if record_events:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(main)
            path = BatchSplineInterpolator(d_ss, d_way[lo:hi], device=self.dev, validate=self.validate)
            pv = JointVelocityConstraint.from_device(d_vlim if d_vlim.dim() == 2 else d_vlim[lo:hi])
            pa = JointAccelerationConstraint.from_device(d_alim if d_alim.dim() == 2 else d_alim[lo:hi])
            res = BatchTOPPRA([pv, pa], path, d_grid, validate=self.validate).compute_parameterization(sd_start, sd_end)
            if record_events:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(main)
                self.kernel_events.append((e0, e1))
            local = dict(K=res.K, sd=res.sd, sdd=res.sdd, status=res.status)
            if not self.gather:
                for k, t in local.items():
                    self.out[k][lo:hi].copy_(t, non_blocking=True)
                continue
            scr = self.scratch[c & 1]
            if self.cuda:
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(done)
                    for t in local.values():
                        t.record_stream(self.side)
                    self._gather_chunk(local, scr, lo, hi)
            else:
                self._gather_chunk(local, scr, lo, hi)
        if self.gather and self.cuda:
            main.wait_stream(self.side)
        return self.out

    def _gather_chunk(self, local, scr, lo, hi):
        """All-gather one chunk of every result tensor and unpack the [world][n] blocks into global path order (rank r's
        chunk sits at r * shard + lo)."""
        n = hi - lo
        for k, t in local.items():
            buf = scr[k][: self.world * n]
            self.dist.all_gather_into_tensor(buf, t.contiguous(), group=self.group)
            view = self.out[k].view((self.world, self.shard) + tuple(self.out[k].shape[1:]))
            view[:, lo:hi].copy_(buf.view((self.world, n) + tuple(t.shape[1:])), non_blocking=True)
