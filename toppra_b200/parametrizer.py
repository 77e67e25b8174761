"""Output trajectories from (gridpoints, sd) — same surface as the reference `toppra/parametrizer.py`
(`ParametrizeConstAccel` :23-158, `ParametrizeSpline` :161-196).

SURVEY.md §8 (f1), "next" row: the O(G) time-stamp recurrence below is host numpy for a single path; the spline
re-fit and every path evaluation run on the GPU through SplineInterpolator."""
import logging

import numpy as np

from . import engine
from .exceptions import ToppraError
from .interpolator import AbstractGeometricPath, SplineInterpolator

logger = logging.getLogger(__name__)


class BatchParametrizeConstAccel(object):
    """Output trajectories of B paths under the piecewise-constant path-acceleration assumption, on the GPU
    (csrc/tb_param.cu).  path: BatchSplineInterpolator; gridpoints: CUDA tensor [G] or [B, G]; velocities: sd [B, G]."""

    def __init__(self, path, gridpoints, velocities):
        self._path = path
        self.d_grid = engine.as_device(gridpoints, path.device)
        self.d_sd = engine.as_device(velocities, path.device)
        self.t_grid, self.us = engine.time_grid(self.d_sd, self.d_grid)

    @property
    def durations(self):
        """CUDA tensor [B]: duration of every trajectory."""
        return self.t_grid[:, -1]

    def __call__(self, ts, order=0):
        """ts: [M] (shared) or [B, M] times -> CUDA tensor [B, M, dof] of q (order 0), qd (1) or qdd (2)."""
        if order not in (0, 1, 2):
            raise ToppraError(f"Order {order} is not supported.")
        ts = engine.as_device(ts, self._path.device)
        return engine.constaccel_eval(self._path.d_ppoly, self._path.d_ss, self.d_grid, self.d_sd, self.t_grid,
                                      self.us, ts, order)


class ParametrizeConstAccel(AbstractGeometricPath):
    """Output trajectory under the piecewise-constant path-acceleration assumption (reference
    parametrizer.py:23-158): on [s_i, s_{i+1}]  u_i = (x_{i+1} - x_i) / (2 ds),
    t_{i+1} = t_i + 2 ds / (sd_i + sd_{i+1}).  Time stamps and evaluation run on the GPU (csrc/tb_param.cu)."""

    def __init__(self, path, gridpoints, velocities):
        self._path = path
        self._ss = np.array(gridpoints, dtype=np.float64)
        self._velocities = np.array(velocities, dtype=np.float64)
        self._xs = self._velocities ** 2
        assert self._ss.shape[0] == self._velocities.shape[0]
        assert len(self._ss.shape) == 1
        assert np.all(self._velocities >= 0)
        self._batch = BatchParametrizeConstAccel(path.as_batch(), self._ss, self._velocities[None])
        self._ts = self._batch.t_grid[0].cpu().numpy()
        self._us = self._batch.us[0].cpu().numpy()

    @property
    def dof(self):
        return self._path.dof

    @property
    def path_interval(self):
        return np.array([self._ts[0], self._ts[-1]])

    @property
    def duration(self):
        return self.path_interval[1] - self.path_interval[0]

    def __call__(self, ts, order=0):
        scalar = isinstance(ts, (int, float))
        ts = np.array([ts], dtype=float) if scalar else np.asarray(ts, dtype=float)
        if order not in (0, 1, 2):
            raise ToppraError(f"Order {order} is not supported.")
        out = self._batch(ts.reshape(-1), order)[0].cpu().numpy().reshape(ts.shape + (self._path.as_batch().dof,))
        if getattr(self._path, "_scalar_dof", False):
            out = out[..., 0]
        return out[0] if scalar else out

    def plot_parametrization(self, show=False, n_sample=500):
        """Four panels: s(t), sd(s), the retimed joint positions and the original path (parametrizer.py:131-158).  Needs
        matplotlib, which is imported here only."""
        import matplotlib.pyplot as plt
        ts = np.linspace(self.path_interval[0], self.path_interval[1], n_sample)
        seg = np.clip(np.searchsorted(self._ts, ts, side="right") - 1, 0, len(self._us) - 1)   # plot-only: s(t), sd(t)
        dt = ts - self._ts[seg]
        vs = self._velocities[seg] + self._us[seg] * dt
        ss = self._ss[seg] + dt * (self._velocities[seg] + 0.5 * self._us[seg] * dt)
        s_dense = np.linspace(self._ss[0], self._ss[-1], n_sample)
        panels = (("path(time)", [(ts, ss, "-", "s(t)"), (self._ts, self._ss, "o", "input")]),
                  ("velocity(path)", [(ss, vs, "-", "v(s)"), (self._ss, self._velocities, "o", "input")]),
                  ("retimed path", [(ts, self(ts, 0), "-", None)]),
                  ("original path", [(s_dense, self._path(s_dense), "-", None)]))
        for k, (title, curves) in enumerate(panels):
            plt.subplot(2, 2, k + 1)
            for x, y, style, label in curves:
                plt.plot(x, y, style, label=label)
            if any(c[3] for c in curves):
                plt.legend()
            plt.title(title)
        plt.tight_layout()
        if show:
            plt.show()


class BatchParametrizeSpline(object):
    """ParametrizeSpline (reference parametrizer.py:161-196) for B paths on the GPU: the time-stamp recurrence with its
    two data-dependent rules (tb_spline_time_stamps, csrc/tb_frows.cu), q at the kept gridpoints (tb_ppoly_eval) and the
    clamped re-fit (tb_spline_fit with first-derivative boundary values q'(s) * sd).

    path: BatchSplineInterpolator; gridpoints: [G] or [B, G]; velocities: sd [B, G]; glen: optional int32 [B] for ragged
    grids.  The knot lists are ragged whenever the grids are, or an increment below 1e-8 is dropped: paths are then
    grouped by knot count and each group is fitted with one launch (`self.groups`: list of (path indices,
    BatchSplineInterpolator over time)).  `durations` is a CUDA tensor [B]."""

    def __init__(self, path, gridpoints, velocities, glen=None):
        torch = engine.torch_mod()
        self._path = path
        dev = path.device
        d_grid = engine.as_device(gridpoints, dev)
        d_sd = engine.as_device(velocities, dev)
        B, G = d_sd.shape
        if glen is not None and d_grid.dim() == 1:
            d_grid = d_grid.expand(B, G).contiguous()
        self.t_knots, self.s_knots, self.nkeep = engine.spline_time_stamps(d_sd, d_grid, glen)
        q_knots = path.eval_device(self.s_knots, 0)                               # [B, G, dof]
        s0 = d_grid[..., 0:1] if d_grid.dim() == 2 else d_grid[0:1]
        if glen is None:
            s1 = d_grid[..., -1:] if d_grid.dim() == 2 else d_grid[-1:]
            v1 = d_sd[:, -1]
        else:
            last = (glen.to(torch.int64) - 1).unsqueeze(1)
            s1 = torch.gather(d_grid, 1, last)
            v1 = torch.gather(d_sd, 1, last)[:, 0]
        # boundary conditions: first derivatives q'(path_interval) * sd at both ends (reference :189-196)
        bc0 = path.eval_device(s0, 1)[:, 0, :] * d_sd[:, 0:1]
        bc1 = path.eval_device(s1, 1)[:, 0, :] * v1.unsqueeze(1)
        n_host = self.nkeep.cpu().numpy()       # grouping by knot count needs the counts on the host (B ints)
        self.durations = torch.gather(self.t_knots, 1, (self.nkeep.to(torch.int64) - 1).unsqueeze(1))[:, 0]
        self.groups = []
        from .interpolator import BatchSplineInterpolator
        for n in np.unique(n_host):
            idx = np.nonzero(n_host == n)[0]
            sel = engine.as_device(idx.astype(np.int64), dev, dtype=torch.int64)
            full = len(idx) == B
            t = self.t_knots[:, :n] if full else self.t_knots[sel, :n]
            q = q_knots[:, :n] if full else q_knots[sel, :n]
            b0, b1 = (bc0, bc1) if full else (bc0[sel], bc1[sel])
            fit = BatchSplineInterpolator(t.contiguous(), q.contiguous(), ((1, b0.contiguous()), (1, b1.contiguous())),
                                          device=dev, validate=False)
            self.groups.append((idx, fit))

    def __call__(self, ts, order=0):
        """ts: [M] shared or [B, M] times -> CUDA tensor [B, M, dof]."""
        torch = engine.torch_mod()
        ts = engine.as_device(ts, self._path.device)
        B = self.t_knots.shape[0]
        out = None
        for idx, fit in self.groups:
            t = ts if ts.dim() == 1 else (ts if len(idx) == B else ts[engine.as_device(idx.astype(np.int64), ts.device, dtype=torch.int64)])
            val = fit.eval_device(t.contiguous(), order)
            if len(idx) == B:
                return val
            if out is None:
                out = torch.empty((B,) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
            out[engine.as_device(idx.astype(np.int64), val.device, dtype=torch.int64)] = val
        return out


class ParametrizeSpline(SplineInterpolator):
    """Output trajectory by cubic-spline interpolation of q(s_i) at the gridpoint time stamps
    t_i = t_{i-1} + ds / mean(sd_{i-1}, sd_i) (5 s for a stalled segment; increments < 1e-8 dropped), with the
    first derivatives at both ends clamped to q'(s) * sd (reference parametrizer.py:161-196).  The recurrence runs on
    the GPU (tb_spline_time_stamps): this is the B = 1 case of `BatchParametrizeSpline`."""

    def __init__(self, path, gridpoints, velocities):
        gridpoints = np.ascontiguousarray(gridpoints, dtype=np.float64)
        velocities = np.ascontiguousarray(velocities, dtype=np.float64)
        bpath = path.as_batch()
        t, s, nkeep = engine.spline_time_stamps(engine.as_device(velocities[None], bpath.device),
                                                engine.as_device(gridpoints, bpath.device))
        n = int(nkeep[0].item())
        t_grid = t[0, :n].cpu().numpy()
        q_grid = path(s[0, :n].cpu().numpy())
        bc = ((1, path(path.path_interval[0], 1) * velocities[0]),
              (1, path(path.path_interval[1], 1) * velocities[-1]))
        super(ParametrizeSpline, self).__init__(t_grid, q_grid, bc)
