"""Output trajectories from (gridpoints, sd) — same surface as the reference `toppra/parametrizer.py`
(`ParametrizeConstAccel` :23-158, `ParametrizeSpline` :161-196).

SURVEY.md §8 (f1), "next" row: the O(G) time-stamp recurrence below is host numpy for a single path; the spline
re-fit and every path evaluation run on the GPU through SplineInterpolator."""
import logging

import numpy as np

from . import engine
from .constants import TINY
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


class ParametrizeSpline(SplineInterpolator):
    """Output trajectory by cubic-spline interpolation of q(s_i) at the gridpoint time stamps
    t_i = t_{i-1} + ds / mean(sd_{i-1}, sd_i) (5 s for a stalled segment; increments < 1e-8 dropped), with the
    first derivatives at both ends clamped to q'(s) * sd."""

    def __init__(self, path, gridpoints, velocities):
        gridpoints = np.asarray(gridpoints, dtype=np.float64)
        velocities = np.asarray(velocities, dtype=np.float64)
        t_grid = np.zeros_like(gridpoints)
        skip = []
        for i in range(1, len(t_grid)):
            sd_average = (velocities[i - 1] + velocities[i]) / 2
            delta_s = gridpoints[i] - gridpoints[i - 1]
            delta_t = delta_s / sd_average if sd_average > TINY else 5
            t_grid[i] = t_grid[i - 1] + delta_t
            if delta_t < TINY:
                skip.append(i)
        t_grid = np.delete(t_grid, skip)
        gridpoints = np.delete(gridpoints, skip)
        q_grid = path(gridpoints)
        bc = ((1, path(path.path_interval[0], 1) * velocities[0]),
              (1, path(path.path_interval[1], 1) * velocities[-1]))
        super(ParametrizeSpline, self).__init__(t_grid, q_grid, bc)
