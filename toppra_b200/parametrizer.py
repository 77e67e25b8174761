"""Output trajectories from (gridpoints, sd) — same surface as the reference `toppra/parametrizer.py`
(`ParametrizeConstAccel` :23-158, `ParametrizeSpline` :161-196).

SURVEY.md §8 (f1), "next" row: the O(G) time-stamp recurrence below is host numpy for a single path; the spline
re-fit and every path evaluation run on the GPU through SplineInterpolator."""
import logging

import numpy as np

from .constants import TINY
from .exceptions import ToppraError
from .interpolator import AbstractGeometricPath, SplineInterpolator

logger = logging.getLogger(__name__)


class ParametrizeConstAccel(AbstractGeometricPath):
    """Output trajectory under the piecewise-constant path-acceleration assumption:
    on [s_i, s_{i+1}]  u_i = (x_{i+1} - x_i) / (2 ds),  t_{i+1} = t_i + 2 ds / (sd_i + sd_{i+1})."""

    def __init__(self, path, gridpoints, velocities):
        self._path = path
        self._ss = np.array(gridpoints, dtype=np.float64)
        self._velocities = np.array(velocities, dtype=np.float64)
        self._xs = self._velocities ** 2
        assert self._ss.shape[0] == self._velocities.shape[0]
        assert len(self._ss.shape) == 1
        assert np.all(self._velocities >= 0)
        ds = np.diff(self._ss)
        self._us = 0.5 * (self._xs[1:] - self._xs[:-1]) / ds
        dts = 2 * ds / (self._velocities[:-1] + self._velocities[1:])
        ts = np.zeros(len(self._ss))
        for i in range(len(dts)):  # sequential sum keeps the reference's rounding order
            ts[i + 1] = ts[i] + dts[i]
        self._ts = ts

    @property
    def dof(self):
        return self._path.dof

    @property
    def path_interval(self):
        return np.array([self._ts[0], self._ts[-1]])

    @property
    def duration(self):
        return self.path_interval[1] - self.path_interval[0]

    def _eval_params(self, ts):
        idx = np.searchsorted(self._ts, ts, side="right") - 1
        idx = np.where(idx == len(self._us), idx - 1, idx)
        dt = ts - self._ts[idx]
        us = self._us[idx]
        vs = self._velocities[idx] + dt * us
        ss = self._ss[idx] + dt * self._velocities[idx] + 0.5 * dt ** 2 * us
        return ss, vs, us

    def __call__(self, ts, order=0):
        scalar = isinstance(ts, (int, float))
        ts = np.array([ts], dtype=float) if scalar else np.asarray(ts, dtype=float)
        ss, vs, us = self._eval_params(ts)
        if order == 0:
            out = self._path(ss)
        elif order == 1:
            out = np.multiply(self._path(ss, 1), vs[:, np.newaxis])
        elif order == 2:
            out = (np.multiply(self._path(ss, 2), vs[:, np.newaxis] ** 2)
                   + np.multiply(self._path(ss, 1), us[:, np.newaxis]))
        else:
            raise ToppraError(f"Order {order} is not supported.")
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
