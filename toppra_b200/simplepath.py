"""`SimplePath` — cubic-Hermite path through positions (and optional first derivatives); public surface of the reference's
`toppra/simplepath.py:7-91`.  The reference keeps one scipy `BPoly.from_derivatives` per joint; here the same cubics are
written once in the local power basis and handed to the GPU path evaluation (`PPolyPath`), so values agree with the
reference to rounding (the basis differs), not bit for bit."""
import numpy as np

from .interpolator import PPolyPath


class SimplePath(PPolyPath):
    """x: (n,) positions of the waypoints; y: (n,) or (n, dof) values; yd: first derivatives or None (then: zero at
    both ends, central differences inside, simplepath.py:62-72)."""

    def __init__(self, x, y, yd=None, device=None):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        if yd is not None:
            yd = np.asarray(yd, dtype=np.float64)
            if yd.ndim == 1:
                yd = yd.reshape(-1, 1)
        if x.ndim != 1 or x.shape[0] != y.shape[0] or x.shape[0] < 2:
            raise ValueError("SimplePath needs x of shape (n,) and y of shape (n,) or (n, dof), n >= 2")
        self._y = y
        self._yd = yd
        d = self._slopes(x, y, yd)
        h = np.diff(x)[:, None]
        y0, y1, d0, d1 = y[:-1], y[1:], d[:-1], d[1:]
        coeffs = np.stack(((2 * (y0 - y1) + h * (d0 + d1)) / h ** 3,      # cubic Hermite on [x_i, x_i+1] in powers of
                           (3 * (y1 - y0) - h * (2 * d0 + d1)) / h ** 2,  # (s - x_i): value/slope match at both ends
                           d0, y0))
        super(SimplePath, self).__init__(coeffs, x, device=device)

    @staticmethod
    def _slopes(x, y, yd):
        if yd is not None:
            return np.array(yd, dtype=np.float64)
        d = np.zeros_like(y)
        d[1:-1] = (y[2:] - y[:-2]) / (x[2:] - x[:-2])[:, None]
        return d

    @property
    def waypoints(self):
        return self._y
