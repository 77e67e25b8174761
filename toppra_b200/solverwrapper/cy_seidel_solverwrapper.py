"""Module-level mirror of the reference's compiled Seidel module, so `import toppra.solverwrapper.cy_seidel_solverwrapper`
call sites keep working after the switch (reference: toppra/solverwrapper/cy_seidel_solverwrapper.pyx).

  solve_lp1d   pyx:42-63   one 1-variable LP   -> (result, optval, optvar, active)
  solve_lp2d   pyx:65-87   one 2-variable LP   -> (result, optval, optvar[2], active[2])
  seidelWrapper pyx:392-703 the stage-wise solver wrapper (here: B200SolverWrapper)

The two LP functions run ONE problem through the batched device kernels (tb_lp1d_batch / tb_lp2d_batch, csrc/tb_scan.cu):
they exist for API parity and for the reference's own known-answer tests, not for speed — batches go through
`engine.lp1d_batch` / `engine.lp2d_batch` or, better, a whole scan."""
import numpy as np

from .. import engine
from .solverwrapper import B200SolverWrapper


def _rows(*arrays):
    """Row arrays of one LP as [1, n]; None or an empty list mean 'no rows' (pyx:56-61, 75-85)."""
    if arrays[0] is None or len(arrays[0]) == 0:
        return tuple(np.zeros((1, 0)) for _ in arrays)
    return tuple(np.ascontiguousarray(x, dtype=np.float64).reshape(1, -1) for x in arrays)


def solve_lp1d(v, a, b, low, high):
    """max v[0] x + v[1]  s.t.  a x + b <= 0, low <= x <= high.  result 1 = optimal, 0 = infeasible; `active` is the index
    of the binding row, -1 / -2 for the lower / upper bound (pyx:93-144)."""
    a, b = _rows(a, b)
    res, optval, optvar, active = engine.lp1d_batch(np.asarray(v, dtype=np.float64).reshape(1, 2), a, b,
                                                    np.array([low], dtype=np.float64), np.array([high], dtype=np.float64))
    return int(res[0]), float(optval[0]), float(optvar[0]), int(active[0])


def solve_lp2d(v, a, b, c, low, high, active_c):
    """max v[0] x0 + v[1] x1 + v[2]  s.t.  a x0 + b x1 + c <= 0, low <= x <= high, warm-started from the row pair
    `active_c` (pyx:149-390).  Returns (result, optval, optvar[2], active[2])."""
    a, b, c = _rows(a, b, c)
    res, optval, optvar, active = engine.lp2d_batch(np.asarray(v, dtype=np.float64).reshape(1, 3), a, b, c,
                                                    np.asarray(low, dtype=np.float64).reshape(1, 2),
                                                    np.asarray(high, dtype=np.float64).reshape(1, 2),
                                                    np.asarray(active_c).reshape(1, 2))
    return int(res[0]), float(optval[0]), optvar[0].copy(), active[0].copy()


class seidelWrapper(B200SolverWrapper):
    """The reference's constructor signature: `solve_lp1d` defaults to 0 there (pyx:425); the algorithms pass True
    (reachability_algorithm.py:123-125)."""

    def __init__(self, constraint_list, path, path_discretization, solve_lp1d=0):
        super(seidelWrapper, self).__init__(constraint_list, path, path_discretization, solve_lp1d=bool(solve_lp1d))


__all__ = ["solve_lp1d", "solve_lp2d", "seidelWrapper"]
