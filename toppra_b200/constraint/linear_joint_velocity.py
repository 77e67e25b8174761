"""Joint velocity constraint — same surface as the reference
`toppra/constraint/linear_joint_velocity.py:7-53`; numbers from csrc/tb_coeff.cu."""
import numpy as np

from .linear_constraint import LinearConstraint, RecordContext
from .. import engine


def _limits(lim, what):
    """(dof,) symmetric | (dof, 2) | batched (B, dof, 2) -> float array with last dim 2."""
    lim = np.array(lim, dtype=float)
    if np.isnan(lim).any():
        raise ValueError("Bad %s given: %s" % (what, lim))
    if lim.ndim == 1:
        lim = np.vstack((-np.array(lim), np.array(lim))).T
    return lim


def single_path_context(path, gridpoints):
    """RecordContext for the reference-style call `constraint.compute_constraint_params(path, gridpoints)`."""
    grid = np.ascontiguousarray(gridpoints, dtype=np.float64)
    bpath = path.as_batch()
    return RecordContext(bpath, engine.as_device(grid, bpath.device), grid, path)


class JointVelocityConstraint(LinearConstraint):
    """A Joint Velocity Constraint class.

    Parameters
    ----------
    vlim: np.ndarray
        Shape (dof, 2): lower and upper velocity bounds of joint j are vlim[j, 0], vlim[j, 1];
        shape (dof,): symmetric bounds.  toppra_b200 extension: shape (B, dof, 2) gives every path of a
        batch its own limits."""

    def __init__(self, vlim):
        super(JointVelocityConstraint, self).__init__()
        self.vlim = _limits(vlim, "velocity")
        self.dof = self.vlim.shape[-2]
        self._assert_valid_limits()
        self._d_cache = {}

    def _assert_valid_limits(self):
        assert self.vlim.shape[-1] == 2, "Wrong input shape."
        flat = self.vlim.reshape(-1, 2)
        bad = np.nonzero(flat[:, 0] >= flat[:, 1])[0]
        if len(bad):
            raise ValueError("Bad velocity limits: {:} (lower limit) > {:} (higher limit)".format(
                flat[bad[0], 0], flat[bad[0], 1]))
        self._format_string = "    Velocity limit: \n"
        if self.vlim.ndim == 2:
            for i in range(self.vlim.shape[0]):
                self._format_string += "      J{:d}: {:}".format(i + 1, self.vlim[i]) + "\n"

    @classmethod
    def from_device(cls, d_vlim):
        """Limits already resident on the GPU: a (dof, 2) or (B, dof, 2) float64 CUDA tensor (toppra_b200 extension for
        pipelines whose inputs never touch the host; the caller vouches for lower < upper)."""
        obj = cls.__new__(cls)
        LinearConstraint.__init__(obj)
        if d_vlim.dim() not in (2, 3) or d_vlim.shape[-1] != 2:
            raise ValueError("device limits must have shape (dof, 2) or (B, dof, 2)")
        obj.vlim, obj.dof = d_vlim, int(d_vlim.shape[-2])
        obj._format_string = "    Velocity limit: (device tensor)\n"
        obj._d_cache = {str(d_vlim.device): d_vlim}
        return obj

    def device_limits(self, device):
        key = str(device)
        if key not in self._d_cache:
            self._d_cache[key] = engine.as_device(self.vlim, device)
        return self._d_cache[key]

    def compute_constraint_params(self, path, gridpoints):
        if path.dof != self.get_dof():
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.get_dof(), path.dof))
        ctx = single_path_context(path, gridpoints)
        records, _ = engine.alloc_records(1, ctx.G, 0, ctx.device)
        # raw xbound of _create_velocity_constraint (no clipping to the solver box): write_xbound = 2
        engine.coeff_velacc(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, ctx.limits(self.device_limits(ctx.device)), None,
                            False, records, 0, 0, 2)
        xbound = records[0, :, 0:2].cpu().numpy().copy()
        return None, None, None, None, None, None, xbound

    # device protocol
    def num_rows(self, ctx):
        return 0

    def append_records(self, ctx, records, R_total, row0):
        if ctx.bpath.dof != self.get_dof():
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.get_dof(), ctx.bpath.dof))
        engine.coeff_velacc(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, ctx.limits(self.device_limits(ctx.device)), None,
                            False, records, R_total, 0, 3)


class JointVelocityConstraintVarying(LinearConstraint):
    """Joint velocity limits that vary along the path (reference linear_joint_velocity.py:56-87).

    vlim_func: (float) -> np.ndarray (dof, 2): lower and upper velocity bounds at path position s.  The function is
    user code, evaluated on the host once per gridpoint; the bound itself is computed on the GPU."""

    def __init__(self, vlim_func):
        super(JointVelocityConstraintVarying, self).__init__()
        self.dof = vlim_func(0).shape[0]
        self._format_string = "    Varying Velocity limit: \n"
        self.vlim_func = vlim_func

    def _limits_grid(self, ctx):
        if ctx.grid_host is None:
            raise NotImplementedError("JointVelocityConstraintVarying needs host gridpoints shared by all paths")
        vlim_grid = np.array([self.vlim_func(s) for s in ctx.grid_host], dtype=np.float64)
        return engine.as_device(vlim_grid, ctx.device)

    def compute_constraint_params(self, path, gridpoints):
        if path.dof != self.get_dof():
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.get_dof(), path.dof))
        ctx = single_path_context(path, gridpoints)
        records, _ = engine.alloc_records(1, ctx.G, 0, ctx.device)
        engine.xbound_varying(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, self._limits_grid(ctx), records, 0, 2)
        xbound = records[0, :, 0:2].cpu().numpy().copy()
        return None, None, None, None, None, None, xbound

    def num_rows(self, ctx):
        return 0

    def append_records(self, ctx, records, R_total, row0):
        if ctx.bpath.dof != self.get_dof():
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.get_dof(), ctx.bpath.dof))
        engine.xbound_varying(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, self._limits_grid(ctx), records, R_total, 3)
