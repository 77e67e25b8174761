"""Joint acceleration constraint — same surface as the reference
`toppra/constraint/linear_joint_acceleration.py:7-104`; numbers from csrc/tb_coeff.cu."""
import numpy as np

from .constraint import DiscretizationType
from .linear_constraint import LinearConstraint
from .linear_joint_velocity import _limits, single_path_context
from .. import engine


class JointAccelerationConstraint(LinearConstraint):
    """qdd_min <= q'(s_i) u_i + q''(s_i) x_i <= qdd_max, i.e. a = q', b = q'', c = 0, F = [I; -I],
    g = [qdd_max; -qdd_min].

    Parameters
    ----------
    alim: (dof, 2) lower/upper bounds, (dof,) symmetric, or batched (B, dof, 2) (toppra_b200 extension).
    discretization_scheme: Collocation (0) or Interpolation (1, default)."""

    def __init__(self, alim, discretization_scheme=DiscretizationType.Interpolation):
        super(JointAccelerationConstraint, self).__init__()
        self.alim = _limits(alim, "velocity")
        self.dof = self.alim.shape[-2]
        self.set_discretization_type(discretization_scheme)
        assert self.alim.shape[-1] == 2, "Wrong input shape."
        self._format_string = "    Acceleration limit: \n"
        if self.alim.ndim == 2:
            for i in range(self.alim.shape[0]):
                self._format_string += "      J{:d}: {:}".format(i + 1, self.alim[i]) + "\n"
        self.identical = True
        self._d_cache = {}

    @classmethod
    def from_device(cls, d_alim, discretization_scheme=DiscretizationType.Interpolation):
        """Limits already resident on the GPU: a (dof, 2) or (B, dof, 2) float64 CUDA tensor (toppra_b200 extension)."""
        obj = cls.__new__(cls)
        LinearConstraint.__init__(obj)
        if d_alim.dim() not in (2, 3) or d_alim.shape[-1] != 2:
            raise ValueError("device limits must have shape (dof, 2) or (B, dof, 2)")
        obj.alim, obj.dof = d_alim, int(d_alim.shape[-2])
        obj.set_discretization_type(discretization_scheme)
        obj._format_string = "    Acceleration limit: (device tensor)\n"
        obj.identical = True
        obj._d_cache = {str(d_alim.device): d_alim}
        return obj

    def device_limits(self, device):
        key = str(device)
        if key not in self._d_cache:
            self._d_cache[key] = engine.as_device(self.alim, device)
        return self._d_cache[key]

    @property
    def interpolation(self):
        return self.discretization_type == DiscretizationType.Interpolation

    def compute_constraint_params(self, path, gridpoints, *args, **kwargs):
        if path.dof != self.dof:
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.dof, path.dof))
        if self.alim.ndim != 2:
            raise ValueError("compute_constraint_params needs (dof, 2) limits; batched limits go through BatchTOPPRA")
        ctx = single_path_context(path, gridpoints)
        dof = self.dof
        R = self.num_rows(ctx)
        records, _ = engine.alloc_records(1, ctx.G, R, ctx.device)
        engine.coeff_velacc(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, None, ctx.limits(self.device_limits(ctx.device)),
                            self.interpolation, records, R, 0, 0)
        rec = records[0].cpu().numpy()
        arow, brow = rec[:, 0:R], rec[:, R:2 * R]
        F_single = np.zeros((dof * 2, dof))
        g_single = np.zeros(dof * 2)
        g_single[0:dof] = self.alim[:, 1]
        g_single[dof:] = -self.alim[:, 0]
        F_single[0:dof, :] = np.eye(dof)
        F_single[dof:, :] = -np.eye(dof)
        if not self.interpolation:
            a = arow[:, :dof].copy()
            b = brow[:, :dof].copy()
            return a, b, np.zeros_like(a), F_single, g_single, None, None
        # rows are [a, -a, a+, -a+] (F = blkdiag([I;-I],[I;-I])): columns of the lifted a are (a, a+)
        a = np.concatenate((arow[:, :dof], arow[:, 2 * dof:3 * dof]), axis=1)
        b = np.concatenate((brow[:, :dof], brow[:, 2 * dof:3 * dof]), axis=1)
        F = np.zeros((4 * dof, 2 * dof))
        F[:2 * dof, :dof] = F_single
        F[2 * dof:, dof:] = F_single
        g = np.concatenate((g_single, g_single))
        return a, b, np.zeros_like(a), F, g, None, None

    # device protocol
    def num_rows(self, ctx):
        return (4 if self.interpolation else 2) * self.dof

    def append_records(self, ctx, records, R_total, row0):
        if ctx.bpath.dof != self.dof:
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.dof, ctx.bpath.dof))
        engine.coeff_velacc(ctx.bpath.d_ppoly, ctx.bpath.d_ss, ctx.d_grid, None, ctx.limits(self.device_limits(ctx.device)),
                            self.interpolation, records, R_total, row0, 0)
