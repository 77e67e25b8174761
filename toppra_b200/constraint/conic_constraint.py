"""Conic constraints — same surface as the reference `toppra/constraint/conic_constraint.py:6-124`.

`RobustLinearConstraint` robustifies a CanonicalLinear constraint against perturbations
[da, db, dc] = diag(ru, rx, rc) e, ||e||_2 <= 1 of every row:
    a u + b x + c + || diag(ru, rx, rc) [u, x, 1] ||_2 <= 0.
The reference solves the resulting stage problems with ECOS; here they are solved exactly on the GPU
(csrc/tb_robust.cu).  Parity with ECOS holds to solver tolerance only (unpinned)."""
import numpy as np

from .constraint import Constraint, ConstraintType, DiscretizationType
from .linear_joint_velocity import single_path_context
from .. import engine


class ConicConstraint(Constraint):
    """Base class for all canonical conic constraints (reference conic_constraint.py:6-44)."""

    constraint_type = ConstraintType.CanonicalConic      # class-level defaults; instances override dof / scheme
    discretization_type = DiscretizationType.Collocation
    n_extra_vars = 0
    dof = -1
    _format_string = ""

    def compute_constraint_params(self, path, gridpoints):
        raise NotImplementedError("%s does not say how its rows and ellipsoids are computed" % type(self).__name__)


class RobustLinearConstraint(ConicConstraint):
    """Robustified version of a CanonicalLinear constraint.

    Parameters
    ----------
    cnst: the base LinearConstraint
    ellipsoid_axes_lengths: (3,) non-negative axes (ru, rx, rc) of the perturbation ellipsoid
    discretization_scheme: Collocation (default, as in the reference) or Interpolation"""

    def __init__(self, cnst, ellipsoid_axes_lengths, discretization_scheme=DiscretizationType.Collocation):
        axes = np.asarray(ellipsoid_axes_lengths, dtype=np.float64).reshape(-1)
        if axes.shape != (3,):
            raise ValueError("ellipsoid_axes_lengths must hold the three axes (ru, rx, rc); got {:}".format(
                ellipsoid_axes_lengths))
        if (axes < 0).any():
            raise ValueError("Perturbation must be non-negative. Input {:}".format(ellipsoid_axes_lengths))
        assert cnst.get_constraint_type() == ConstraintType.CanonicalLinear, "only linear constraints can be robustified"
        self.base_constraint, self.ellipsoid_axes_lengths = cnst, ellipsoid_axes_lengths
        self.dof = cnst.get_dof()
        self.set_discretization_type(discretization_scheme)
        self._format_string = "    Robust constraint generated from a canonical linear constraint\n"

    def compute_constraint_params(self, path, gridpoints):
        """(a, b, c, P, ubound, xbound): rows a = F a0, b = F b0, c = F c0 - g of the base constraint
        (shape (N+1, d)) and P (N+1, d+2, 3, 3) = diag(ellipsoid axes)  (reference :95-124)."""
        self.base_constraint.set_discretization_type(self.discretization_type)
        ctx = single_path_context(path, gridpoints)
        d = self.base_constraint.num_rows(ctx)
        records, _ = engine.alloc_records(1, ctx.G, d, ctx.device)
        engine.init_bounds(records, d)
        self.base_constraint.append_records(ctx, records, d, 0)
        rec = records[0].cpu().numpy()
        a, b, c = rec[:, 0:d].copy(), rec[:, d:2 * d].copy(), rec[:, 2 * d:3 * d].copy()
        N = len(gridpoints) - 1
        P = np.zeros((N + 1, d + 2, 3, 3))
        P[:] = np.diag(self.ellipsoid_axes_lengths)
        # the base constraint's own ubound / xbound pass through (reference :97-99, :124); the in-scope bases
        # (JointAcceleration, SecondOrder, JointTorque) return None for both, user-defined ones may not
        u_, x_ = self.base_constraint.compute_constraint_params(path, gridpoints)[5:7]
        return a, b, c, P, u_, x_

    # device protocol (see LinearConstraint): the rows are those of the base constraint; the solver gets the
    # row range and the ellipsoid through `conic_info`
    def num_rows(self, ctx):
        self.base_constraint.set_discretization_type(self.discretization_type)
        return self.base_constraint.num_rows(ctx)

    def append_records(self, ctx, records, R_total, row0):
        self.base_constraint.set_discretization_type(self.discretization_type)
        self.base_constraint.append_records(ctx, records, R_total, row0)

    def ellipsoid(self):
        return np.asarray(self.ellipsoid_axes_lengths, dtype=np.float64).reshape(3)
