"""Cartesian velocity-norm constraint — the C++-only constraint of the reference
(`cpp/src/toppra/constraint/cartesian_velocity_norm.hpp:10-95`, `.cpp:23-54`) with a Python surface.

    v = J(p(s)) p'(s) sd  is the 6-D velocity of a frame;   ||v||_S^2 = v^T S v <= limit
    =>  one CanonicalLinear row per gridpoint:  a = 0,  b = v1^T S v1  (v1 = the frame velocity for sd = 1),  c = 0,
        F = [1],  g = [limit]                                                                 (.cpp:36-52)

`compute_velocity(q, qd) -> v[6]` is USER code (the C++ class leaves `computeVelocity` pure virtual), called once per
gridpoint on the host with q = p(s), qd = p'(s).  Constant (S, limit) gives an identical F/g like the C++ constant
constructor; `velocity_limit(s) -> (S, limit)` gives the varying form (`computeVelocityLimit`, .hpp:73-80)."""
import numpy as np

from .linear_constraint import LinearConstraint


class CartesianVelocityNorm(LinearConstraint):
    def __init__(self, compute_velocity, S=None, limit=None, velocity_limit=None, dof=None):
        super(CartesianVelocityNorm, self).__init__()
        self.compute_velocity = compute_velocity
        self.velocity_limit = velocity_limit
        self.dof = dof
        if velocity_limit is None:
            if S is None or limit is None:
                raise ValueError("CartesianVelocityNorm needs (S, limit) or velocity_limit(s) -> (S, limit)")
            self.S, self.limit = np.array(S, dtype=float), float(limit)
            self._check(self.S, self.limit)
        else:
            self.S, self.limit = np.zeros((6, 6)), 1.0  # like the C++ varying constructor: set per gridpoint
        self.identical = velocity_limit is None
        self._format_string = "    Velocity norm limit: {:}\n".format(self.limit if self.identical else "varying")

    @staticmethod
    def _check(S, limit):  # .cpp:16-21
        if limit < 0:
            raise ValueError("Velocity limit should be positive.")
        if np.shape(S) != (6, 6):
            raise ValueError("S matrix should be of size 6x6.")

    def get_dof(self):
        return self.dof

    def compute_constraint_params(self, path, gridpoints):
        if self.dof is not None and path.dof != self.dof:
            raise ValueError("Wrong dimension: constraint dof ({:d}) not equal to path dof ({:d})".format(
                self.dof, path.dof))
        gridpoints = np.asarray(gridpoints, dtype=float)
        G = gridpoints.shape[0]
        q, qd = path(gridpoints), path(gridpoints, 1)
        a, b, c = np.zeros((G, 1)), np.zeros((G, 1)), np.zeros((G, 1))
        F = np.ones((1, 1)) if self.identical else np.ones((G, 1, 1))
        g = np.full(1, self.limit) if self.identical else np.zeros((G, 1))
        S, limit = self.S, self.limit
        for i in range(G):
            if not self.identical:
                S, limit = self.velocity_limit(gridpoints[i])
                S, limit = np.asarray(S, dtype=float), float(limit)
                self._check(S, limit)
                g[i, 0] = limit
            v = np.asarray(self.compute_velocity(q[i], qd[i]), dtype=float)
            if v.shape != (6,):
                raise ValueError("compute_velocity must return the 6-D frame velocity")
            b[i, 0] = v.dot(S.dot(v))  # v^T (S v), .cpp:43-46
        return a, b, c, F, g, None, None
