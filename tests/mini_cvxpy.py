"""TEST INFRASTRUCTURE: a tiny linear-programming subset of the cvxpy modelling API, backed by scipy.optimize.linprog.

The reference's tests validate constraint parameters and stage LPs against cvxpy (tests/tests/constraint/
test_joint_velocity.py:50-83, tests/tests/lpsolvers/seidel/test_lp2d.py:74-116, tests/tests/solverwrapper/
test_basic_can_linear.py:85-164).  cvxpy is not installed in this image; tests/ref_suite_plugin.py registers this module
under the name `cvxpy` so that those tests validate against an INDEPENDENT LP solver (HiGHS) instead of being skipped.

Only what those tests use exists: Variable (scalar or 1-D), affine arithmetic with numpy operands (`*` is scaling when one
side is a scalar and matrix / inner product otherwise, as in the cvxpy version the reference was written against),
<=, >=, Minimize / Maximize, Problem.solve / .status / .value, Variable.value.  quad_form and norm (QP / SOCP) skip."""
import numpy as np
from scipy.optimize import linprog

ECOS = "ECOS"
CVXOPT = "CVXOPT"


class SolverError(Exception):
    pass


def _skip(what):
    import pytest
    pytest.skip("mini_cvxpy: %s is outside the LP subset" % what)


class Expr(object):
    """Affine expression sum_v M_v v + const with shape () or (m,)."""

    __array_ufunc__ = None      # numpy operands defer to the reflected operators below

    def __init__(self, terms, const, scalar):
        self.terms = terms              # {Variable: (m, var.size) matrix}
        self.const = np.asarray(const, dtype=float).reshape(-1)
        self.scalar = scalar            # shape () (m == 1) or (m,)

    @property
    def m(self):
        return self.const.shape[0]

    # ---- helpers ----------------------------------------------------------------------------------------------
    @staticmethod
    def _lift(other):
        if isinstance(other, Expr):
            return other
        arr = np.asarray(other, dtype=float)
        if arr.ndim > 1:
            raise NotImplementedError("mini_cvxpy: only scalars and vectors")
        return Expr({}, arr.reshape(-1), arr.ndim == 0)

    def _broadcast(self, m):
        if self.m == m:
            return self
        assert self.m == 1, "mini_cvxpy: shape mismatch"
        return Expr({v: np.repeat(M, m, axis=0) for v, M in self.terms.items()}, np.repeat(self.const, m), False)

    def __add__(self, other):
        other = Expr._lift(other)
        m = max(self.m, other.m)
        a, b = self._broadcast(m), other._broadcast(m)
        terms = dict(a.terms)
        for v, M in b.terms.items():
            terms[v] = terms[v] + M if v in terms else M
        return Expr(terms, a.const + b.const, self.scalar and other.scalar)

    __radd__ = __add__

    def __neg__(self):
        return Expr({v: -M for v, M in self.terms.items()}, -self.const, self.scalar)

    def __sub__(self, other):
        return self + (-Expr._lift(other))

    def __rsub__(self, other):
        return Expr._lift(other) + (-self)

    def _times(self, coef):
        """coef (numeric) * self."""
        if isinstance(coef, Expr):
            if coef.terms and self.terms:
                raise NotImplementedError("mini_cvxpy: product of two expressions")
            if not coef.terms:
                return self._times(coef.const[0] if coef.scalar else coef.const)
            return coef._times(self.const[0] if self.scalar else self.const)
        coef = np.asarray(coef, dtype=float)
        if coef.ndim == 0:
            return Expr({v: coef * M for v, M in self.terms.items()}, coef * self.const, self.scalar)
        if coef.ndim == 1 and self.scalar:                       # vector of coefficients times a scalar expression
            col = coef.reshape(-1, 1)
            return Expr({v: col * M for v, M in self.terms.items()}, coef * self.const[0], False)
        if coef.ndim == 1:                                       # inner product
            assert coef.shape[0] == self.m
            return Expr({v: coef.reshape(1, -1).dot(M) for v, M in self.terms.items()}, [coef.dot(self.const)], True)
        if coef.ndim == 2:                                       # matrix times vector expression
            assert coef.shape[1] == self.m
            return Expr({v: coef.dot(M) for v, M in self.terms.items()}, coef.dot(self.const), False)
        raise NotImplementedError

    def __mul__(self, other):
        return self._times(other)

    __rmul__ = __mul__
    __matmul__ = __mul__
    __rmatmul__ = __mul__

    def __getitem__(self, idx):
        return Expr({v: M[idx:idx + 1] for v, M in self.terms.items()}, self.const[idx:idx + 1], True)

    def __le__(self, other):
        return Constraint(self - other)

    def __ge__(self, other):
        return Constraint(Expr._lift(other) - self)

    # no __eq__: expressions are dictionary keys (identity); equality constraints are outside the subset


class Variable(Expr):
    def __init__(self, n=None, **kwargs):
        self.size = 1 if n is None else int(n)
        self._value = None
        super(Variable, self).__init__({self: np.eye(self.size)}, np.zeros(self.size), n is None)

    @property
    def value(self):
        if self._value is None:
            return None
        return float(self._value[0]) if self.scalar else self._value.copy()


class Constraint(object):
    """expr <= 0, row by row."""

    def __init__(self, expr):
        self.expr = expr


class _Objective(object):
    def __init__(self, expr, sign):
        expr = Expr._lift(expr)
        assert expr.m == 1, "mini_cvxpy: the objective must be a scalar"
        self.expr, self.sign = expr, sign


def Minimize(expr):
    return _Objective(expr, 1.0)


def Maximize(expr):
    return _Objective(expr, -1.0)


def quad_form(x, P):
    _skip("quad_form")


def norm(x, p=2):
    _skip("norm")


class Problem(object):
    def __init__(self, objective, constraints=()):
        self.objective, self.constraints = objective, list(constraints)
        self.status, self.value = None, None

    def solve(self, *args, **kwargs):
        variables = []
        for e in [self.objective.expr] + [c.expr for c in self.constraints]:
            for v in e.terms:
                if v not in variables:
                    variables.append(v)
        offset, n = {}, 0
        for v in variables:
            offset[v] = n
            n += v.size

        def dense(e):
            A = np.zeros((e.m, n))
            for v, M in e.terms.items():
                A[:, offset[v]:offset[v] + v.size] += M
            return A

        cost = self.objective.sign * dense(self.objective.expr)[0]
        A = np.vstack([dense(c.expr) for c in self.constraints]) if self.constraints else np.zeros((0, n))
        b = -np.concatenate([c.expr.const for c in self.constraints]) if self.constraints else np.zeros(0)
        res = linprog(cost, A_ub=A, b_ub=b, bounds=[(None, None)] * n, method="highs")
        if res.status == 0:
            self.status = "optimal"
            for v in variables:
                v._value = np.array(res.x[offset[v]:offset[v] + v.size])
            self.value = float(self.objective.sign * res.fun + self.objective.expr.const[0])
        elif res.status == 2:
            self.status, self.value = "infeasible", (np.inf if self.objective.sign > 0 else -np.inf)
        elif res.status == 3:
            self.status, self.value = "unbounded", (-np.inf if self.objective.sign > 0 else np.inf)
        else:
            raise SolverError(res.message)
        if self.status != "optimal":
            for v in variables:
                v._value = None
        return self.value
