"""Independent optimality certificate for the robust (conic) stage problems of BASELINE config 4.

The kernel (csrc/tb_robust.cu) and its C restatement (oracle/toppra_robust_oracle.c) share one closed form, so a
test that only compares the two proves nothing about optimality.  This module knows NOTHING of that closed form: it
evaluates the problem as the reference states it for ECOS (ecos_solverwrapper.py:90-207, conic_constraint.py:95-124)

    rows      f_j(u, x) = a_j u + b_j x + c_j + || (ru u, rx x, rc) ||_2  <= 0
    x_next    K_lo[i+1] <= x + 2 delta_i u <= K_hi[i+1]
    x box     max(xbound_lo, -1000) <= x <= min(1e4, xbound_hi, 1000)

directly, in 80-bit long doubles: F_i(x) = min_u max_j f_j(u, x) over the u-interval the x_next rows leave, by
golden-section search (f is convex in u).  x is feasible iff F_i(x) <= 0; the feasible set is convex, hence an
interval in x, so "K_hi is feasible and K_hi (1 + 1e-9) + 1e-12 is not" certifies maximality (likewise K_lo)."""
import numpy as np

LD = np.longdouble
ECOS_INFTY, ECOS_MAXX = 1000.0, 10000.0


def _F(x, a, b, c, ell, two_delta, klo, khi, iters=200):
    """min over u of the largest row value at x, per stage.  x, two_delta, klo, khi: [N]; a, b, c: [N, R] -> [N]
    (+inf where the x_next rows leave no u)."""
    x = x.astype(LD)
    ru, rx, rc = (LD(e) for e in ell)
    ulo = (klo.astype(LD) - x) / two_delta.astype(LD)
    uhi = (khi.astype(LD) - x) / two_delta.astype(LD)
    empty = ulo > uhi
    g2 = (rx * x) ** 2 + rc * rc
    bxc = b.astype(LD) * x[:, None] + c.astype(LD)
    aL = a.astype(LD)

    def fmax(u):
        nrm = np.sqrt((ru * u) ** 2 + g2)
        return np.max(aL * u[:, None] + bxc + nrm[:, None], axis=1)

    lo, hi = ulo.copy(), np.maximum(uhi, ulo)
    invphi = LD(0.6180339887498948482)
    x1 = hi - invphi * (hi - lo)
    x2 = lo + invphi * (hi - lo)
    f1, f2 = fmax(x1), fmax(x2)
    for _ in range(iters):
        left = f1 < f2
        hi = np.where(left, x2, hi)
        lo = np.where(left, lo, x1)
        nx1 = hi - invphi * (hi - lo)
        nx2 = lo + invphi * (hi - lo)
        x1n = np.where(left, nx1, x2)
        x2n = np.where(left, x1, nx2)
        f1n, f2n = fmax(x1n), fmax(x2n)
        x1, x2, f1, f2 = x1n, x2n, f1n, f2n
    best = np.minimum(np.minimum(f1, f2), np.minimum(fmax(ulo), fmax(np.maximum(uhi, ulo))))
    return np.where(empty, LD(np.inf), best)


def certify_controllable_sets(rows, xbound, grid, K, ell, feas_tol=1e-9):
    """rows [G, 3, R] (all rows robust), xbound [G, 2], K [G, 2] of ONE path.  Returns (worst feasibility residual of the
    K endpoints, number of endpoints certified extremal, number that sit on the x box)."""
    N = len(grid) - 1
    a, b, c = rows[:N, 0], rows[:N, 1], rows[:N, 2]
    two_delta = 2 * np.diff(grid)
    klo, khi = K[1:, 0], K[1:, 1]
    box_lo = np.maximum(xbound[:N, 0], -ECOS_INFTY)
    box_hi = np.minimum(np.minimum(xbound[:N, 1], ECOS_MAXX), ECOS_INFTY)
    xl, xu = K[:N, 0], K[:N, 1]
    resid = max(float(np.max(_F(xu, a, b, c, ell, two_delta, klo, khi))),
                float(np.max(_F(xl, a, b, c, ell, two_delta, klo, khi))))
    assert resid <= feas_tol, "K endpoint infeasible by %g" % resid
    # upper ends: on the box, or infeasible just beyond
    on_box_u = xu >= box_hi
    beyond = xu * (1 + 1e-9) + 1e-12
    Fb = _F(beyond, a, b, c, ell, two_delta, klo, khi)
    assert np.all(on_box_u | (Fb > 0)), "K upper end is not maximal at stages %s" % np.nonzero(~(on_box_u | (Fb > 0)))[0]
    # lower ends: on the box / clamped at 0 (reachability_algorithm.py:190-191), or infeasible just below
    on_box_l = (xl <= np.maximum(box_lo, 0.0))
    below = xl * (1 - 1e-9) - 1e-12
    Fl = _F(below, a, b, c, ell, two_delta, klo, khi)
    assert np.all(on_box_l | (Fl > 0)), "K lower end is not minimal at stages %s" % np.nonzero(~(on_box_l | (Fl > 0)))[0]
    return resid, int((~on_box_u).sum() + (~on_box_l).sum()), int(on_box_u.sum() + on_box_l.sum())


def slsqp_extreme_x(a, b, c, ell, two_delta, klo, khi, box_lo, box_hi, sign, x0):
    """max (sign=+1) / min (sign=-1) of x over the stage's cone program with an unrelated solver (scipy SLSQP on the
    smooth form: rc > 0 makes the norm differentiable).  Returns x* or None if the solver did not converge."""
    from scipy.optimize import minimize
    ru, rx, rc = ell

    def cons(z):
        u, x = z
        nrm = np.sqrt((ru * u) ** 2 + (rx * x) ** 2 + rc * rc)
        return np.concatenate((-(a * u + b * x + c + nrm), [x + two_delta * u - klo, khi - x - two_delta * u]))

    def jac(z):
        u, x = z
        nrm = np.sqrt((ru * u) ** 2 + (rx * x) ** 2 + rc * rc)
        J = np.empty((len(a) + 2, 2))
        J[:-2, 0] = -(a + ru * ru * u / nrm)
        J[:-2, 1] = -(b + rx * rx * x / nrm)
        J[-2] = [two_delta, 1.0]
        J[-1] = [-two_delta, -1.0]
        return J

    res = minimize(lambda z: -sign * z[1], x0, jac=lambda z: np.array([0.0, -sign]), method="SLSQP",
                   constraints=[{"type": "ineq", "fun": cons, "jac": jac}], bounds=[(None, None), (box_lo, box_hi)],
                   options={"ftol": 1e-15, "maxiter": 500})
    if not res.success or np.min(cons(res.x)) < -1e-9:
        return None
    return float(res.x[1])
