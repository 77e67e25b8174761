"""Robust (conic) TOPP-RA — BASELINE config 4.  Parity with the reference is UNPINNED for this path (the reference
needs ECOS, absent here); the tests follow SURVEY.md §8c: (i) zero ellipsoid == linear Seidel result,
(ii) feasibility residuals of the returned parameterisation, (iii) monotonicity in the ellipsoid size, plus the
reference's own checks for this path (coefficients: tests/tests/constraint/test_robust_can_linear.py:21-56;
sanity: tests/tests/retime/test_retime_wconic_constraints.py:31-48).  GPU == CPU restatement bit-for-bit."""
import numpy as np
import pytest

from oracle import oracle as orc
from problems import make_batch, make_path

ELL = [1e-3, 5e-2, 9e-3]   # defaults of examples/plot_robust_kinematics.py:26-28


def _rows(seed, G=100, vel_active=False, interp=True):
    ss = np.linspace(0, 1, 5)
    grid = np.linspace(0, 1, G)
    way, vlim, alim = make_path(seed, vel_active=vel_active)
    c = orc.cubic_spline_fit(ss, way)
    lin = orc.solve_velacc(c, ss, grid, vlim, alim, interp, 0, 0, want_rows=True)
    return grid, lin


def _residuals(rows, grid, sd, u, ell):
    """max over stages/rows of  a u + b x + c + ||diag(ell) [u, x, 1]||  at the returned (u_i, x_i)."""
    x = sd[:-1] ** 2
    a, b, c = rows[:-1, 0], rows[:-1, 1], rows[:-1, 2]
    norm = np.sqrt((ell[0] * u) ** 2 + (ell[1] * x) ** 2 + ell[2] ** 2)
    return np.max(a * u[:, None] + b * x[:, None] + c + norm[:, None])


@pytest.mark.parametrize("seed", [1000, 1003, 1007])
def test_oracle_zero_ellipsoid_equals_linear(seed):
    grid, lin = _rows(seed)
    z = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, 28, [0, 0, 0])
    assert z["status"] == 0 == lin["status"]
    np.testing.assert_allclose(z["K"], lin["K"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(z["sd"], lin["sd"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("vel_active", [False, True])
def test_oracle_feasibility_and_monotonicity(vel_active):
    for seed in (1000, 1001, 1002, 1003):
        grid, lin = _rows(seed, vel_active=vel_active)
        prev_K, prev_T = lin["K"][:, 1], None
        for scale in (0.0, 0.5, 1.0, 2.0):
            ell = [scale * e for e in ELL]
            r = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, 28, ell)
            assert r["status"] == 0
            assert _residuals(lin["rows"], grid, r["sd"], r["u"], ell) <= 1e-9
            assert np.all(r["K"][:, 1] <= prev_K + 1e-12)          # controllable sets shrink with the ellipsoid
            T = np.sum(2 * np.diff(grid) / (r["sd"][1:] + r["sd"][:-1]))
            assert prev_T is None or T >= prev_T - 1e-12             # robust trajectories are slower
            prev_K, prev_T = r["K"][:, 1], T
            x = r["sd"] ** 2
            assert np.all(x <= r["K"][:, 1] * (1 + 1e-12) + 1e-14) and np.all(x >= r["K"][:, 0] - 1e-14)


def test_oracle_infeasible_rows_fail_uncontrollable():
    # the reference example draws alim = rand*2 without the +10: joints with alim < rc make the rows infeasible
    grid, lin = _rows(1000)
    rows = lin["rows"].copy()
    rows[:, 2, :] = -1e-4   # c = -amax = -1e-4 > -rc
    r = orc.solve_rows_robust(rows, lin["xbound"], grid, 0, 28, ELL)
    assert r["status"] == 3 and np.isnan(r["sd"]).all()


# ---- GPU ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ta():
    import toppra_b200
    return toppra_b200


@pytest.mark.gpu
@pytest.mark.parametrize("vel_active,scheme", [(False, 1), (True, 1), (False, 0)])
def test_gpu_matches_cpu_restatement_bit_for_bit(ta, vel_active, scheme):
    B, G = 48, 120
    ss, way, vlim, alim = make_batch(B, 3000, vel_active=vel_active)   # cfg 4 seeds 3000+b
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    base = ta.constraint.JointAccelerationConstraint(alim)
    cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.RobustLinearConstraint(base, ELL, scheme)]
    inst = ta.BatchTOPPRA(cons, path, grid)
    res = inst.compute_parameterization(0.0, 0.0, counters=True)
    h = res.to_host()
    R = inst.R
    assert R == (28 if scheme else 14) and inst.conic[:2] == (0, R)
    rec = inst.records.cpu().numpy()
    c_gpu = path.d_ppoly.cpu().numpy()
    for b in range(B):
        lin = orc.solve_velacc(c_gpu[b], ss, grid, vlim[b], alim[b], bool(scheme), 0, 0, want_rows=True)
        assert np.array_equal(rec[b, :, :3 * R].reshape(G, 3, R), lin["rows"])
        o = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, R, ELL)
        assert h["status"][b] == o["status"]
        assert np.array_equal(h["K"][b], o["K"], equal_nan=True)
        assert np.array_equal(h["sd"][b], o["sd"], equal_nan=True) and np.array_equal(h["sdd"][b], o["u"], equal_nan=True)
    assert int(res.counters[:, 0].min()) > 0


@pytest.mark.gpu
def test_gpu_zero_ellipsoid_equals_lp_path(ta):
    B, G = 32, 100
    ss, way, vlim, alim = make_batch(B, 3000)
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    vel, acc = ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)
    lin = ta.BatchTOPPRA([vel, acc], path, grid).compute_parameterization(0, 0).to_host()
    rob = ta.BatchTOPPRA([vel, ta.constraint.RobustLinearConstraint(acc, [0, 0, 0], 1)], path, grid
                         ).compute_parameterization(0, 0).to_host()
    assert not rob["status"].any()
    np.testing.assert_allclose(rob["K"], lin["K"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rob["sd"], lin["sd"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dist_scheme", [0, 1])
def test_gpu_robust_coefficients(ta, dist_scheme):
    """tests/tests/constraint/test_robust_can_linear.py:21-50."""
    dof = 5
    np.random.seed(0)
    alim_ = np.random.rand(5)
    alim = np.vstack((-alim_, alim_)).T
    cnst = ta.constraint.JointAccelerationConstraint(alim)
    np.random.seed(0)
    path = ta.SplineInterpolator(np.linspace(0, 1, 5), np.random.randn(5, dof))
    ro = ta.constraint.RobustLinearConstraint(cnst, [0.1, 2, .3], dist_scheme)
    assert ro.get_constraint_type() == ta.constraint.ConstraintType.CanonicalConic and ro.get_dof() == 5
    grid = np.linspace(0, path.duration, 10)
    a, b, c, P, _, _ = ro.compute_constraint_params(path, grid)
    cnst.set_discretization_type(dist_scheme)
    a0, b0, c0, F0, g0, _, _ = cnst.compute_constraint_params(path, grid)
    for i in range(10):
        np.testing.assert_allclose(a[i], F0.dot(a0[i]))
        np.testing.assert_allclose(b[i], F0.dot(b0[i]))
        np.testing.assert_allclose(c[i], F0.dot(c0[i]) - g0)
        for j in range(a0.shape[1]):
            np.testing.assert_allclose(P[i, j], np.diag([0.1, 2, .3]))
    with pytest.raises(ValueError):
        ta.constraint.RobustLinearConstraint(cnst, [-0.1, 2, .3])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2])
def test_gpu_toppra_conic_api(ta, seed):
    """tests/tests/retime/test_retime_wconic_constraints.py:31-48 with solver_wrapper='ecos' mapped to K2r."""
    vlims = np.array([[-1, 1], [-1, 2], [-1, 4]], dtype=float)
    alims = np.array([[-1, 1], [-1, 2], [-1, 4]], dtype=float)
    vel_c = ta.constraint.JointVelocityConstraint(vlims)
    acc_c = ta.constraint.JointAccelerationConstraint(alims, 0)
    ro_acc_c = ta.constraint.RobustLinearConstraint(acc_c, [1e-4, 1e-4, 5e-4], 0)
    np.random.seed(seed)
    path = ta.SplineInterpolator(np.linspace(0, 1, 5), np.random.randn(5, 3))
    acc_c.set_discretization_type(1)
    ro_acc_c.set_discretization_type(1)
    inst = ta.algorithm.TOPPRA([vel_c, ro_acc_c], path, solver_wrapper="ecos")
    X = inst.compute_feasible_sets()
    assert np.all(X >= 0) and not np.any(np.isnan(X))
    K = inst.compute_controllable_sets(0, 0)
    assert np.all(K >= 0) and not np.any(np.isnan(K))
    traj = inst.compute_trajectory(0, 0)
    assert traj is not None and 0 < traj.duration < 20
    with pytest.raises(AssertionError):
        ta.algorithm.TOPPRA([vel_c, ro_acc_c], path, solver_wrapper="seidel")   # reference :78-84


# ---- optimality: independent certificate + unrelated solver ------------------------------------
def _cfg4_problem(seed, G=200):
    """cfg 4: limits of cfg 2, RobustLinearConstraint(acc-interp, ELL, Interpolation); rows/xbound from the linear oracle."""
    ss = np.linspace(0, 1, 5)
    grid = np.linspace(0, 1, G)
    way, vlim, alim = make_path(seed)
    lin = orc.solve_velacc(orc.cubic_spline_fit(ss, way), ss, grid, vlim, alim, True, 0, 0, want_rows=True)
    return grid, lin


def test_oracle_controllable_sets_are_extremal():
    """Every K endpoint of the restatement is feasible and either sits on the x box or is infeasible one part in 1e9
    further out, evaluated by tests/robust_check.py which does not know the closed form (8 cfg-4 paths x 199 stages)."""
    from robust_check import certify_controllable_sets
    n_ext = 0
    for seed in range(3000, 3008):
        grid, lin = _cfg4_problem(seed)
        r = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, 28, ELL)
        assert r["status"] == 0
        _, ext, _ = certify_controllable_sets(lin["rows"], lin["xbound"], grid, r["K"], ELL)
        n_ext += ext
    assert n_ext >= 8 * 199  # every upper end is an interior extremal point (the lower ends sit on x = 0 here)
    # a non-zero terminal velocity lifts the lower ends off the box near the end of the path: min-x certified too
    n_low = 0
    for seed in range(3000, 3004):
        grid, lin = _cfg4_problem(seed)
        r = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, 28, ELL, 0.0, 0.1)
        assert r["status"] == 0 and abs(r["K"][-1, 0] - 0.01) < 1e-15
        _, ext, _ = certify_controllable_sets(lin["rows"], lin["xbound"], grid, r["K"], ELL)
        n_low += ext - 199
    assert n_low > 0


def test_oracle_vs_slsqp_on_sampled_stage_problems():
    """>= 200 stage SOCPs (max x and min x) re-solved with scipy SLSQP on the cone form of
    ecos_solverwrapper.py:112-188: agreement to 1e-7 (relative + absolute), the tolerance DESIGN.md states for cfg 4."""
    from robust_check import slsqp_extreme_x, ECOS_INFTY, ECOS_MAXX
    rng = np.random.RandomState(0)
    done = 0
    for seed in range(3000, 3012):
        grid, lin = _cfg4_problem(seed)
        r = orc.solve_rows_robust(lin["rows"], lin["xbound"], grid, 0, 28, ELL)
        K, rows, xb = r["K"], lin["rows"], lin["xbound"]
        for i in rng.choice(len(grid) - 1, size=20, replace=False):
            a, b, c = rows[i, 0], rows[i, 1], rows[i, 2]
            td = 2 * (grid[i + 1] - grid[i])
            lo, hi = max(xb[i, 0], -ECOS_INFTY), min(xb[i, 1], ECOS_MAXX, ECOS_INFTY)
            xmid = 0.5 * (K[i, 0] + K[i, 1])
            umid = (0.5 * (K[i + 1, 0] + K[i + 1, 1]) - xmid) / td
            for sign, ref in ((1, K[i, 1]), (-1, K[i, 0])):
                xs = slsqp_extreme_x(a, b, c, ELL, td, K[i + 1, 0], K[i + 1, 1], lo, hi, sign, np.array([umid, xmid]))
                if xs is None:
                    continue
                assert abs(max(xs, 0.0) - ref) <= 1e-7 * (1 + abs(ref)), (seed, i, sign, xs, ref)
                done += 1
    assert done >= 200


@pytest.mark.gpu
def test_gpu_controllable_sets_are_extremal(ta):
    """The same certificate on the KERNEL's output: 64 cfg-4 paths (seeds 3000+b), every stage."""
    from robust_check import certify_controllable_sets
    B, G = 64, 200
    ss, way, vlim, alim = make_batch(B, 3000)
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    base = ta.constraint.JointAccelerationConstraint(alim)
    cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.RobustLinearConstraint(base, ELL, 1)]
    inst = ta.BatchTOPPRA(cons, path, grid)
    h = inst.compute_parameterization(0.0, 0.0).to_host()
    assert not h["status"].any()
    R = inst.R
    rec = inst.records.cpu().numpy()
    n_ext = 0
    for b in range(B):
        rows = rec[b, :, :3 * R].reshape(G, 3, R)
        _, ext, _ = certify_controllable_sets(rows, rec[b, :, 3 * R:3 * R + 2], grid, h["K"][b], ELL)
        n_ext += ext
    assert n_ext >= B * (G - 1)
