"""Batched device forms of the callers around the scan and `ubound` from a constraint
(row a7, cy_seidel_solverwrapper.pyx:512-515), against tests/golden/frows_batch.npz = outputs of the UNMODIFIED reference
for 16 paths each (tests/golden/make_golden.py frows_batch):

  * `BatchSplineInterpolator.propose_gridpoints`  vs  interpolator.propose_gridpoints (ragged grids, bit-exact)
  * `BatchTOPPRA(gridpoints=None | grid, glen)`   vs  TOPPRA on each path's own grid (tb_scan*_ragged, bit-exact)
  * `BatchTOPPRA.compute_reachable_sets`          vs  compute_reachable_sets (tb_reachable_sets, bit-exact)
  * `BatchTOPPRAsd`                               vs  TOPPRAsd (tb_sd_bisect, bit-exact)
  * `BatchParametrizeSpline`                      vs  ParametrizeSpline (knots bit-exact; evaluations 1e-9)
  * a user-defined LinearConstraint with a ubound vs  the reference (TB_SCAN_UBOUND records, bit-exact)

Every test runs twice: on the GPU (-m gpu) and, under -m "not gpu", on the oracle-backed engine double
(tests/cpu_engine.py), which pins the oracle and the Python host logic to the same reference outputs."""
import numpy as np
import pytest

import cpu_engine


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "cpu_double"])
def ta(request, monkeypatch):
    if request.param == "cpu_double":
        return cpu_engine.install(monkeypatch)
    import toppra_b200
    return toppra_b200


def _same(a, b):
    """array_equal with NaN == NaN."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _cons(ta, g, sl=slice(None)):
    return [ta.constraint.JointVelocityConstraint(g["vlim"][sl]), ta.constraint.JointAccelerationConstraint(g["alim"][sl])]


def test_propose_gridpoints_batch_ragged_bit_exact(ta, golden):
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    for tag, kw in (("pg_default", {}), ("pg_toppra", dict(max_err_threshold=1e-3, min_nb_points=100)),
                    ("pg_coarse", dict(max_err_threshold=5e-2, max_seg_length=0.3, min_nb_points=20))):
        grid, glen = path.propose_gridpoints(**kw)
        grid, glen = grid.cpu().numpy(), glen.cpu().numpy()
        assert np.array_equal(glen, g[tag + "_len"]), tag
        assert grid.shape[1] == glen.max()
        for b in range(len(glen)):
            assert np.array_equal(grid[b, :glen[b]], g[tag + "_grid"][b, :glen[b]]), (tag, b)
            assert np.all(grid[b, glen[b]:] == 1.0)          # padded with the end of the path
    # the single-path function is the B = 1 case
    one = ta.propose_gridpoints(ta.SplineInterpolator(g["ss"], g["way"][3]), max_err_threshold=1e-3, min_nb_points=100)
    assert isinstance(one, list) and np.array_equal(one, g["pg_toppra_grid"][3, :g["pg_toppra_len"][3]])
    with pytest.raises(ValueError, match="Unable to find a good gridpoint"):
        path.propose_gridpoints(max_err_threshold=1e-12, max_iteration=3)


@pytest.mark.parametrize("fused", [True, False])
def test_ragged_batch_solve_bit_exact(ta, golden, fused):
    """gridpoints=None: every path solved on ITS proposed grid in one launch (fused vel+acc scan and record scan)."""
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    inst = ta.BatchTOPPRA(_cons(ta, g), path, gridpoints=None, fused=fused)
    assert inst.fused == fused and np.array_equal(inst.glen.cpu().numpy(), g["pg_toppra_len"])
    h = inst.compute_parameterization(0.0, 0.0).to_host()
    assert np.array_equal(h["status"], g["ragged_status"]) and not h["status"].any()
    assert _same(h["K"], g["ragged_K"]) and _same(h["sd"], g["ragged_sd"]) and _same(h["sdd"], g["ragged_sdd"])
    # explicit ragged grids + validation of the REAL ends only
    grid = np.where(np.isnan(g["pg_toppra_grid"]), 1.0, g["pg_toppra_grid"])
    inst2 = ta.BatchTOPPRA(_cons(ta, g), path, grid, glen=g["pg_toppra_len"], fused=fused)
    h2 = inst2.compute_parameterization(0.0, 0.0).to_host()
    assert _same(h2["sd"], g["ragged_sd"])
    bad = grid.copy()
    bad[2, 5] = bad[2, 4]
    with pytest.raises(ValueError, match="Bad input gridpoints"):
        ta.BatchTOPPRA(_cons(ta, g), path, bad, glen=g["pg_toppra_len"])


def test_ragged_batch_in_chunks_bit_exact(ta, golden):
    """Ragged grids through a record buffer that holds only 5 of the 16 paths: chunks carry their own slice of the grid
    lengths; results equal the reference's golden (and hence the one-launch solve) bit for bit."""
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    whole = ta.BatchTOPPRA(_cons(ta, g), path, gridpoints=None, fused=False)
    rows = sum(c.num_rows(whole.ctx) for c in whole.constraints)
    per_path = 8 * ta.engine.record_doubles(rows) * whole.G
    inst = ta.BatchTOPPRA(_cons(ta, g), path, gridpoints=None, fused=False, max_record_bytes=5 * per_path)
    assert inst.chunk_size() == 5 and whole.chunk_size() == 16
    h = inst.compute_parameterization(0.0, 0.0).to_host()
    assert not h["status"].any()
    assert _same(h["K"], g["ragged_K"]) and _same(h["sd"], g["ragged_sd"]) and _same(h["sdd"], g["ragged_sdd"])


def test_reachable_sets_batch_bit_exact(ta, golden):
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    inst = ta.BatchTOPPRA(_cons(ta, g), path, g["grid"])
    L, X, fail = inst.compute_reachable_sets(g["sdmin"], g["sdmax"])
    assert np.array_equal(X.cpu().numpy(), g["X"])
    assert _same(L.cpu().numpy(), g["L"])
    nan_rows = np.isnan(g["L"]).any(axis=2)               # some paths cannot start at sdmin: "Path not parametrizable"
    first_nan = np.where(nan_rows.any(axis=1), nan_rows.argmax(axis=1), -1)
    assert np.array_equal(fail.cpu().numpy(), first_nan) and (first_nan >= 0).sum() >= 2
    for b in np.nonzero(first_nan >= 0)[0]:
        assert not g["L"][b, first_nan[b] + 1:].any()     # rows after the failure stay 0 (np.zeros)
    # single-path API = the B = 1 case (and it records X like the reference)
    b = 5
    one = ta.algorithm.TOPPRA(_cons(ta, g, b), ta.SplineInterpolator(g["ss"], g["way"][b]), gridpoints=g["grid"])
    assert _same(one.compute_reachable_sets(g["sdmin"][b], g["sdmax"][b]), g["L"][b])
    assert np.array_equal(one.problem_data.X, g["X"][b])


def test_toppra_sd_batch_bit_exact(ta, golden):
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    inst = ta.BatchTOPPRAsd(_cons(ta, g), path, g["grid"])
    inst.set_desired_duration(g["sd_desired"])
    res = inst.compute_parameterization(0.0, 0.0)
    h = res.to_host()
    assert np.array_equal(h["status"], g["sd_status"])
    assert np.array_equal(h["sd"], g["sd_sd"]) and np.array_equal(h["sdd"], g["sd_sdd"])
    alpha = res.alpha.cpu().numpy()
    assert (alpha[0::4] == 1.0).all()                                     # shorter than the fastest: unachievable
    assert ((alpha[1::4] > 0) & (alpha[1::4] < 1)).all() and (alpha[3::4] < 1e-9).all()   # 1e6 x fastest: almost slowest
    inst.set_desired_duration(1e30)                                       # longer than the slowest: unachievable
    assert (inst.compute_parameterization(0.0, 0.0).alpha.cpu().numpy() == 0.0).all()
    # durations of the blends hit the target within the reference's atol
    dur = np.sum(2 * np.diff(g["grid"]) / (h["sd"][:, 1:] + h["sd"][:, :-1] + 1e-9), axis=1)
    inside = np.arange(16) % 4 == 1
    assert np.all(np.abs(dur[inside] - g["sd_desired"][inside]) <= 1e-5 + 1e-12)
    # single-path class = the B = 1 case
    b = 9
    one = ta.algorithm.TOPPRAsd(_cons(ta, g, b), ta.SplineInterpolator(g["ss"], g["way"][b]), gridpoints=g["grid"])
    one.set_desired_duration(g["sd_desired"][b])
    sdd, sd, _, _ = one.compute_parameterization(0, 0, return_data=True)
    assert np.array_equal(sd, g["sd_sd"][b]) and np.array_equal(sdd, g["sd_sdd"][b])


def test_parametrize_spline_batch(ta, golden):
    g = golden("frows_batch")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    traj = ta.BatchParametrizeSpline(path, g["grid"], g["ps_vel"])
    assert np.array_equal(traj.nkeep.cpu().numpy(), g["ps_n"])                 # path 6 drops three knots
    t = traj.t_knots.cpu().numpy()
    for b in range(16):
        assert np.array_equal(t[b, :g["ps_n"][b]], g["ps_t"][b, :g["ps_n"][b]]), b
    assert np.array_equal(traj.durations.cpu().numpy(), g["ps_dur"])
    assert len(traj.groups) == 2
    ts = g["ps_ts"][None, :] * g["ps_dur"][:, None]
    for order, key in ((0, "ps_q"), (1, "ps_qd"), (2, "ps_qdd")):
        got = traj(ts, order).cpu().numpy()
        np.testing.assert_allclose(got, g[key], rtol=1e-9, atol=1e-9 * max(1.0, np.abs(g[key]).max()))
    # single-path class = the B = 1 case
    b = 6
    one = ta.ParametrizeSpline(ta.SplineInterpolator(g["ss"], g["way"][b]), g["grid"], g["ps_vel"][b])
    assert np.array_equal(one.ss_waypoints, g["ps_t"][b, :g["ps_n"][b]])


def test_ubound_from_a_constraint_bit_exact(ta, golden):
    """A LinearConstraint subclass that returns ubound (and xbound): seidelWrapper.__init__ intersects it into
    low/high[:, 0] (pyx:512-515); here the stage records carry the pair and the kernels take TB_SCAN_UBOUND."""
    g = golden("frows_batch")
    grid = g["grid"]

    class UBoundConstraint(ta.constraint.LinearConstraint):
        def __init__(self, acc, ulim):
            super(UBoundConstraint, self).__init__()
            self.acc, self.ulim = acc, ulim
            self.discretization_type = acc.discretization_type
            self.identical = True

        def get_dof(self):
            return self.acc.get_dof()

        def compute_constraint_params(self, path, gridpoints, *a):
            pa, pb, pc, F, gg, _, _ = self.acc.compute_constraint_params(path, gridpoints)
            n = len(gridpoints)
            ub = np.stack((-self.ulim * (1.0 + gridpoints), self.ulim * (2.0 - gridpoints)), axis=1)
            xb = np.stack((np.zeros(n), 40.0 + 30 * gridpoints), axis=1)
            return pa, pb, pc, F, gg, ub, xb

    tight = 0
    for b in range(8):
        path = ta.SplineInterpolator(g["ss"], g["way"][b])
        mk = lambda: [ta.constraint.JointVelocityConstraint(g["vlim"][b]),  # noqa: E731
                      UBoundConstraint(ta.constraint.JointAccelerationConstraint(g["alim"][b]), g["ub_ulim"][b])]
        inst = ta.algorithm.TOPPRA(mk(), path, gridpoints=grid, solver_wrapper="seidel")
        assert ta.engine.has_ubound(inst.solver_wrapper.records, inst.solver_wrapper.R)
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert np.array_equal(K, g["ub_K"][b]) and np.array_equal(sd, g["ub_sd"][b]) and np.array_equal(sdd, g["ub_sdd"][b])
        assert list(ta.algorithm.ParameterizationReturnCode).index(inst.problem_data.return_code) == g["ub_status"][b]
        hi = g["ub_ulim"][b] * (2.0 - grid[:-1])
        tight += int(np.sum(np.abs(sdd - hi) < 1e-6) + np.sum(np.abs(sdd + g["ub_ulim"][b] * (1.0 + grid[:-1])) < 1e-6))
        assert np.array_equal(ta.algorithm.TOPPRA(mk(), path, gridpoints=grid).compute_feasible_sets(), g["ub_X"][b])
        assert _same(ta.algorithm.TOPPRA(mk(), path, gridpoints=grid).compute_reachable_sets(0.0, 0.3), g["ub_L"][b])
        # the per-stage plugin interface sees the same bounds
        w = inst.solver_wrapper
        rows = w.rows()
        assert np.array_equal(rows["low"][:, 0], np.maximum(-1e8, -g["ub_ulim"][b] * (1.0 + grid)))
    assert tight > 50       # the u-bound is active on many stages: the test exercises it


@pytest.mark.parametrize("n", [2, 3, 4, 5, 9, 20])
def test_periodic_spline_bit_exact(ta, golden, n):
    """bc_type='periodic' (scipy CubicSpline; reference SplineInterpolator forwards bc_type, interpolator.py:419): the
    condensed cyclic system of scipy _cubic.py restated in K0; coefficients and evaluations equal scipy's bit for bit."""
    g = golden("spline_periodic")
    x, y = g["x_%d" % n], g["y_%d" % n]
    path = ta.SplineInterpolator(x, y, bc_type="periodic")
    assert np.array_equal(path.cspl.c, g["c_%d" % n])
    for order, key in ((0, "q"), (1, "qd"), (2, "qdd")):
        assert np.array_equal(path(g["s_%d" % n], order), g["%s_%d" % (key, n)]), (n, order)
    if n > 2:
        bad = y.copy()
        bad[-1, 0] += 1e-3
        with pytest.raises(ValueError, match="identical"):
            ta.SplineInterpolator(x, bad, bc_type="periodic")
    with pytest.raises(ValueError, match="both"):
        ta.SplineInterpolator(x, y, bc_type=("periodic", "natural"))


def test_periodic_path_solve(ta, golden):
    """TOPPRA on a closed path.  The kernels evaluate the LAST gridpoint on the last spline segment, the reference (through
    scipy's periodic extrapolation) on the first one at ds = 0: the same number up to rounding, so the parameterisation
    agrees to 1e-9 instead of bit for bit (documented in DESIGN.md)."""
    g = golden("spline_periodic")
    path = ta.SplineInterpolator(g["solve_ss"], g["solve_way"], bc_type="periodic")
    cons = [ta.constraint.JointVelocityConstraint(g["solve_vlim"]), ta.constraint.JointAccelerationConstraint(g["solve_alim"])]
    inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["solve_grid"], solver_wrapper="seidel")
    sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
    assert g["solve_status"] == 0
    np.testing.assert_allclose(K, g["solve_K"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sd, g["solve_sd"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sdd, g["solve_sdd"], rtol=1e-8, atol=1e-8)
