"""GPU (-m gpu; replayed on CPU by tests/test_host_pipeline_cpu.py): "PPoly in"  — paths handed over as
piecewise cubics in scipy's PPoly layout instead of waypoints: `PPolyPath` for the single-path API and
`BatchSplineInterpolator.from_ppoly` for batches.  Fed with the coefficients of the reference's own spline fits, the
solver must reproduce the reference's golden parameterisations bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ta():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import toppra_b200
    return toppra_b200


def _cons(ta, g, b=None):
    vlim = g["vlim"] if b is None else g["vlim"][b]
    alim = g["alim"] if b is None else g["alim"][b]
    return [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)]


def test_ppoly_path_single(ta, golden):
    from scipy.interpolate import PPoly
    g = golden("cfg1_seed9")
    spline = ta.SplineInterpolator(g["ss"], g["way"])
    s = np.linspace(0, 1, 33)
    for path in (ta.PPolyPath(g["c"], g["ss"]), ta.PPolyPath(PPoly(g["c"], g["ss"]))):
        assert path.dof == 7 and path.path_interval.tolist() == [0.0, 1.0] and path.duration == 1.0
        for order in (0, 1, 2):
            assert np.array_equal(path(s, order), spline(s, order))
        assert path(0.3).shape == (7,) and np.array_equal(path.eval(s), path(s)) and np.array_equal(path.evaldd(s), path(s, 2))
        np.testing.assert_allclose(path.waypoints[1], g["way"], rtol=0, atol=1e-14)
        inst = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"], solver_wrapper="seidel")
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert np.array_equal(K, g["K"]) and np.array_equal(sd, g["sd"]) and np.array_equal(sdd, g["sdd"])
    traj = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"]).compute_trajectory(0, 0)
    assert abs(traj.duration - float(g["traj_duration"])) < 1e-9
    ca = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"], parametrizer="ParametrizeConstAccel")
    assert ca.compute_trajectory(0, 0).duration == float(g["ca_duration"])
    with pytest.raises(ValueError):
        path(0.5, 3)
    with pytest.raises(ValueError):
        ta.PPolyPath(np.zeros((5, 3, 2)), np.linspace(0, 1, 4))          # quartic pieces
    with pytest.raises(ValueError):
        ta.PPolyPath(np.zeros((4, 3, 2)), np.array([0.0, 0.5, 0.4, 1.0]))   # breaks not increasing


def test_scalar_and_low_degree_pieces(ta):
    """(k, nseg) coefficients = a 1-DOF path; k < 4 is zero-padded: a parabola q(s) = 1 + 2 s - 3 s^2 on two pieces."""
    x = np.array([0.0, 0.4, 1.0])
    c = np.array([[-3.0, -3.0], [2.0, 2.0 - 6 * 0.4], [1.0, 1.0 + 2 * 0.4 - 3 * 0.16]])
    path = ta.PPolyPath(c, x)
    s = np.linspace(0, 1, 11)
    assert path.dof == 1 and path(s).shape == (11,)
    np.testing.assert_allclose(path(s), 1 + 2 * s - 3 * s ** 2, rtol=0, atol=1e-15)
    np.testing.assert_allclose(path(s, 1), 2 - 6 * s, rtol=0, atol=1e-15)
    np.testing.assert_allclose(path(s, 2), -6 * np.ones(11), rtol=0, atol=0)
    cubic = ta.PPolyPath(np.concatenate((np.zeros((1, 2)), c)), x)
    assert np.array_equal(cubic(s, 1), path(s, 1))


def test_batch_from_ppoly(ta, golden):
    g = golden("cfg2_seeds1000")
    bpath = ta.BatchSplineInterpolator.from_ppoly(g["ss"], g["c"])
    assert (bpath.B, bpath.n, bpath.dof, bpath.nseg) == (16, 5, 7, 4)
    np.testing.assert_allclose(bpath.d_wp.cpu().numpy(), g["way"], rtol=0, atol=1e-14)
    h = ta.BatchTOPPRA(_cons(ta, g), bpath, g["grid"]).compute_parameterization(0.0, 0.0).to_host()
    assert np.array_equal(h["status"], g["status"])
    assert np.array_equal(h["K"], g["K"]) and np.array_equal(h["sd"], g["sd"]) and np.array_equal(h["sdd"], g["sdd"])
    # per-path breaks, chunked solve
    bp2 = ta.BatchSplineInterpolator.from_ppoly(np.tile(g["ss"], (16, 1)), g["c"])
    W = ta.engine.record_doubles(28)
    chunked = ta.BatchTOPPRA(_cons(ta, g), bp2, g["grid"], max_record_bytes=5 * len(g["grid"]) * W * 8, fused=False)
    h2 = chunked.compute_parameterization(0.0, 0.0).to_host()
    assert np.array_equal(h2["K"], g["K"]) and np.array_equal(h2["sd"], g["sd"])
    for bad in (g["c"][0], np.zeros((2, 5, 4, 7))):
        with pytest.raises(ValueError):
            ta.BatchSplineInterpolator.from_ppoly(g["ss"], bad)
    with pytest.raises(ValueError):
        ta.BatchSplineInterpolator.from_ppoly(np.linspace(0, 1, 4), g["c"])


def test_simple_path_and_polynomial_path(ta, golden):
    """SimplePath (reference simplepath.py, cubic Hermite) and PolynomialPath (interpolator.py:584-686, degree <= 3 here)
    against the reference's own evaluations and parameterisations.  The reference evaluates Bernstein / power-series
    forms, this package local cubics: tolerance 1e-12 on values (they are O(1..30)), 1e-9 relative on the solution."""
    g = golden("other_paths")
    vel, acc = ta.constraint.JointVelocityConstraint(g["vlim"]), ta.constraint.JointAccelerationConstraint(g["alim"])
    paths = (("sp_auto", ta.SimplePath(g["x"], g["y"])), ("sp_yd", ta.SimplePath(g["x"], g["y"], g["yd"])),
             ("poly", ta.PolynomialPath([[1, 2, 3], [-2, 3, 4, 5], [0.5, -1.0]], s_start=0.0, s_end=2.5)))
    for tag, path in paths:
        assert path.dof == 3 and path.path_interval.tolist() == [0.0, 2.5] and path.duration == 2.5
        for order in (0, 1, 2):
            np.testing.assert_allclose(path(g["s"], order), g["%s_q%d" % (tag, order)], rtol=1e-12, atol=1e-12)
        inst = ta.algorithm.TOPPRA([vel, acc], path, gridpoints=g["grid"], solver_wrapper="seidel")
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert inst.problem_data.return_code == ta.algorithm.STATUS_CODES[int(g[tag + "_status"])]
        np.testing.assert_allclose(K, g[tag + "_K"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(sd, g[tag + "_sd"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(sdd, g[tag + "_sdd"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(paths[0][1].waypoints, g["y"])
    # the reference's own unit tests (tests/tests/interpolators/test_simple_path.py, test_poly_interpolator.py)
    f = ta.SimplePath([0, 1, 2], np.array([0, 1, 1]))
    assert f(1) == 1.0 and f(2) == 1.0 and f(np.linspace(0, 2, 200), 1).shape == (200, 1)
    np.testing.assert_allclose(f(1, 1), 0.5)
    np.testing.assert_allclose([f(0, 1), f(2, 1)], 0, atol=1e-15)
    fd = ta.SimplePath([0, 1, 2], np.array([0, 1, 1]), np.array([0, 2, 0]))
    np.testing.assert_allclose([fd(0, 1)[0], fd(1, 1)[0], fd(2, 1)[0]], [0, 2.0, 0], atol=1e-15)
    fm = ta.SimplePath([0, 1, 2], np.array([[0, 0], [1, 2], [1, 2]]))
    assert fm(0.5).shape == (2,) and fm.dof == 2
    np.testing.assert_allclose(fm(1), [1, 2])
    pi = ta.PolynomialPath([1, 2, 3], s_start=0, s_end=2)
    assert pi.dof == 1
    np.testing.assert_allclose(pi.eval([0, 0.5, 1]), [1, 2.75, 6])
    np.testing.assert_allclose(pi.evald([0, 0.5, 1]), [2, 5, 8])
    np.testing.assert_allclose(pi.evaldd([0, 0.5, 1]), [6, 6, 6])
    np.testing.assert_allclose(pi.path_interval, [0, 2])
    p2 = ta.PolynomialPath([[1, 2, 3], [-2, 3, 4, 5]])
    np.testing.assert_allclose(p2.eval([0, 0.5, 1]), [[1, -2], [2.75, 1.125], [6, 10]])
    np.testing.assert_allclose(p2.evald([0, 0.5, 1]), [[2, 3], [5, 10.75], [8, 26]])
    np.testing.assert_allclose(p2.evaldd([0, 0.5, 1]), [[6, 8], [6, 23], [6, 38]])
    with pytest.raises(NotImplementedError):
        ta.PolynomialPath([1, 0, 0, 0, 1])


def test_univariate_spline_interpolator(ta):
    """Smoothing-spline path (reference interpolator.py:508-581) through "PPoly in": values against scipy's own evaluation
    of the fitted splines (what the reference's __call__ returns) to rounding, the solve against the oracle fed with the
    converted coefficients bit for bit."""
    from scipy.interpolate import UnivariateSpline
    from oracle import oracle as orc
    rng = np.random.RandomState(5)
    ss = np.linspace(0, 2.0, 40)
    way = np.stack([np.sin(2 * ss), np.cos(ss) * ss, 0.3 * ss ** 2], axis=1) + 0.02 * rng.randn(40, 3)
    path = ta.UnivariateSplineInterpolator(ss, way)
    assert path.dof == 3 and list(path.path_interval) == [0.0, 2.0] and path.duration == 2.0
    assert np.array_equal(path.waypoints[1], way)
    s = np.linspace(0, 2.0, 257)
    for k in range(3):
        spl = UnivariateSpline(ss, way[:, k])
        for order in (0, 1, 2):
            expect = spl(s) if order == 0 else spl.derivative(order)(s)
            np.testing.assert_allclose(path(s, order)[:, k], expect, rtol=1e-10, atol=1e-10)
    assert path(0.7).shape == (3,) and path([0.1, 0.2], 1).shape == (2, 3)
    vlim = np.array([[-3.0, 3.0]] * 3)
    alim = np.array([[-8.0, 8.0]] * 3)
    grid = np.linspace(0, 2.0, 201)
    inst = ta.algorithm.TOPPRA([ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)],
                               path, gridpoints=grid, solver_wrapper="seidel")
    sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
    bp = path.as_batch()
    o = orc.solve_velacc(bp.d_ppoly[0].cpu().numpy(), bp.d_ss.cpu().numpy().reshape(-1), grid, vlim, alim, True, 0, 0)
    assert o["status"] == 0 and inst.problem_data.return_code == ta.algorithm.ParameterizationReturnCode.Ok
    assert np.array_equal(K, o["K"]) and np.array_equal(sd, o["sd"])
    scalar = ta.UnivariateSplineInterpolator(ss, way[:, 0])
    assert scalar.dof == 1 and scalar(s).shape == (257, 1)      # np.array(data).T in the reference: always 2-D
    with pytest.raises(AssertionError):
        ta.UnivariateSplineInterpolator(ss + 1.0, way)
