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
    chunked = ta.BatchTOPPRA(_cons(ta, g), bp2, g["grid"], max_record_bytes=5 * len(g["grid"]) * W * 8)
    h2 = chunked.compute_parameterization(0.0, 0.0).to_host()
    assert np.array_equal(h2["K"], g["K"]) and np.array_equal(h2["sd"], g["sd"])
    for bad in (g["c"][0], np.zeros((2, 5, 4, 7))):
        with pytest.raises(ValueError):
            ta.BatchSplineInterpolator.from_ppoly(g["ss"], bad)
    with pytest.raises(ValueError):
        ta.BatchSplineInterpolator.from_ppoly(np.linspace(0, 1, 4), g["c"])
