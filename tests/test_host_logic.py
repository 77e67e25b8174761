"""CPU: host-side logic that needs no GPU — argument validation, enums, API surface, numpy utilities."""
import numpy as np
import pytest

import toppra_b200 as ta
from toppra_b200 import constraint
from toppra_b200.algorithm import ParameterizationReturnCode, STATUS_CODES


def test_public_names_match_reference_surface():
    for name in ("SplineInterpolator", "ParametrizeSpline", "ParametrizeConstAccel", "constraint", "algorithm",
                 "solverwrapper", "BatchTOPPRA", "BatchSplineInterpolator"):
        assert hasattr(ta, name)
    for name in ("JointVelocityConstraint", "JointVelocityConstraintVarying", "JointAccelerationConstraint",
                 "SecondOrderConstraint", "JointTorqueConstraint", "RobustLinearConstraint", "LinearConstraint", "Constraint",
                 "ConstraintType", "DiscretizationType", "canlinear_colloc_to_interpolate"):  # reference constraint/__init__.py
        assert hasattr(constraint, name)
    for name in ("TOPPRA", "ParameterizationData", "ParameterizationReturnCode", "ParameterizationAlgorithm"):
        assert hasattr(ta.algorithm, name)
    assert ta.constants.TINY == 1e-8 and ta.constants.SMALL == 1e-5 and ta.constants.MAX_TRIES == 10


def test_status_code_order_matches_header():
    # include/toppra_b200.h TB_STATUS_* <-> reference enum order (algorithm.py:49-56)
    assert [c.name for c in STATUS_CODES] == ["Ok", "ErrUnknown", "ErrShortPath", "FailUncontrollable",
                                              "ErrForwardPassFail"]
    assert str(ParameterizationReturnCode.Ok).startswith("<ParameterizationReturnCode.Ok")


def test_velocity_constraint_validation():
    c = constraint.JointVelocityConstraint([1.0, 2.0])
    assert c.vlim.tolist() == [[-1.0, 1.0], [-2.0, 2.0]] and c.get_dof() == 2
    assert c.get_constraint_type() == constraint.ConstraintType.CanonicalLinear
    with pytest.raises(ValueError):
        constraint.JointVelocityConstraint([[1.0, -1.0], [0, 1]])  # lower >= upper, linear_joint_velocity.py:34-37
    with pytest.raises(ValueError):
        constraint.JointVelocityConstraint([np.nan, 1.0])
    cb = constraint.JointVelocityConstraint(np.tile(np.array([[-1.0, 1.0]]), (5, 3, 1)))
    assert cb.get_dof() == 3 and cb.vlim.shape == (5, 3, 2)


def test_acceleration_constraint_schemes():
    c = constraint.JointAccelerationConstraint([1.0, 2.0, 3.0])
    assert c.get_discretization_type() == constraint.DiscretizationType.Interpolation and c.identical
    c.set_discretization_type(0)
    assert c.get_discretization_type() == constraint.DiscretizationType.Collocation
    c.set_discretization_type(constraint.DiscretizationType.Interpolation)
    assert c.interpolation
    with pytest.raises(NotImplementedError):
        c.set_discretization_type(7)
    assert "Acceleration limit" in repr(c)


def test_canlinear_colloc_to_interpolate_matches_reference(golden):
    """Host utility (API parity) against the reference's lifted parameters (golden acc_a / acc_b)."""
    g = golden("cfg2_seeds1000")
    qs, qss, grid = g["qs"][0], g["qss"][0], g["grid"]
    dof = qs.shape[1]
    F = np.vstack((np.eye(dof), -np.eye(dof)))
    gg = np.r_[g["alim"][0][:, 1], -g["alim"][0][:, 0]]
    a, b, c, F2, g2, _, _ = constraint.canlinear_colloc_to_interpolate(qs, qss, np.zeros_like(qs), F, gg, None, None,
                                                                      grid, identical=True)
    assert np.array_equal(a, g["acc_a"][0]) and np.array_equal(b, g["acc_b"][0])
    assert np.array_equal(F2, g["acc_F"][0]) and np.array_equal(g2, g["acc_g"][0])
    assert constraint.canlinear_colloc_to_interpolate(None, None, None, None, None, None, None, grid)[0] is None


def test_available_solvers():
    av = dict(ta.solverwrapper.available_solvers(output_msg=False))
    assert av["seidel"] and not av["ecos"]
    assert ta.solverwrapper.check_solver_availability("seidel")


def test_parse_bc_errors():
    from toppra_b200 import engine
    import torch
    if not torch.cuda.is_available():
        assert engine.parse_bc("periodic", 1, 2, "cpu") == ((3, None), (3, None))
        with pytest.raises(ValueError, match="both"):      # scipy: 'periodic' is defined for both curve ends
            engine.parse_bc(("periodic", "natural"), 1, 2, "cpu")
        with pytest.raises(ValueError):
            engine.parse_bc("bogus", 1, 2, "cpu")


class _FakePath(object):
    """Stand-in for a path object: polynomial q(s) evaluated in numpy (the constraints' host side only calls it)."""
    dof = 3

    def __call__(self, s, order=0):
        s = np.asarray(s, dtype=float)
        base = np.stack((np.sin(s), s ** 2, 1.0 - s), axis=-1)
        d1 = np.stack((np.cos(s), 2 * s, -np.ones_like(s)), axis=-1)
        d2 = np.stack((-np.sin(s), 2 * np.ones_like(s), np.zeros_like(s)), axis=-1)
        return (base, d1, d2)[order]


def test_joint_torque_constraint_host_tuple(golden):
    """JointTorqueConstraint (reference joint_torque.py:77-116): the host 7-tuple follows the reference's formulas for any
    path object; the golden case compares numbers produced by the reference class itself (tests/test_gpu_parity.py does
    it with the device path evaluation, bit for bit)."""
    inv_dyn = lambda q, qd, qdd: 2.0 * qdd + 0.3 * qd * qd + np.sin(q)  # noqa: E731
    tl = np.array([[-3.0, 4.0], [-5.0, 6.0], [-7.0, 8.0]])
    fric = np.array([0.1, 0.2, 0.3])
    grid = np.linspace(0, 1, 9)
    path = _FakePath()
    c0 = constraint.JointTorqueConstraint(inv_dyn, tl, fric)
    assert c0.identical and c0.get_dof() == 3 and c0.discretization_type == constraint.DiscretizationType.Collocation
    a, b, c, F, g, ub, xb = c0.compute_constraint_params(path, grid)
    q, qd, qdd = path(grid), path(grid, 1), path(grid, 2)
    np.testing.assert_array_equal(c, np.sin(q) + fric * np.sign(qd))
    np.testing.assert_allclose(a, 2.0 * qd, rtol=0, atol=1e-15)
    np.testing.assert_allclose(b, 2.0 * qdd + 0.3 * qd * qd, rtol=0, atol=1e-15)
    assert ub is None and xb is None
    np.testing.assert_array_equal(F, np.vstack((np.eye(3), -np.eye(3))))
    np.testing.assert_array_equal(g, np.r_[tl[:, 1], -tl[:, 0]])
    c1 = constraint.JointTorqueConstraint(inv_dyn, tl, fric, discretization_scheme=1)
    a1, b1, cc1, F1, g1, _, _ = c1.compute_constraint_params(path, grid)
    ra, rb, rc, rF, rg, _, _ = constraint.canlinear_colloc_to_interpolate(a, b, c, F, g, None, None, grid, identical=True)
    for x, y in ((a1, ra), (b1, rb), (cc1, rc), (F1, rF), (g1, rg)):
        np.testing.assert_array_equal(x, y)
    assert F1.shape == (12, 6) and a1.shape == (9, 6)
    with pytest.raises(ValueError):
        constraint.JointTorqueConstraint(inv_dyn, np.ones((4, 2)), np.zeros(4)).compute_constraint_params(path, grid)
    assert "Torque limit" in repr(c0)


def test_cartesian_velocity_norm_host_tuple():
    """CartesianVelocityNorm (cpp/src/toppra/constraint/cartesian_velocity_norm.cpp:23-54): a = c = 0, b = v^T S v,
    F = [1], g = [limit]; constant limit -> identical F, varying limit -> per-gridpoint F and g; argument checks of
    CartesianVelocityNorm::check (.cpp:16-21)."""
    path = _FakePath()
    grid = np.linspace(0, 1, 7)
    J = np.arange(18.0).reshape(6, 3) / 10.0
    vel = lambda q, qd: J.dot(qd) * (1.0 + q[0])  # noqa: E731
    S = np.diag([1.0, 2.0, 3.0, 0.5, 0.5, 0.5])
    c0 = constraint.CartesianVelocityNorm(vel, S, 2.5, dof=3)
    a, b, c, F, g, ub, xb = c0.compute_constraint_params(path, grid)
    want = np.array([vel(q, qd).dot(S.dot(vel(q, qd))) for q, qd in zip(path(grid), path(grid, 1))])
    np.testing.assert_array_equal(b[:, 0], want)
    assert c0.identical and not a.any() and not c.any() and F.shape == (1, 1) and g.tolist() == [2.5] and ub is None and xb is None
    c1 = constraint.CartesianVelocityNorm(vel, velocity_limit=lambda s: (S * (1 + s), 1.0 + s))
    a, b, c, F, g, _, _ = c1.compute_constraint_params(path, grid)
    assert not c1.identical and F.shape == (7, 1, 1) and g.shape == (7, 1)
    np.testing.assert_array_equal(g[:, 0], 1.0 + grid)
    np.testing.assert_allclose(b[:, 0], want * (1 + grid), rtol=1e-15)
    for bad in (dict(S=np.eye(5), limit=1.0), dict(S=S, limit=-1.0), dict()):
        with pytest.raises(ValueError):
            constraint.CartesianVelocityNorm(vel, **bad)
    with pytest.raises(ValueError):
        constraint.CartesianVelocityNorm(vel, S, 1.0, dof=4).compute_constraint_params(path, grid)
    with pytest.raises(ValueError):
        constraint.CartesianVelocityNorm(lambda q, qd: qd, S, 1.0).compute_constraint_params(path, grid)


class _FakeTrajectory(object):
    def __init__(self, path, gridpoints, sd_vec):
        self.args = (path, gridpoints, sd_vec)
        self.path_interval = np.array([0.0, 2.5])


def test_parameterization_algorithm_base():
    """ParameterizationAlgorithm (reference algorithm.py:65-194): gridpoint validation, problem_data, parametrizer choice,
    compute_trajectory returning None unless the solve reports Ok."""
    from toppra_b200.algorithm import ParameterizationAlgorithm, ParameterizationData

    class Path(_FakePath):
        path_interval = np.array([0.0, 1.0])

    class Algo(ParameterizationAlgorithm):
        code = ParameterizationReturnCode.Ok

        def compute_parameterization(self, sd_start, sd_end, return_data=False):
            self._problem_data.return_code = self.code
            self._problem_data.sd_vec = np.full(len(self.gridpoints), sd_start + 1.0)

    grid = [0.0, 0.25, 0.5, 1.0]
    alg = Algo(["c"], Path(), grid, parametrizer="ParametrizeConstAccel")
    assert alg.constraints == ["c"] and alg._N == 3 and alg.parametrizer is ta.ParametrizeConstAccel
    assert isinstance(alg.problem_data, ParameterizationData)
    assert alg.problem_data.return_code == ParameterizationReturnCode.ErrUnknown
    assert alg.problem_data.sd_vec is None and alg.problem_data.K is None and alg.problem_data.X is None
    np.testing.assert_array_equal(alg.gridpoints, grid)
    np.testing.assert_array_equal(alg.problem_data.gridpoints, grid)
    assert repr(alg.problem_data).startswith("ParameterizationData(return_code:=<ParameterizationReturnCode.ErrUnknown")
    assert repr(alg.problem_data).endswith("N=4)")
    assert Algo([], Path(), grid).parametrizer is ta.ParametrizeSpline
    assert Algo([], Path(), grid, parametrizer="ParametrizeSpline").parametrizer is ta.ParametrizeSpline
    with pytest.raises(NotImplementedError):
        ParameterizationAlgorithm([], Path(), grid).compute_parameterization(0, 0)
    for bad in ([0.0, 0.5, 0.9], [0.1, 0.5, 1.0]):                 # ends must be the path interval
        with pytest.raises(ValueError, match="Invalid manually supplied gridpoints"):
            Algo([], Path(), bad)
    for bad in ([0.0, 0.5, 0.5, 1.0], [0.0, 0.7, 0.3, 1.0]):       # strictly increasing
        with pytest.raises(ValueError, match="Bad input gridpoints"):
            Algo([], Path(), bad)
    alg.parametrizer = _FakeTrajectory
    traj = alg.compute_trajectory(0.5, 0.0)
    assert isinstance(traj, _FakeTrajectory) and traj.args[0] is alg.path
    np.testing.assert_array_equal(traj.args[1], grid)
    np.testing.assert_array_equal(traj.args[2], np.full(4, 1.5))
    alg.code = ParameterizationReturnCode.FailUncontrollable
    assert alg.compute_trajectory() is None
    assert STATUS_CODES == tuple(ParameterizationReturnCode) and repr(STATUS_CODES[3]) == str(STATUS_CODES[3])
    assert repr(STATUS_CODES[0]) == "<ParameterizationReturnCode.Ok: 'Ok: Successful parametrization'>"
