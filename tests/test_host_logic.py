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
    for name in ("JointVelocityConstraint", "JointAccelerationConstraint", "SecondOrderConstraint", "LinearConstraint",
                 "ConstraintType", "DiscretizationType", "canlinear_colloc_to_interpolate"):
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
        with pytest.raises(NotImplementedError):
            engine.parse_bc("periodic", 1, 2, "cpu")
        with pytest.raises(ValueError):
            engine.parse_bc("bogus", 1, 2, "cpu")
