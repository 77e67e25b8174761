"""Constraints — same names as the reference `toppra/constraint/__init__.py`."""
from .constraint import ConstraintType, DiscretizationType, Constraint
from .linear_constraint import LinearConstraint, canlinear_colloc_to_interpolate, RecordContext
from .linear_joint_acceleration import JointAccelerationConstraint
from .linear_joint_velocity import JointVelocityConstraint, JointVelocityConstraintVarying
from .linear_second_order import SecondOrderConstraint
from .joint_torque import JointTorqueConstraint
from .cartesian_velocity_norm import CartesianVelocityNorm
from .conic_constraint import ConicConstraint, RobustLinearConstraint

__all__ = ["ConstraintType", "DiscretizationType", "Constraint", "LinearConstraint",
           "canlinear_colloc_to_interpolate", "JointAccelerationConstraint", "JointVelocityConstraint", "JointVelocityConstraintVarying",
           "SecondOrderConstraint", "JointTorqueConstraint", "CartesianVelocityNorm", "RecordContext", "ConicConstraint", "RobustLinearConstraint"]
