"""Base class and enums of path-parametrization constraints — same surface as the reference
`toppra/constraint/constraint.py:10-103`."""
from enum import Enum
import logging

logger = logging.getLogger(__name__)


class ConstraintType(Enum):
    """Type of path parametrization constraint."""

    Unknown = -1
    CanonicalLinear = 0
    CanonicalConic = 1


class DiscretizationType(Enum):
    """Discretization scheme: Collocation (0) or Interpolation (1)."""

    Collocation = 0
    Interpolation = 1


class Constraint(object):
    """The base constraint class."""

    def __repr__(self):
        string = self.__class__.__name__ + "(\n"
        string += "    Type: {:}".format(self.constraint_type) + "\n"
        string += "    Discretization Scheme: {:}".format(self.discretization_type) + "\n"
        string += self._format_string
        string += ")"
        return string

    def get_dof(self):
        return self.dof

    def get_no_extra_vars(self):
        return self.n_extra_vars

    def get_constraint_type(self):
        return self.constraint_type

    def get_discretization_type(self):
        return self.discretization_type

    def set_discretization_type(self, discretization_type):
        """Discretization type: Collocation or Interpolation (int 0/1 or the enum)."""
        if discretization_type == 0:
            self.discretization_type = DiscretizationType.Collocation
        elif discretization_type == 1:
            self.discretization_type = DiscretizationType.Interpolation
        elif (discretization_type == DiscretizationType.Collocation
              or discretization_type == DiscretizationType.Interpolation):
            self.discretization_type = discretization_type
        else:
            raise NotImplementedError("Discretization type: {:} not implemented!".format(discretization_type))

    def compute_constraint_params(self, path, gridpoints, *args, **kwargs):
        """Evaluate parameters of the constraint."""
        raise NotImplementedError
