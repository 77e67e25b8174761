"""Enums and the abstract base of path-parameterisation constraints.  Public surface (names, getters, repr layout)
follows the reference's `toppra/constraint/constraint.py:10-103`; the code is this package's own."""
import enum


class ConstraintType(enum.Enum):
    """How a constraint presents itself to the solver: rows (CanonicalLinear) or second-order cones (CanonicalConic)."""

    Unknown = -1
    CanonicalLinear = 0
    CanonicalConic = 1


class DiscretizationType(enum.Enum):
    """Collocation: the constraint holds at the gridpoints; Interpolation: also at the next gridpoint expressed in
    the current stage's variables (linear_constraint.py:84-192)."""

    Collocation = 0
    Interpolation = 1


_BY_VALUE = {t.value: t for t in DiscretizationType}


class Constraint(object):
    """Abstract constraint.  Subclasses set `constraint_type`, `discretization_type`, `dof`, `n_extra_vars` and
    `_format_string` (the body of the repr) and implement `compute_constraint_params`."""

    def compute_constraint_params(self, path, gridpoints, *args, **kwargs):
        raise NotImplementedError

    def set_discretization_type(self, discretization_type):
        """Accepts the enum or its integer value (0 = Collocation, 1 = Interpolation)."""
        if isinstance(discretization_type, DiscretizationType):
            self.discretization_type = discretization_type
            return
        for value, scheme in _BY_VALUE.items():
            if discretization_type == value:
                self.discretization_type = scheme
                return
        raise NotImplementedError("Discretization type: {:} not implemented!".format(discretization_type))

    def get_discretization_type(self):
        return self.discretization_type

    def get_constraint_type(self):
        return self.constraint_type

    def get_dof(self):
        return self.dof

    def get_no_extra_vars(self):
        return self.n_extra_vars

    def __repr__(self):
        head = ["{:}(".format(type(self).__name__),
                "    Type: {:}".format(self.constraint_type),
                "    Discretization Scheme: {:}".format(self.discretization_type)]
        return "\n".join(head) + "\n" + self._format_string + ")"
