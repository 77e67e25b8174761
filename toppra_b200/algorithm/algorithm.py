"""Result types and the abstract base of the parameterisation algorithms.  Names, fields, enum values, messages and
error behaviour follow the reference's `toppra/algorithm/algorithm.py:27-194` (they are the drop-in surface); the code
is this package's own."""
import enum
import logging
import time

import numpy as np

from .. import interpolator
from .. import parametrizer as tparam

logger = logging.getLogger(__name__)


class ParameterizationReturnCode(enum.Enum):
    """Outcome of a parameterisation attempt; the member ORDER is the kernels' status integer
    (include/toppra_b200.h TB_STATUS_*), the values are the reference's messages (algorithm.py:49-56)."""

    Ok = "Ok: Successful parametrization"
    ErrUnknown = "Error: Unknown issue"
    ErrShortPath = "Error: Input path is very short"
    FailUncontrollable = "Error: Instance is not controllable"
    ErrForwardPassFail = "Error: Forward pass fail. Numerical errors occured"

    def __str__(self):  # like the reference, str() shows the full <Class.Member: 'message'> form (:58-62)
        return enum.Enum.__repr__(self)

    def __repr__(self):
        return enum.Enum.__repr__(self)


#: kernel status integer -> enum member
STATUS_CODES = tuple(ParameterizationReturnCode)


class ParameterizationData(object):
    """What a solve leaves behind (algorithm.py:27-46): `return_code`, `gridpoints`, `sd_vec`, `sdd_vec`, the
    controllable sets `K` and the feasible sets `X`; everything but the code starts as None."""

    _ARRAYS = ("gridpoints", "sd_vec", "sdd_vec", "K", "X")

    def __init__(self, *arg, **kwargs):
        self.return_code = ParameterizationReturnCode.ErrUnknown
        for name in self._ARRAYS:
            setattr(self, name, None)

    def __repr__(self):
        return "ParameterizationData(return_code:={}, N={:d})".format(self.return_code, self.gridpoints.shape[0])


class ParameterizationAlgorithm(object):
    """Base of TOPPRA / TOPPRAsd (algorithm.py:65-194): holds constraints, path, gridpoints and the output parametrizer
    class; subclasses implement `compute_parameterization`."""

    def __init__(self, constraint_list, path, gridpoints=None, parametrizer=None, gridpt_max_err_threshold=1e-3,
                 gridpt_min_nb_points=100):
        self.constraints = constraint_list
        self.path = path
        if gridpoints is None:  # data-dependent grid, interpolator.py:49-122
            gridpoints = interpolator.propose_gridpoints(path, max_err_threshold=gridpt_max_err_threshold,
                                                         min_nb_points=gridpt_min_nb_points)
            logger.info("No gridpoint specified. Automatically choose a gridpoint with %d points", len(gridpoints))
        first, last = path.path_interval[0], path.path_interval[1]
        if first != gridpoints[0] or last != gridpoints[-1]:
            raise ValueError("Invalid manually supplied gridpoints.")
        grid = np.array(gridpoints)
        if bool(np.any(grid[1:] <= grid[:-1])):
            logger.fatal("Input gridpoints are not monotonically increasing.")
            raise ValueError("Bad input gridpoints.")
        self.gridpoints = grid
        self._N = grid.shape[0] - 1
        self._problem_data = ParameterizationData()
        self._problem_data.gridpoints = grid.copy()
        if parametrizer is None or parametrizer == "ParametrizeSpline":
            self.parametrizer = tparam.ParametrizeSpline
        elif parametrizer == "ParametrizeConstAccel":
            self.parametrizer = tparam.ParametrizeConstAccel

    @property
    def constraints(self):
        return self._constraints

    @constraints.setter
    def constraints(self, value):
        self._constraints = value

    @property
    def problem_data(self):
        """The `ParameterizationData` of the last solve."""
        return self._problem_data

    def compute_parameterization(self, sd_start, sd_end, return_data=False):
        raise NotImplementedError

    def compute_trajectory(self, sd_start=0, sd_end=0):
        """Solve, then hand path + gridpoints + velocities to the parametrizer; None when the solve did not succeed
        (algorithm.py:163-194)."""
        began = time.time()
        self.compute_parameterization(sd_start, sd_end)
        data = self.problem_data
        if data.return_code != ParameterizationReturnCode.Ok:
            logger.warning("Fail to parametrize path. Return code: %s", data.return_code)
            return None
        trajectory = self.parametrizer(self.path, data.gridpoints, data.sd_vec)
        logger.info("Successfully parametrize path. Duration: %.3f, previously %.3f)",
                    trajectory.path_interval[1], self.path.path_interval[1])
        logger.info("Finish parametrization in %.3f secs", time.time() - began)
        return trajectory

    def inspect(self, compute=True):
        """Plot what the last solve left in `problem_data` over the gridpoint index: feasible sets X, controllable sets K
        and the squared velocity profile (algorithm.py:196-213).  Needs matplotlib, which is imported here only."""
        import matplotlib.pyplot as plt
        data = self.problem_data
        for sets, style, label in ((data.X, dict(c="green"), "Feasible sets"),
                                   (data.K, dict(c="red", ls="--"), "Controllable sets")):
            if sets is not None:
                plt.plot(sets[:, 0], label=label, **style)
                plt.plot(sets[:, 1], **style)
        if data.sd_vec is not None:
            plt.plot(np.square(data.sd_vec), label="Velocity profile")
        plt.title("Path-position path-velocity plot")
        plt.xlabel("Path position")
        plt.ylabel("Path velocity square")
        plt.legend()
        plt.tight_layout()
        plt.show()
