"""Abstract parametrization algorithm and result types — same surface as the reference
`toppra/algorithm/algorithm.py:27-194`."""
import enum
import logging
import time
from typing import Optional

import numpy as np

from .. import interpolator
from .. import parametrizer as tparam

logger = logging.getLogger(__name__)


class ParameterizationData(object):
    """Internal data and output (reference algorithm.py:27-46)."""

    def __init__(self, *arg, **kwargs) -> None:
        self.return_code: ParameterizationReturnCode = ParameterizationReturnCode.ErrUnknown
        self.gridpoints: Optional[np.ndarray] = None
        self.sd_vec: Optional[np.ndarray] = None
        self.sdd_vec: Optional[np.ndarray] = None
        self.K: Optional[np.ndarray] = None
        self.X: Optional[np.ndarray] = None

    def __repr__(self):
        return "ParameterizationData(return_code:={}, N={:d})".format(self.return_code, self.gridpoints.shape[0])


class ParameterizationReturnCode(enum.Enum):
    """Return codes from a parametrization attempt (reference algorithm.py:49-62)."""

    Ok = "Ok: Successful parametrization"
    ErrUnknown = "Error: Unknown issue"
    ErrShortPath = "Error: Input path is very short"
    FailUncontrollable = "Error: Instance is not controllable"
    ErrForwardPassFail = "Error: Forward pass fail. Numerical errors occured"

    def __repr__(self):
        return super(ParameterizationReturnCode, self).__repr__()

    def __str__(self):
        return super(ParameterizationReturnCode, self).__repr__()


#: kernel status integer (include/toppra_b200.h TB_STATUS_*) -> enum member
STATUS_CODES = (
    ParameterizationReturnCode.Ok,
    ParameterizationReturnCode.ErrUnknown,
    ParameterizationReturnCode.ErrShortPath,
    ParameterizationReturnCode.FailUncontrollable,
    ParameterizationReturnCode.ErrForwardPassFail,
)


class ParameterizationAlgorithm(object):
    """Base parametrization algorithm class (reference algorithm.py:65-194)."""

    def __init__(self, constraint_list, path, gridpoints=None, parametrizer=None,
                 gridpt_max_err_threshold: float = 1e-3, gridpt_min_nb_points: int = 100):
        self.constraints = constraint_list
        self.path = path
        self._problem_data = ParameterizationData()
        if gridpoints is None:
            gridpoints = interpolator.propose_gridpoints(
                path, max_err_threshold=gridpt_max_err_threshold, min_nb_points=gridpt_min_nb_points)
            logger.info("No gridpoint specified. Automatically choose a gridpoint with %d points", len(gridpoints))
        if path.path_interval[0] != gridpoints[0] or path.path_interval[1] != gridpoints[-1]:
            raise ValueError("Invalid manually supplied gridpoints.")
        self.gridpoints = np.array(gridpoints)
        self._problem_data.gridpoints = np.array(gridpoints)
        self._N = len(gridpoints) - 1
        for i in range(self._N):
            if gridpoints[i + 1] <= gridpoints[i]:
                logger.fatal("Input gridpoints are not monotonically increasing.")
                raise ValueError("Bad input gridpoints.")
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
    def problem_data(self) -> ParameterizationData:
        """Data obtained when solving the path parametrization."""
        return self._problem_data

    def compute_parameterization(self, sd_start: float, sd_end: float, return_data: bool = False):
        raise NotImplementedError

    def compute_trajectory(self, sd_start: float = 0, sd_end: float = 0):
        """Compute the resulting joint trajectory (None if the path cannot be parameterised)."""
        t0 = time.time()
        self.compute_parameterization(sd_start, sd_end)
        if self.problem_data.return_code != ParameterizationReturnCode.Ok:
            logger.warning("Fail to parametrize path. Return code: %s", self.problem_data.return_code)
            return None
        outputtraj = self.parametrizer(self.path, self.problem_data.gridpoints, self.problem_data.sd_vec)
        logger.info("Successfully parametrize path. Duration: %.3f, previously %.3f)",
                    outputtraj.path_interval[1], self.path.path_interval[1])
        logger.info("Finish parametrization in %.3f secs", time.time() - t0)
        return outputtraj
