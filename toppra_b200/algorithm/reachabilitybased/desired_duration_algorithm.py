"""TOPPRAsd — TOPP-RA with a specified duration; same surface as the reference
`toppra/algorithm/reachabilitybased/desired_duration_algorithm.py:20-234`.

Everything runs on the GPU: the controllable sets and the fastest pass are one scan launch (csrc/tb_scan.cu, flag
TB_SCAN_SD_FORWARD: the reference's TOPPRAsd forward rules — no retry, x_next - 1e-5 clip), the slowest pass a second one
(TB_SCAN_SD_SLOW), and the bisection on the blend alpha * fastest + (1 - alpha) * slowest is tb_sd_bisect
(csrc/tb_frows.cu), which keeps the reference's running duration sum.  This class is the B = 1 case of
`toppra_b200.BatchTOPPRAsd`."""
import logging

import numpy as np

from .reachability_algorithm import ReachabilityAlgorithm
from .. import algorithm as algo

logger = logging.getLogger(__name__)


class TOPPRAsd(ReachabilityAlgorithm):
    """TOPPRA with specified duration: a convex combination of the fastest and the slowest parameterisation."""

    def set_desired_duration(self, desired_duration: float):
        self.desired_duration = desired_duration

    def compute_parameterization(self, sd_start, sd_end, return_data=False, atol=1e-5):
        """(sdd_vec, sd_vec, v_vec[, K]) like the reference; (None, None, None[, K]) when the instance is not
        controllable.  An unachievable duration returns the fastest / slowest parameterisation (reference :143-152)."""
        assert sd_end >= 0 and sd_start >= 0, "Path velocities must be positive"
        res = self.solver_wrapper.parameterize_sd(sd_start, sd_end, self.desired_duration, atol)
        K = res["K"]
        if algo.STATUS_CODES[res["status"]] == algo.ParameterizationReturnCode.FailUncontrollable:
            if np.isnan(K).any():
                logger.warning("The set of controllable velocities at the beginning is empty!")
            else:
                self.problem_data.K = K
                logger.warning("The initial velocity is not controllable.")
            self._problem_data.return_code = algo.ParameterizationReturnCode.FailUncontrollable
            return (None, None, None, K) if return_data else (None, None, None)
        self.problem_data.K = K
        alpha, fastest, slowest = res["alpha"], res["duration_fast"], res["duration_slow"]
        if fastest > self.desired_duration:
            logger.warning("Desired duration %f seconds is not achievable. Returning the fastest parameterization "
                           "with duration %f seconds", self.desired_duration, fastest)
        elif slowest < self.desired_duration:
            logger.warning("Desired duration %f seconds is not achievable. Returning the slowest parameterization "
                           "with duration %f seconds", self.desired_duration, slowest)
        self.alpha = alpha
        self.problem_data.sd_vec, self.problem_data.sdd_vec = res["sd"], res["u"]
        self.problem_data.return_code = algo.STATUS_CODES[res["blend_status"]]
        v_vec = np.zeros((self._N, self.solver_wrapper.get_no_vars() - 2))
        if return_data:
            return res["u"], res["sd"], v_vec, K
        return res["u"], res["sd"], v_vec
