"""TOPPRAsd — TOPP-RA with a specified duration; same surface as the reference
`toppra/algorithm/reachabilitybased/desired_duration_algorithm.py:20-234`.

The fastest and the slowest parameterisations are two launches of the scan kernel (csrc/tb_scan.cu, flags
TB_SCAN_SD_FORWARD / TB_SCAN_SD_SLOW: the reference's TOPPRAsd forward rules — no retry, x_next - 1e-5 clip); the
bisection on their convex combination is O(N) host arithmetic on two vectors, written like the reference."""
import logging

import numpy as np

from .reachability_algorithm import ReachabilityAlgorithm
from .. import algorithm as algo

logger = logging.getLogger(__name__)


def _compute_duration(xs, deltas):
    sds = np.sqrt(xs)
    t = 0
    for i in range(len(deltas)):
        t += 2 * deltas[i] / (sds[i + 1] + sds[i] + 1e-9)
    return t


class TOPPRAsd(ReachabilityAlgorithm):
    """TOPPRA with specified duration: bisection between the fastest and the slowest parameterisation."""

    def set_desired_duration(self, desired_duration: float):
        self.desired_duration = desired_duration

    def compute_parameterization(self, sd_start, sd_end, return_data=False, atol=1e-5):
        assert sd_end >= 0 and sd_start >= 0, "Path velocities must be positive"
        fast = self.solver_wrapper.parameterize(sd_start, sd_end, sd_forward="fast")
        K = fast["K"]
        if algo.STATUS_CODES[fast["status"]] == algo.ParameterizationReturnCode.FailUncontrollable:
            if np.isnan(K).any():
                logger.warning("The set of controllable velocities at the beginning is empty!")
            else:
                self.problem_data.K = K
                logger.warning("The initial velocity is not controllable.")
            self._problem_data.return_code = algo.ParameterizationReturnCode.FailUncontrollable
            return (None, None, None, K) if return_data else (None, None, None)
        self.problem_data.K = K
        slow = self.solver_wrapper.parameterize(sd_start, sd_end, sd_forward="slow")
        deltas = self.solver_wrapper.get_deltas()
        # with sd_forward the kernel returns the squared velocities x (TOPPRAsd combines those)
        xs, us = fast["sd"], fast["u"]
        xs_slow, us_slow = slow["sd"], slow["u"]
        N = self._N
        v_vec_alpha = np.zeros((N, self.solver_wrapper.get_no_vars() - 2))
        duration = _compute_duration(xs, deltas)
        duration_slow = _compute_duration(xs_slow, deltas)
        if duration > self.desired_duration:
            logger.warning("Desired duration %f seconds is not achievable. Returning the fastest parameterization "
                           "with duration %f seconds", self.desired_duration, duration)
            alpha = 1.0
        elif duration_slow < self.desired_duration:
            logger.warning("Desired duration %f seconds is not achievable. Returning the slowest parameterization "
                           "with duration %f seconds", self.desired_duration, duration_slow)
            alpha = .0
        else:
            alpha_low, alpha_high, diff = 1.0, 0.0, 10
            while diff > atol:
                alpha = 0.5 * (alpha_low + alpha_high)
                duration_alpha = _compute_duration(alpha * xs + (1 - alpha) * xs_slow, deltas)
                if duration_alpha < self.desired_duration:
                    alpha_low = alpha
                    diff = self.desired_duration - duration_alpha
                else:
                    alpha_high = alpha
                    diff = duration_alpha - self.desired_duration
        xs_alpha = alpha * xs + (1 - alpha) * xs_slow
        us_alpha = alpha * us + (1 - alpha) * us_slow
        sd_vec = np.sqrt(xs_alpha)
        sdd_vec = np.copy(us_alpha)
        self.problem_data.sd_vec = sd_vec
        self.problem_data.sdd_vec = sdd_vec
        if np.isnan(sd_vec).any():
            self.problem_data.return_code = algo.ParameterizationReturnCode.ErrUnknown
        else:
            self.problem_data.return_code = algo.ParameterizationReturnCode.Ok
        if return_data:
            return sdd_vec, sd_vec, v_vec_alpha, K
        return sdd_vec, sd_vec, v_vec_alpha
