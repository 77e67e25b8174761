"""TOPPRA — same surface as the reference `time_optimal_algorithm.py:8-92`."""
from .reachability_algorithm import ReachabilityAlgorithm


class TOPPRA(ReachabilityAlgorithm):
    """Time-Optimal Path Parameterization based on Reachability Analysis (TOPPRA).

    >>> instance = algo.TOPPRA([pc_vel, pc_acc], path)
    >>> jnt_traj = instance.compute_trajectory()  # rest-to-rest motion
    >>> instance.problem_data # intermediate result

    The forward step (greedy maximal controllable velocity, reference :55-92) is part of the fused
    backward+forward kernel (csrc/tb_scan.cu)."""
