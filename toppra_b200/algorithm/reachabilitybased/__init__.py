"""Reachability-analysis based algorithms (controllable / reachable / feasible sets on the GPU scan kernels):
the reference's `toppra/algorithm/reachabilitybased` package surface."""
from .reachability_algorithm import ReachabilityAlgorithm
from .desired_duration_algorithm import TOPPRAsd
from .time_optimal_algorithm import TOPPRA

__all__ = ["ReachabilityAlgorithm", "TOPPRA", "TOPPRAsd"]
