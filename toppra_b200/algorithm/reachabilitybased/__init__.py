from .reachability_algorithm import ReachabilityAlgorithm
from .time_optimal_algorithm import TOPPRA

__all__ = ["ReachabilityAlgorithm", "TOPPRA"]
