from .reachability_algorithm import ReachabilityAlgorithm
from .time_optimal_algorithm import TOPPRA
from .desired_duration_algorithm import TOPPRAsd

__all__ = ["ReachabilityAlgorithm", "TOPPRA", "TOPPRAsd"]
