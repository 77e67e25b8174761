"""Parametrization algorithms — same names as the reference `toppra/algorithm/__init__.py`."""
from .algorithm import ParameterizationAlgorithm, ParameterizationData, ParameterizationReturnCode, STATUS_CODES
from .reachabilitybased import TOPPRA, TOPPRAsd, ReachabilityAlgorithm

__all__ = ["ParameterizationData", "ParameterizationAlgorithm", "ParameterizationReturnCode", "TOPPRA",
           "TOPPRAsd", "ReachabilityAlgorithm", "STATUS_CODES"]
