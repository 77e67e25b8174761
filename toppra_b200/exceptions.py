"""Error types of toppra_b200; the names are those of the reference's `toppra/exceptions.py`."""


class ToppraError(Exception):
    pass


class BadInputVelocities(ToppraError):
    """sd_start / sd_end cannot be used (negative values, reachability_algorithm.py:272-276)."""


class SolverNotFound(ToppraError):
    """The requested solver wrapper is not provided by this build."""
