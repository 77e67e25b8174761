"""Exceptions — same names as the reference `toppra/exceptions.py`."""


class ToppraError(Exception):
    """A generic error class."""


class BadInputVelocities(ToppraError):
    """Raised when given input velocity is invalid."""
