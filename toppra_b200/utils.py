"""Small host-side helpers that the reference exposes at package level (reference: toppra/utils.py:15-41).

Only what the hot path's callers use is here: `setup_logging` and the `deprecated` decorator.  The OpenRAVE helpers of the
reference's utils.py / planning_utils.py and `smooth_singularities` (qpOASES-era post-processing) are out of scope
("out of scope")."""
import functools
import logging
import warnings

_FORMAT = "%(levelname)5s [%(filename)s : %(lineno)d] %(message)s"


def setup_logging(level="WARN"):
    """Console logging for the package's loggers (reference: toppra/utils.py:32-41).  Both this package's logger name and
    the reference's ("toppra") are configured, so a caller that switched packages keeps its log output."""
    for name in ("toppra_b200", "toppra"):
        log = logging.getLogger(name)
        log.setLevel(level)
        if not any(getattr(h, "_tb_console", False) for h in log.handlers):
            handler = logging.StreamHandler()
            handler.setLevel(logging.DEBUG)
            handler.setFormatter(logging.Formatter(_FORMAT))
            handler._tb_console = True
            log.addHandler(handler)


def deprecated(func):
    """Decorator: calling `func` emits a DeprecationWarning (reference: toppra/utils.py:15-29)."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        warnings.warn("Call to deprecated function {} in module {}.".format(func.__name__, func.__module__),
                      category=DeprecationWarning, stacklevel=2)
        return func(*args, **kwargs)

    return wrapper
