"""See oracle/stubs/matplotlib/__init__.py.  Any plotting call raises."""


def __getattr__(name):
    raise AttributeError("matplotlib stub: plotting (%s) is not available" % name)
