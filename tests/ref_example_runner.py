"""Runs ONE of the reference's example scripts (unmodified, read in place) and dumps the numeric globals it leaves behind.

usage: python ref_example_runner.py <engine> <script.py> <out.npz>
  engine = reference   the unmodified reference build under oracle/_ref
           cpu_double  toppra_b200 under the name `toppra`, kernels replaced by tests/cpu_engine.py
           gpu         toppra_b200 under the name `toppra`, real engine
matplotlib is replaced by a stand-in that swallows every call (the examples only plot with it).  Used by
tests/test_reference_suite_dropin.py."""
import os
import runpy
import sys
import types

import numpy as np


class _Sink(types.ModuleType):
    """Module / object whose every attribute is a callable returning another sink (plt.subplots(...)[1][0].plot(...))."""

    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Sink(name)

    def __call__(self, *args, **kwargs):
        return _Sink("call")

    def __getitem__(self, idx):
        return _Sink("item")

    def __iter__(self):
        return iter([_Sink("a"), _Sink("b")])


def main(engine, script, out):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    sys.path[:0] = [root, here]
    if engine == "reference":
        from oracle.ref_loader import load_reference
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules[name] = _Sink(name)
        load_reference()
    else:
        os.environ["TB_REF_SUITE_ENGINE"] = engine
        import ref_suite_plugin  # noqa: F401  (import alias + engine double)
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules[name] = _Sink(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.argv = [script]
    g = runpy.run_path(script, run_name="__main__")
    keep = {}
    for k, v in g.items():
        if k.startswith("_") or isinstance(v, (types.ModuleType, _Sink)):
            continue
        if isinstance(v, (float, int, np.floating, np.integer)) and not isinstance(v, bool):
            keep[k] = np.float64(v)
        elif isinstance(v, np.ndarray) and v.dtype.kind in "fi":
            keep[k] = v
        elif hasattr(v, "duration") and hasattr(v, "dof"):          # a trajectory: record its duration
            try:
                keep[k + "__duration"] = np.float64(v.duration)
            except Exception:
                pass
    np.savez(out, **keep)


if __name__ == "__main__":
    main(*sys.argv[1:4])
