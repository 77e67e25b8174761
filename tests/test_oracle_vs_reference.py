"""CPU: live differential test of the oracle against the UNMODIFIED reference build (oracle/_ref), on random
problems beyond the committed fixtures.  Skipped where oracle/_ref was not built (run oracle/build_ref.sh
where /root/reference exists; the build travels to the GPU box)."""
import warnings

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.ref_loader import load_reference, reference_available
from problems import make_path

pytestmark = pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    ta = load_reference()
    import toppra.algorithm as algo
    import toppra.constraint as constraint
    return ta, algo, constraint


@pytest.mark.parametrize("vel_active", [False, True])
def test_random_paths_bit_exact(ref, vel_active):
    ta, algo, constraint = ref
    ss = np.linspace(0, 1, 5)
    for seed in range(5000, 5040):
        G = 60 + (seed % 5) * 35
        grid = np.linspace(0, 1, G)
        way, vlim, alim = make_path(seed, vel_active=vel_active)
        path = ta.SplineInterpolator(ss, way)
        inst = algo.TOPPRA([constraint.JointVelocityConstraint(vlim), constraint.JointAccelerationConstraint(alim)],
                           path, gridpoints=grid, solver_wrapper="seidel")
        sd0 = 0.0 if seed % 3 else 0.02
        sdd, sd, _, K = inst.compute_parameterization(sd0, 0.0, return_data=True)
        c = orc.cubic_spline_fit(ss, way)
        assert np.array_equal(c, path.cspl.c)
        o = orc.solve_velacc(c, ss, grid, vlim, alim, True, sd0, 0.0)
        assert np.array_equal(o["K"], K, equal_nan=True)
        if sd is None:
            assert o["status"] == 3
        else:
            assert np.array_equal(o["sd"], sd, equal_nan=True) and np.array_equal(o["u"], sdd, equal_nan=True)


def test_lp_shims_random(ref):
    """Random LPs with random warm-start pairs through the reference's solve_lp2d shim (pyx:65-87)."""
    import toppra.solverwrapper.cy_seidel_solverwrapper as seidel
    rng = np.random.RandomState(0)
    for trial in range(300):
        n = rng.randint(1, 40)
        v = rng.randn(3)
        a, b = rng.randn(2, n)
        c = -rng.rand(n) if trial % 2 else rng.randn(n) * 0.3 - 0.5
        low, high = np.r_[-1.0, -2.0], np.r_[1.5, 0.7]
        act = rng.randint(-4, n + 2, size=2)
        r0, val0, var0, act0 = seidel.solve_lp2d(v, a, b, c, low, high, act.astype(np.int64))
        r1, val1, var1, act1 = orc.lp2d(v, a, b, c, low, high, act)
        assert r0 == r1
        if r0:
            assert val0 == val1 and np.array_equal(np.asarray(var0), var1) and np.array_equal(np.asarray(act0), act1)
