"""CPU (-m "not gpu"): the scalar model of the CUDA scan kernel's Seidel shortcuts (oracle/shortcut_model.c, same
decision rules and margins as toppra_b200/csrc/tb_scan.cu lp2d_impl) must reproduce the sequential restatement of
cy_solve_lp2d (cy_seidel_solverwrapper.pyx:149-390) bit for bit — optimum, active pair and feasibility — on every
2-variable LP of benchmark-like, velocity-limited, badly scaled and near-degenerate problems, and on the
robustness suite whose spurious-infeasibility cases the reference's tests hold (tests/tests/retime/robustness).
The kernel itself is checked against the oracle by the -m gpu tests."""
import numpy as np
import pytest

from oracle import oracle as orc
from problems import make_batch, make_batch_fast


def _paths(B, vel_active, seed):
    if vel_active:
        ss, way, vlim, alim = make_batch(B, seed, vel_active=True)
    else:
        ss, way, vlim, alim = make_batch_fast(B, seed)
    c = np.stack([orc.cubic_spline_fit(ss, way[b]) for b in range(B)])
    return c, np.tile(ss, (B, 1)), vlim, alim


def test_model_on_benchmark_like_paths():
    grid = np.linspace(0, 1, 200)
    with orc.shortcut_model() as sm:
        for vel_active, B in ((False, 3072), (True, 1024)):
            c, x, vlim, alim = _paths(B, vel_active, 4321)
            r = orc.solve_velacc_batch(c, x, grid, vlim, alim, nthreads=1)
            assert not r["status"].any()
        st = sm.stats()
    assert st["mismatches"] == 0 and st["lps"] == 4096 * 398
    # both shortcuts are actually taken, and they remove most projected re-solves (2.9 -> about 1.1 per LP)
    assert st["a_used"] > 0.9 * 4096 * 199 and st["b_used"] > 0.5 * 3072 * 199
    assert st["resolves_model"] < 0.45 * st["resolves_ref"]


def _degenerate_rows(rng, G, R0, kind, eps):
    a, b = rng.randn(G, R0), rng.randn(G, R0)
    c = -rng.rand(G, R0) * 10 ** rng.uniform(-1, 1)
    if kind == 0:    # every coefficient perturbed
        a2, b2, c2 = a * (1 + eps * rng.randn(G, R0)), b * (1 + eps * rng.randn(G, R0)), c * (1 + eps * rng.randn(G, R0))
    elif kind == 1:  # the same line with another normal length
        sc = 10 ** rng.uniform(-3, 3, size=(G, R0))
        a2, b2, c2 = a * sc, b * sc, c * sc + eps * rng.randn(G, R0)
    elif kind == 2:  # parallel, offset by eps
        a2, b2, c2 = a.copy(), b.copy(), c + eps * rng.randn(G, R0)
    else:            # slightly rotated
        a2, b2, c2 = a + eps * rng.randn(G, R0), b.copy(), c.copy()
    perm = rng.permutation(2 * R0)
    return np.stack((np.concatenate((a, a2), 1)[:, perm], np.concatenate((b, b2), 1)[:, perm],
                     np.concatenate((c, c2), 1)[:, perm]), axis=1)


def test_model_on_near_degenerate_rows():
    rng = np.random.RandomState(5)
    with orc.shortcut_model() as sm:
        for it in range(12000):
            G, R0 = rng.randint(5, 40), rng.randint(2, 10)
            rows = _degenerate_rows(rng, G, R0, it % 4, 10 ** rng.uniform(-14, -6))
            xb = np.stack((np.zeros(G), np.full(G, 10 ** rng.uniform(-2, 3))), axis=1)
            orc.solve_rows(rows, xb, np.linspace(0, 1, G), 0, 0)
        st = sm.stats()
    assert st["mismatches"] == 0 and st["lps"] > 400000 and st["a_used"] > 40000 and st["b_used"] > 4000


def test_model_on_badly_scaled_problems():
    """Coefficients down to 1e-8 (every pair of rows looks 'parallel' to the 1e-10 test of pyx:339-345), optima up to the
    +-1e10 sentinel of the 1-D LP (pyx:376-383): the cases in which a skipped visit would have ended the reference's solve."""
    rng = np.random.RandomState(6)
    with orc.shortcut_model() as sm:
        for it in range(12000):
            G, R = rng.randint(5, 40), rng.randint(2, 20)
            sa, sb, sc = 10 ** rng.uniform(-8, 1), 10 ** rng.uniform(-8, 1), 10 ** rng.uniform(-3, 5)
            if it % 3 == 0:   # smooth along the path: warm-start pairs stay valid
                a = np.cumsum(rng.randn(G, R) * 0.05, 0) * sa + rng.randn(1, R) * sa
                b = np.cumsum(rng.randn(G, R) * 0.05, 0) * sb + rng.randn(1, R) * sb
            else:
                a, b = rng.randn(G, R) * sa, rng.randn(G, R) * sb
            rows = np.stack((a, b, -rng.rand(G, R) * sc), axis=1)
            xb = np.stack((np.zeros(G), np.full(G, 1e8 if it % 2 else 10 ** rng.uniform(-2, 9))), axis=1)
            orc.solve_rows(rows, xb, np.linspace(0, 1, G), 0, 0)
        for it in range(2400):
            dof, nway = rng.randint(1, 8), rng.randint(4, 8)
            way = rng.randn(nway, dof) * 10 ** rng.uniform(-6, 3)
            vl = 10 ** rng.uniform(-2, 4) * (1 + rng.rand(dof))
            al = 10 ** rng.uniform(-2, 5) * (1 + rng.rand(dof))
            ss = np.linspace(0, 10 ** rng.uniform(-2, 2), nway)
            c = orc.cubic_spline_fit(ss, way, "clamped" if it % 2 else "not-a-knot")
            orc.solve_velacc(c, ss, np.linspace(ss[0], ss[-1], rng.randint(10, 120)), np.stack((-vl, vl), 1),
                             np.stack((-al, al), 1), True, 0, 0)
        st = sm.stats()
    assert st["mismatches"] == 0 and st["lps"] > 400000 and st["a_declined"] > 4000


def test_model_on_robustness_suite(golden):
    g = golden("p4_robustness_suite")
    with orc.shortcut_model() as sm:
        for name in g["names"]:
            o = orc.solve_velacc(g[name + "_c"], g[name + "_ss"], g[name + "_grid"], g[name + "_vlim"], g[name + "_alim"],
                                 True, 0, 0)
            assert o["status"] == int(g[name + "_status"])
        st = sm.stats()
    assert st["mismatches"] == 0 and st["lps"] > 1000


def test_model_margins_matter():
    """With the violation margin switched off the model must disagree on near-degenerate rows: the test above is able to
    see a wrong shortcut."""
    rng = np.random.RandomState(5)
    with orc.shortcut_model() as sm:
        sm._lib.orc_shortcut_model_margins.argtypes = [orc.ctypes.c_double, orc.ctypes.c_double]
        sm._lib.orc_shortcut_model_margins(0.0, 90.0)
        try:
            for it in range(400):
                G, R0 = rng.randint(5, 40), rng.randint(2, 10)
                rows = _degenerate_rows(rng, G, R0, it % 4, 10 ** rng.uniform(-14, -6))
                xb = np.stack((np.zeros(G), np.full(G, 10 ** rng.uniform(-2, 3))), axis=1)
                orc.solve_rows(rows, xb, np.linspace(0, 1, G), 0, 0)
            st = sm.stats()
        finally:
            sm._lib.orc_shortcut_model_margins(1e-7, 90.0)
    assert st["mismatches"] > 0
