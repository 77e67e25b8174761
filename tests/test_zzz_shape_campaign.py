"""GPU (-m gpu; the harness itself is replayed on the engine double under -m "not gpu"): the kernels against the oracle on
RANDOMLY SHAPED problems — the shapes of scripts/fuzz_oracle_vs_reference.py, whose CPU campaign pins the oracle to the
unmodified reference on 55 000 such problems.  dof 1..14 (one, two rows per lane), 2..12 knots on non-uniform breakpoints,
2..400 gridpoints (non-uniform, on breakpoints), every boundary condition, both discretisations, active velocity bounds,
non-zero and inadmissible boundary velocities, tiny motions.  Each problem goes through the single-path API (K1 records +
record scan, feasible sets) and through BatchTOPPRA (fused scan where the shape allows it).  Bit for bit.

Runs last (file name) so that a shape nobody thought of cannot hide the rest of the suite behind `-x`."""
import numpy as np
import pytest

import cpu_engine
from oracle import oracle as orc
from problems import random_shaped_problem

N_PROBLEMS = 120


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "cpu_double"])
def ta(request, monkeypatch):
    if request.param == "cpu_double":
        return cpu_engine.install(monkeypatch)
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import toppra_b200
    return toppra_b200


def _same(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_kernels_equal_the_oracle_on_randomly_shaped_problems(ta):
    codes = {"Ok": 0, "ErrUnknown": 1, "ErrShortPath": 2, "FailUncontrollable": 3, "ErrForwardPassFail": 4}
    seen_status, fused_runs, two_rows_per_lane = set(), 0, 0
    for seed in range(500000, 500000 + N_PROBLEMS):
        p = random_shaped_problem(np.random.RandomState(seed))
        tag = "seed %d (dof %d, %d knots, %d gridpoints, %s, scheme %d)" % (
            seed, p["way"].shape[1], len(p["ss"]), len(p["grid"]), p["bc"], p["interp"])
        path = ta.SplineInterpolator(p["ss"], p["way"], bc_type=p["bc"])
        c = np.ascontiguousarray(path.cspl.c)
        fit = orc.cubic_spline_fit(p["ss"], p["way"], p["bc"])
        if p["bc"] == "not-a-knot":
            assert _same(c, fit), "spline fit, " + tag
        else:   # same algebra on both sides; the tolerance only guards the comparison against the order of a future refit
            np.testing.assert_allclose(c, fit, rtol=1e-11, atol=1e-12 * max(1.0, np.abs(fit).max()), err_msg=tag)
        cons = [ta.constraint.JointVelocityConstraint(p["vlim"]),
                ta.constraint.JointAccelerationConstraint(p["alim"], discretization_scheme=p["interp"])]
        inst = ta.algorithm.TOPPRA(cons, path, gridpoints=p["grid"], solver_wrapper="seidel")
        two_rows_per_lane += inst.solver_wrapper.nC > 32
        sdd, sd, _, K = inst.compute_parameterization(p["sd0"], p["sd1"], return_data=True)
        o = orc.solve_velacc(c, p["ss"], p["grid"], p["vlim"], p["alim"], bool(p["interp"]), p["sd0"], p["sd1"])
        assert _same(K, o["K"]), "K, " + tag
        assert codes[inst.problem_data.return_code.name] == o["status"], "status, " + tag
        seen_status.add(o["status"])
        if sd is not None:
            assert _same(sd, o["sd"]) and _same(sdd, o["u"]), "sd / u, " + tag
        lin = orc.solve_velacc(c, p["ss"], p["grid"], p["vlim"], p["alim"], bool(p["interp"]), 0, 0, want_rows=True)
        X = ta.algorithm.TOPPRA(cons, path, gridpoints=p["grid"], solver_wrapper="seidel").compute_feasible_sets()
        assert _same(X, orc.Wrapper(p["grid"], lin["rows"], lin["xbound"]).compute_feasible_sets()), "feasible sets, " + tag
        # the batched entry: three copies of the path in one launch (the fused vel+acc scan wherever it applies)
        bpath = ta.BatchSplineInterpolator.from_ppoly(p["ss"], np.repeat(c[None], 3, axis=0))
        binst = ta.BatchTOPPRA(cons, bpath, gridpoints=p["grid"])
        fused_runs += bool(binst.fused)
        h = binst.compute_parameterization(p["sd0"], p["sd1"]).to_host()
        for b in range(3):
            assert h["status"][b] == o["status"] and _same(h["K"][b], o["K"]), "batched K / status, " + tag
            if o["status"] == 0:
                assert _same(h["sd"][b], o["sd"]) and _same(h["sdd"][b], o["u"]), "batched sd / u, " + tag
    assert {0, 3} <= seen_status and fused_runs >= 40 and two_rows_per_lane >= 5
