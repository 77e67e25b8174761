"""BASELINE config 3 at its stated shape: 6-DOF paths, 500 gridpoints, JointVelocity + JointAcceleration +
SecondOrderConstraint torque rows (R = 48, nC = 50), SURVEY.md section 8d.

  * reference golden at G = 500 (tests/golden/torque_dof6_g500.npz, 4 paths solved by the unmodified reference with the
    numpy inv_dyn): the oracle fed with the numpy-callback rows reproduces it bit for bit (CPU); the single-path GPU
    API with the numpy callback does too (GPU);
  * 256 paths (seeds 2000+b): the DEVICE-MODEL path of BatchTOPPRA (tb_coeff_second_order: q, q', q'' and the three
    inverse-dynamics terms inside one kernel) against the oracle fed with numpy-callback rows.  The device sin/cos and
    summation order differ from numpy's, so the comparison uses SURVEY section 8d's stated tolerance
    |dK| <= 1e-9 + 1e-8 |K| (same for sd); statuses must match exactly."""
import numpy as np
import pytest

from oracle import oracle as orc
from problems import inv_dyn_numpy, inv_dyn_torch, make_torque_problem

MODEL = ("coupled_cosine", [2.0, 0.3, 0.1, 4.9])   # the closed-form model of tests/problems.py / SURVEY section 8d cfg 3


def oracle_torque_solve(ss, way, vlim, alim, taulim, grid):
    """Rows like the reference builds them: vel+acc rows from the oracle's own K1 restatement, torque rows from the numpy
    inv_dyn (3 calls per gridpoint, linear_second_order.py:142-165) lifted with canlinear_colloc_to_interpolate and
    F = [I; -I], g = [tau_max; -tau_min] (pyx:483-510)."""
    from toppra_b200.constraint.linear_constraint import canlinear_colloc_to_interpolate
    c = orc.cubic_spline_fit(ss, way)
    lin = orc.solve_velacc(c, ss, grid, vlim, alim, True, 0, 0, want_rows=True)
    q, qd, qdd = (orc.ppoly_eval(c, ss, grid, o) for o in (0, 1, 2))
    zero = np.zeros(q.shape[1])
    cv = np.array([inv_dyn_numpy(p, zero, zero) for p in q])
    av = np.array([inv_dyn_numpy(p, zero, ps) for p, ps in zip(q, qd)]) - cv
    bv = np.array([inv_dyn_numpy(p, ps, pss) for p, ps, pss in zip(q, qd, qdd)]) - cv
    dof = q.shape[1]
    F = np.vstack((np.eye(dof), -np.eye(dof)))
    g = np.concatenate((taulim[:, 1], -taulim[:, 0]))
    a2, b2, c2, F2, g2, _, _ = canlinear_colloc_to_interpolate(av, bv, cv, F, g, None, None, grid, identical=True)
    tau_rows = np.stack((a2.dot(F2.T), b2.dot(F2.T), c2.dot(F2.T) - g2), axis=1)   # [G, 3, 4 dof]
    rows = np.concatenate((lin["rows"], tau_rows), axis=2)
    return orc.solve_rows(rows, lin["xbound"], grid, 0.0, 0.0), c


def test_oracle_reproduces_reference_golden_g500(golden):
    g = golden("torque_dof6_g500")
    for b in range(g["way"].shape[0]):
        o, _ = oracle_torque_solve(g["ss"], g["way"][b], g["vlim"][b], g["alim"][b], g["taulim"][b], g["grid"])
        assert o["status"] == g["status"][b] == 0
        assert np.array_equal(o["K"], g["K"][b]) and np.array_equal(o["sd"], g["sd"][b]) and np.array_equal(o["u"], g["sdd"][b])


@pytest.fixture(scope="module")
def ta():
    import toppra_b200
    return toppra_b200


@pytest.mark.gpu
def test_gpu_single_path_numpy_callback_bit_exact_g500(ta, golden):
    g = golden("torque_dof6_g500")
    for b in range(2):
        path = ta.SplineInterpolator(g["ss"], g["way"][b])
        cons = [ta.constraint.JointVelocityConstraint(g["vlim"][b]), ta.constraint.JointAccelerationConstraint(g["alim"][b]),
                ta.constraint.SecondOrderConstraint.joint_torque_constraint(inv_dyn_numpy, g["taulim"][b], np.zeros(6))]
        inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="seidel")
        assert inst.solver_wrapper.nC == 50
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert np.array_equal(K, g["K"][b]) and np.array_equal(sd, g["sd"][b]) and np.array_equal(sdd, g["sdd"][b])


@pytest.mark.gpu
def test_gpu_device_model_batch_vs_oracle_at_cfg3_shape(ta, golden):
    B, G, dof = 256, 500, 6
    ss = np.linspace(0, 1, 5)
    grid = np.linspace(0, 1, G)
    probs = [make_torque_problem(2000 + b) for b in range(B)]
    way, vlim, alim, taulim = (np.stack([p[i] for p in probs]) for i in range(4))
    bpath = ta.BatchSplineInterpolator(ss, way)

    def solve(**kw):
        cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim),
                ta.constraint.SecondOrderConstraint.joint_torque_constraint(kw.pop("inv_dyn", None), taulim, np.zeros(dof), **kw)]
        inst = ta.BatchTOPPRA(cons, bpath, grid)
        assert inst.R == 48 and not inst.fused
        return inst.compute_parameterization(0, 0).to_host()

    h = solve(device_model=MODEL)
    ht = solve(inv_dyn=inv_dyn_torch, batched=True)       # general tensor-callback fallback: same tolerance
    g = golden("torque_dof6_g500")                        # the first 4 paths are the reference golden
    for b in range(4):
        assert h["status"][b] == g["status"][b]
        np.testing.assert_allclose(h["K"][b], g["K"][b], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(h["sd"][b], g["sd"][b], rtol=1e-8, atol=1e-9)
    worst = 0.0
    for b in range(B):
        o, c = oracle_torque_solve(ss, way[b], vlim[b], alim[b], taulim[b], grid)
        assert h["status"][b] == o["status"] == ht["status"][b], b
        if o["status"] != 0:
            continue
        for got in (h, ht):
            np.testing.assert_allclose(got["K"][b], o["K"], rtol=1e-8, atol=1e-9)
            np.testing.assert_allclose(got["sd"][b], o["sd"], rtol=1e-8, atol=1e-9)
        worst = max(worst, float(np.max(np.abs(h["sd"][b] - o["sd"]) / (1 + np.abs(o["sd"])))))
    assert (h["status"] == 0).sum() > 0.9 * B
    print("cfg3 device model vs oracle: worst relative sd deviation %.3g over %d paths" % (worst, B))
