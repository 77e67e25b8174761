"""GPU (-m gpu): parity of the CUDA path against the committed golden vectors (reference outputs) and the
oracle.  Every comparison is BIT-EXACT (fp64, kernels built with -fmad=false) unless a tolerance is written."""
import numpy as np
import pytest

from conftest import BATCH_CASES
from test_oracle_vs_golden import (LP1D_KATS, LP2D_KATS, CPP_K_MAX, CPP_PARAM, CPP_FEAS_MAX)

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.fixture(scope="module")
def ta():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import toppra_b200
    return toppra_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _cons(ta, g, b=None):
    vlim = g["vlim"] if b is None else g["vlim"][b]
    alim = g["alim"] if b is None else g["alim"][b]
    scheme = int(g["scheme"]) if "scheme" in g else 1
    return [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim, scheme)]


# ---- LP layer on device ------------------------------------------------------------------------------------
def test_lp1d_kats(ta):
    for v, a, b, low, high, res, optval, optvar, active in LP1D_KATS:
        r, val, var, act = ta.engine.lp1d_batch(np.array([v], float), np.array([a], float).reshape(1, -1),
                                                np.array([b], float).reshape(1, -1), np.array([low]), np.array([high]))
        assert (r[0], val[0], var[0], act[0]) == (res, optval, optvar, active)
    r, _, _, _ = ta.engine.lp1d_batch(np.array([[1.0, 2]]), np.array([[-1.0, 1.0]]), np.array([[0.0, 0.5]]),
                                      np.array([-1.0]), np.array([1.0]))
    assert r[0] == 0


def test_lp2d_kats(ta):
    for v, a, b, c, low, high, active_c, res, optval, optvar, active in LP2D_KATS:
        n = len(a)
        r, val, var, act = ta.engine.lp2d_batch(np.array([v], float), np.array(a, float).reshape(1, n),
                                                np.array(b, float).reshape(1, n), np.array(c, float).reshape(1, n),
                                                np.array([low], float), np.array([high], float), np.array([active_c]))
        assert r[0] == res
        if res:
            np.testing.assert_allclose(val[0], optval)
            np.testing.assert_allclose(var[0], optvar)
            assert set(act[0].tolist()) == set(active)


def test_lp2d_random100_one_launch(ta, golden):
    g = golden("lp2d_random100")
    low = np.tile(g["low"], (100, 1))
    high = np.tile(g["high"], (100, 1))
    r, val, var, act = ta.engine.lp2d_batch(g["v"], g["a"], g["b"], g["c"], low, high, g["active_in"])
    assert _eq(r, g["res"])
    ok = g["res"] == 1
    assert _eq(val[ok], g["optval"][ok]) and _eq(var[ok], g["optvar"][ok]) and _eq(act[ok], g["active_out"][ok])


def test_lp2d_many_rows_vs_oracle(ta, orc):
    """Rows-per-lane 2..4 (n up to 128) and random warm-start pairs."""
    rng = np.random.RandomState(7)
    for n in (33, 50, 64, 65, 96, 100, 128):
        B = 64
        v = rng.randn(B, 3)
        a, b = rng.randn(2, B, n)
        c = np.where(rng.rand(B, 1) < 0.5, -rng.rand(B, n), rng.randn(B, n) * 0.3 - 0.6)
        low = np.tile([-1.0, -2.0], (B, 1))
        high = np.tile([1.5, 0.7], (B, 1))
        act = rng.randint(-4, n + 2, size=(B, 2))
        r, val, var, aout = ta.engine.lp2d_batch(v, a, b, c, low, high, act)
        for i in range(B):
            r0, val0, var0, act0 = orc.lp2d(v[i], a[i], b[i], c[i], low[i], high[i], act[i])
            assert r[i] == r0
            if r0:
                assert val[i] == val0 and _eq(var[i], var0) and _eq(aout[i], act0)


# ---- spline fit / evaluation -------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 4, 5, 9, 20])
def test_spline_fit_all_boundary_conditions(ta, golden, n):
    g = golden("spline_fits")
    x, y = g["x_%d" % n], g["y_%d" % n]
    cases = {"not-a-knot": "not-a-knot", "clamped": "clamped", "natural": "natural",
             "first": ((1, g["d0_%d" % n]), (1, g["d1_%d" % n])), "mixed": ((2, g["d0_%d" % n]), (1, g["d1_%d" % n]))}
    for key, bc in cases.items():
        path = ta.SplineInterpolator(x, y, bc_type=bc)
        ref = g["c_%d_%s" % (n, key)]
        # tolerance: scipy's LAPACK banded solve; in practice the results are bit-equal
        np.testing.assert_allclose(path.cspl.c, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_spline_interpolator_api(ta, golden):
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    assert path.dof == 7 and path.duration == 1.0 and _eq(path.path_interval, [0, 1])
    assert _eq(path.waypoints[0], g["ss"]) and _eq(path.waypoints[1], g["way"])
    assert _eq(path.cspl.c, g["c"])
    assert _eq(path(g["grid"], 1), g["qs"]) and _eq(path(g["grid"], 2), g["qss"])
    assert _eq(path.evald(g["grid"]), g["qs"]) and _eq(path.cspld(g["grid"]), g["qs"])
    assert path(0.3).shape == (7,) and path([0.3, 0.4], 1).shape == (2, 7)
    assert _eq(path(0.3, 1), path(np.array([0.3]), 1)[0])
    with pytest.raises(ValueError):
        path(0.1, 3)
    # scalar-dof path
    p1 = ta.SplineInterpolator([0, 0.5, 1, 2], [0.0, 1.0, 0.5, 2.0])
    assert p1.dof == 1 and p1([0.1, 0.2]).shape == (2,)
    # single waypoint -> constant path
    p0 = ta.SplineInterpolator([0], [[1.0, 2.0]])
    assert _eq(p0([0, 1.0]), [[1, 2], [1, 2]]) and _eq(p0(0.5, 1), [0, 0])


# ---- constraint parameters (reference 7-tuple contract) ----------------------------------------------------
def test_constraint_params_contract(ta, golden):
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    pc_vel, pc_acc = _cons(ta, g)
    out = pc_vel.compute_constraint_params(path, g["grid"])
    assert all(o is None for o in out[:6]) and _eq(out[6], g["xbound"])
    a, b, c, F, gg, ub, xb = pc_acc.compute_constraint_params(path, g["grid"])
    assert _eq(a, g["acc_a"]) and _eq(b, g["acc_b"]) and _eq(F, g["acc_F"]) and _eq(gg, g["acc_g"])
    assert not c.any() and ub is None and xb is None
    pc_acc.set_discretization_type(0)
    a, b, c, F, gg, _, _ = pc_acc.compute_constraint_params(path, g["grid"])
    assert _eq(a, g["qs"]) and _eq(b, g["qss"]) and F.shape == (14, 7)
    with pytest.raises(ValueError):
        ta.constraint.JointVelocityConstraint(np.ones(3)).compute_constraint_params(path, g["grid"])


# ---- whole algorithm ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", BATCH_CASES)
def test_batch_cases_bit_exact(ta, golden, case):
    """BatchTOPPRA (per-path limits) on the golden batches: fit, records, K, sd, sdd, status."""
    g = golden(case)
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    assert _eq(path.d_ppoly.cpu().numpy(), g["c"])
    inst = ta.BatchTOPPRA(_cons(ta, g), path, g["grid"])
    res = inst.compute_parameterization(float(g["sd_start"]), float(g["sd_end"]))
    h = res.to_host()
    assert _eq(h["status"], g["status"])
    assert _eq(h["K"], g["K"]) and _eq(h["sd"], g["sd"]) and _eq(h["sdd"], g["sdd"])
    R = inst.R
    assert inst.fused and inst.records is None  # vel+acc: rows are built inside the scan, only xbound is materialised
    assert _eq(inst.xbound.cpu().numpy(), np.stack((np.maximum(-1e8, g["xbound"][:, :, 0]),
                                                     np.minimum(1e8, g["xbound"][:, :, 1])), axis=-1))
    # the same problem through materialised stage records (K1 -> K2): identical results, records == reference rows
    inst = ta.BatchTOPPRA(_cons(ta, g), path, g["grid"], fused=False)
    h2 = inst.compute_parameterization(float(g["sd_start"]), float(g["sd_end"])).to_host()
    for key in ("status", "K", "sd", "sdd"):
        assert _eq(h2[key], h[key]), key
    rec = inst.records.cpu().numpy()
    assert _eq(rec[:, :, 3 * R], np.maximum(-1e8, g["xbound"][:, :, 0]))
    assert _eq(rec[:, :, 3 * R + 1], np.minimum(1e8, g["xbound"][:, :, 1]))
    for b in range(rec.shape[0]):
        F = g["acc_F"][b]
        assert _eq(rec[b, :, 0:R], g["acc_a"][b].dot(F.T)) and _eq(rec[b, :, R:2 * R], g["acc_b"][b].dot(F.T))


@pytest.mark.parametrize("case", ["cfg2_seeds1000", "p1_velocity_active", "p3_inadmissible_start", "p6_collocation"])
def test_single_path_api_bit_exact(ta, golden, case):
    """The drop-in classes: TOPPRA(constraints, SplineInterpolator, gridpoints, 'seidel')."""
    g = golden(case)
    for b in range(min(4, g["way"].shape[0])):
        path = ta.SplineInterpolator(g["ss"], g["way"][b])
        inst = ta.algorithm.TOPPRA(_cons(ta, g, b), path, gridpoints=g["grid"], solver_wrapper="seidel")
        sdd, sd, v, K = inst.compute_parameterization(float(g["sd_start"]), float(g["sd_end"]), return_data=True)
        assert _eq(K, g["K"][b])
        code = ta.algorithm.STATUS_CODES[int(g["status"][b])]
        assert inst.problem_data.return_code == code
        if code == ta.algorithm.ParameterizationReturnCode.Ok:
            assert _eq(sd, g["sd"][b]) and _eq(sdd, g["sdd"][b]) and v.shape == (len(g["grid"]) - 1, 0)
            assert _eq(inst.problem_data.sd_vec, g["sd"][b]) and _eq(inst.problem_data.K, g["K"][b])
        else:
            assert sd is None and sdd is None and v is None


def test_cfg1_example_and_trajectory(ta, golden):
    """BASELINE config 1 (examples/plot_kinematics.py): feasible/controllable sets, auto gridpoints, trajectory."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    inst = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"])
    assert _eq(inst.compute_feasible_sets(), g["X"])
    assert _eq(inst.compute_controllable_sets(0.0, 0.5), g["K_0_05"])
    traj = inst.compute_trajectory(0, 0)
    assert inst.problem_data.return_code == ta.algorithm.ParameterizationReturnCode.Ok
    assert _eq(inst.problem_data.sd_vec, g["sd"])
    # output trajectory: tolerance 1e-9 — time stamps are a host prefix sum, the re-fit runs on
    # the GPU; scipy's banded solve vs ours differ at rounding level only
    np.testing.assert_allclose(traj.duration, g["traj_duration"], rtol=1e-12)
    np.testing.assert_allclose(traj(g["traj_ts"]), g["traj_q"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(traj(g["traj_ts"], 1), g["traj_qd"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(traj(g["traj_ts"], 2), g["traj_qdd"], rtol=1e-7, atol=1e-6)
    # automatic gridpoints (propose_gridpoints): identical grid, then identical solution
    auto = ta.algorithm.TOPPRA(_cons(ta, g), path)
    assert _eq(auto.gridpoints, g["auto_grid"])
    _, sd, _, K = auto.compute_parameterization(0, 0, return_data=True)
    assert _eq(sd, g["auto_sd"]) and _eq(K, g["auto_K"])
    # const-accel parametrizer
    inst2 = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"], parametrizer="ParametrizeConstAccel")
    traj2 = inst2.compute_trajectory(0, 0)
    assert abs(traj2.duration - g["traj_duration"]) < 1e-9
    qd = traj2(np.linspace(0, traj2.duration, 200), 1)
    assert np.all(np.abs(qd) <= g["vlim"][:, 1] * (1 + 1e-6))


def test_cpp_2dof_collocation_golden(ta, golden):
    g = golden("cpp_2dof_collocation")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    cons = [ta.constraint.JointVelocityConstraint([1.0, 1.0]),
            ta.constraint.JointAccelerationConstraint([0.2, 0.2], discretization_scheme=0)]
    inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="seidel")
    sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
    assert _eq(K, g["K"]) and _eq(sd, g["sd"]) and _eq(sdd, g["sdd"])
    np.testing.assert_allclose(K[:, 1], CPP_K_MAX, atol=1e-6)       # cpp/tests/test_algorithm.cpp:109-117
    np.testing.assert_allclose(sd ** 2, CPP_PARAM, atol=1e-6)       # :132-140
    X = inst.compute_feasible_sets()
    assert _eq(X, g["X"])
    np.testing.assert_allclose(X[:, 1], CPP_FEAS_MAX, atol=1e-6)    # :161-169


def test_stagewise_plugin_interface(ta, golden):
    """solve_stagewise_optim of the solver wrapper (the reference's plugin boundary), warm start chained."""
    g = golden("stagewise_6dof")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    cons = [ta.constraint.JointVelocityConstraint(g["vlim"]), ta.constraint.JointAccelerationConstraint(g["alim"])]
    w = ta.solverwrapper.seidelWrapper(cons, path, g["grid"], solve_lp1d=1)   # as the golden was generated
    assert w.get_no_vars() == 2 and w.get_no_stages() == 200 and len(w.get_deltas()) == 200
    for row in g["cases"]:
        i, gg, xb, xnb, ref = int(row[0]), row[1:3], row[3:5], row[5:7], row[7:9]
        res = w.solve_stagewise_optim(i, None, gg, xb[0], xb[1], xnb[0], xnb[1])
        assert _eq(res, ref), (row, res)
    assert len(w.params) == 2 and w.params[0][6].shape == (201, 2)


def test_robustness_suite(ta, golden):
    g = golden("p4_robustness_suite")
    for name in g["names"]:
        path = ta.SplineInterpolator(g[name + "_ss"], g[name + "_way"], bc_type="clamped")
        np.testing.assert_allclose(path.cspl.c, g[name + "_c"], rtol=1e-12, atol=1e-15)
        if not _eq(path.cspl.c, g[name + "_c"]):
            continue  # a last-bit difference in the fit changes the degenerate LPs; compared via the oracle below
        cons = [ta.constraint.JointVelocityConstraint(g[name + "_vlim"]),
                ta.constraint.JointAccelerationConstraint(g[name + "_alim"])]
        inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g[name + "_grid"], solver_wrapper="seidel")
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert inst.problem_data.return_code == ta.algorithm.STATUS_CODES[int(g[name + "_status"])], name
        assert _eq(K, g[name + "_K"]), name
        if sd is not None:
            assert _eq(sd, g[name + "_sd"]) and _eq(sdd, g[name + "_sdd"]), name


def test_torque_second_order(ta, golden):
    """cfg-3 shape: vel + acc + SecondOrderConstraint.joint_torque_constraint.
    Reference-style numpy callback: bit-exact.  Batched tensor callback: rtol 1e-9 (torch's matmul/cos order)."""
    from problems import inv_dyn_numpy, inv_dyn_torch
    g = golden("torque_dof6")
    B = g["way"].shape[0]
    for b in range(B):
        path = ta.SplineInterpolator(g["ss"], g["way"][b])
        cons = [ta.constraint.JointVelocityConstraint(g["vlim"][b]), ta.constraint.JointAccelerationConstraint(g["alim"][b]),
                ta.constraint.SecondOrderConstraint.joint_torque_constraint(inv_dyn_numpy, g["taulim"][b], np.zeros(6))]
        inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="seidel")
        assert inst.solver_wrapper.nC == 50
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        assert _eq(K, g["K"][b]) and _eq(sd, g["sd"][b]) and _eq(sdd, g["sdd"][b])
        a, bb, c, F, gg, _, _ = cons[2].compute_constraint_params(path, g["grid"])
        assert _eq(a, g["tau_a"][b]) and _eq(bb, g["tau_b"][b]) and _eq(c, g["tau_c"][b])
    bpath = ta.BatchSplineInterpolator(g["ss"], g["way"])
    cons = [ta.constraint.JointVelocityConstraint(g["vlim"]), ta.constraint.JointAccelerationConstraint(g["alim"]),
            ta.constraint.SecondOrderConstraint.joint_torque_constraint(inv_dyn_torch, g["taulim"], np.zeros(6),
                                                                        batched=True)]
    h = ta.BatchTOPPRA(cons, bpath, g["grid"]).compute_parameterization(0, 0).to_host()
    assert _eq(h["status"], g["status"])
    np.testing.assert_allclose(h["K"], g["K"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(h["sd"], g["sd"], rtol=1e-9, atol=1e-12)


def test_joint_torque_constraint(ta, golden):
    """JointTorqueConstraint (reference joint_torque.py:7-116): vel + torque with dry friction, Collocation
    (the reference's default for this class) and Interpolation, bit-exact incl. the host 7-tuple."""
    from problems import inv_dyn_numpy
    g = golden("joint_torque_dof6")
    for scheme in (0, 1):
        t = "s%d_" % scheme
        for b in range(g[t + "way"].shape[0]):
            path = ta.SplineInterpolator(g["ss"], g[t + "way"][b])
            pc_tau = ta.constraint.JointTorqueConstraint(inv_dyn_numpy, g[t + "taulim"][b], g[t + "fric"][b],
                                                         discretization_scheme=scheme)
            assert pc_tau.identical and pc_tau.get_dof() == 6
            cons = [ta.constraint.JointVelocityConstraint(g[t + "vlim"][b]), pc_tau]
            inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="seidel")
            assert inst.solver_wrapper.nC == 2 + (12 if scheme == 0 else 24)
            sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
            assert _eq(K, g[t + "K"][b]) and _eq(sd, g[t + "sd"][b]) and _eq(sdd, g[t + "sdd"][b]), (scheme, b)
            a, bb, c, F, gg, ub, xb = pc_tau.compute_constraint_params(path, g["grid"])
            assert ub is None and xb is None and _eq(F, g[t + "F"][b]) and _eq(gg, g[t + "g"][b])
            assert _eq(a, g[t + "a"][b]) and _eq(bb, g[t + "b"][b]) and _eq(c, g[t + "c"][b]), (scheme, b)
    with pytest.raises(ValueError):
        ta.constraint.JointTorqueConstraint(inv_dyn_numpy, np.ones((5, 2)), np.zeros(5)).compute_constraint_params(
            path, g["grid"])


def test_cartesian_velocity_norm(ta, orc, golden):
    """CartesianVelocityNorm (C++-only in the reference: cpp/src/toppra/constraint/cartesian_velocity_norm.cpp:23-54,
    SURVEY §8 f4): one row (0, v^T S v, -limit) per gridpoint.  Constant and varying limit; checked against the oracle
    fed with the same rows (bit-exact) and through the constraint's meaning: v^T S v * sd^2 <= limit everywhere."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    grid = g["grid"]
    rng = np.random.RandomState(3)
    J0 = rng.randn(6, 7)
    S = np.diag([1.0, 1.0, 1.0, 0.1, 0.1, 0.1])

    def frame_velocity(q, qd):
        return (J0 + 0.2 * np.sin(q)[None, :]).dot(qd)

    def varying(s):
        return S * (1.0 + 0.5 * s), 0.6 + 0.4 * np.cos(3.0 * s) ** 2

    qs, qds = path(grid), path(grid, 1)
    for cart in (ta.constraint.CartesianVelocityNorm(frame_velocity, S, 0.8, dof=7),
                 ta.constraint.CartesianVelocityNorm(frame_velocity, velocity_limit=varying)):
        a, b, c, F, gg, ub, xb = cart.compute_constraint_params(path, grid)
        lim = np.array([0.8 if cart.identical else varying(s)[1] for s in grid])
        Ss = [S if cart.identical else varying(s)[0] for s in grid]
        want_b = np.array([frame_velocity(q, qd).dot(Si.dot(frame_velocity(q, qd))) for q, qd, Si in zip(qs, qds, Ss)])
        assert not a.any() and not c.any() and ub is None and xb is None and _eq(b[:, 0], want_b)
        assert (F.shape, gg.shape) == (((1, 1), (1,)) if cart.identical else ((len(grid), 1, 1), (len(grid), 1)))
        cons = [ta.constraint.JointVelocityConstraint(g["vlim"]), cart]
        inst = ta.algorithm.TOPPRA(cons, path, gridpoints=grid, solver_wrapper="seidel")
        assert inst.solver_wrapper.nC == 3
        sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
        rows = np.stack((np.zeros_like(b), b, -lim[:, None]), axis=1)
        o = orc.solve_rows(rows, np.stack((g["xbound"][:, 0], np.minimum(g["xbound"][:, 1], 1e8)), axis=1), grid, 0, 0)
        assert o["status"] == 0 and inst.problem_data.return_code == ta.algorithm.ParameterizationReturnCode.Ok
        assert _eq(K, o["K"]) and _eq(sd, o["sd"]) and _eq(sdd, o["u"])
        used = want_b * sd ** 2 / lim
        assert used.max() <= 1 + 1e-9 and used.max() > 0.999   # never exceeded, active somewhere
    with pytest.raises(ValueError):
        ta.constraint.CartesianVelocityNorm(frame_velocity, np.eye(5), 1.0)
    with pytest.raises(ValueError):
        ta.constraint.CartesianVelocityNorm(frame_velocity, S, -1.0)


def test_errors(ta, golden):
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    cons = _cons(ta, g)
    with pytest.raises(ValueError):
        ta.algorithm.TOPPRA(cons, path, gridpoints=np.linspace(0, 0.9, 10))      # algorithm.py:109-113
    with pytest.raises(ValueError):
        ta.algorithm.TOPPRA(cons, path, gridpoints=[0, 0.5, 0.4, 1.0])           # algorithm.py:117-120
    inst = ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"])
    with pytest.raises(ta.exceptions.BadInputVelocities):
        inst.compute_parameterization(-1, 0)                                      # reachability_algorithm.py:272-276
    with pytest.raises(AssertionError):
        ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="gurobi")
    with pytest.raises(NotImplementedError):
        ta.algorithm.TOPPRA(cons, path, gridpoints=g["grid"], solver_wrapper="qpoases")
    assert inst.compute_trajectory(20.0, 0) is None                               # inadmissible -> None
    assert inst.problem_data.return_code == ta.algorithm.ParameterizationReturnCode.FailUncontrollable


def test_host_buffer_cabi_entry(ta, golden):
    """tb_solve_velacc_host: the torch-free C entry (K0 -> K1 -> K2 + copies) with plain host buffers."""
    g = golden("cfg2_seeds1000")
    out = ta.engine.solve_velacc_host(g["ss"], g["way"], g["grid"], g["vlim"], g["alim"], True)
    assert _eq(out["status"], g["status"]) and _eq(out["K"], g["K"]) and _eq(out["sd"], g["sd"]) and _eq(out["u"], g["sdd"])
    out = ta.engine.solve_velacc_host(g["ss"], g["way"], g["grid"], g["vlim"][0], g["alim"][0], True)
    assert _eq(out["K"][0], g["K"][0])


def test_custom_linear_constraint_generic_rows(ta, orc, golden):
    """A user-defined CanonicalLinear constraint with a NON-identical F (like the reference's test constraint in
    tests/tests/solverwrapper/test_basic_can_linear.py:18-50): rows F_i a_i, F_i b_i, F_i c_i - g_i are assembled on
    the device from the host 7-tuple.  Checked against the oracle fed with numpy-assembled rows (tolerance 1e-12:
    the dot-product order of numpy's BLAS is not pinned)."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    grid = g["grid"]
    N, dof = len(grid) - 1, 7
    rng = np.random.RandomState(0)

    class Custom(ta.constraint.LinearConstraint):
        def __init__(self):
            super().__init__()
            self.dof = dof
            self.identical = False
            self._format_string = ""

        def get_dof(self):
            return dof

        def compute_constraint_params(self, path_, gridpoints):
            qs, qss = path_(gridpoints, 1), path_(gridpoints, 2)
            G = len(gridpoints)
            F = np.tile(np.vstack((np.eye(dof), -np.eye(dof)))[None], (G, 1, 1)) * (1 + 0.1 * rng.rand(G, 1, 1))
            gg = np.tile(np.r_[g["alim"][:, 1], -g["alim"][:, 0]][None], (G, 1)) * 1.5
            xb = np.stack((np.zeros(G), np.full(G, 0.9)), axis=1)
            return qs, qss, 0.01 * np.ones_like(qs), F, gg, None, xb

    cons = [ta.constraint.JointVelocityConstraint(g["vlim"]), Custom()]
    inst = ta.algorithm.TOPPRA(cons, path, gridpoints=grid, solver_wrapper="seidel")
    sdd, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
    a, b, c, F, gg, _, xb = cons[1]._host_params(inst.solver_wrapper.ctx)
    rows = np.stack((np.einsum("gkm,gm->gk", F, a), np.einsum("gkm,gm->gk", F, b), np.einsum("gkm,gm->gk", F, c) - gg), axis=1)
    xbound = np.stack((np.maximum(g["xbound"][:, 0], xb[:, 0]), np.minimum(np.minimum(g["xbound"][:, 1], 1e8), xb[:, 1])), axis=1)
    o = orc.solve_rows(rows, xbound, grid, 0, 0)
    assert o["status"] == 0 and inst.problem_data.return_code == ta.algorithm.ParameterizationReturnCode.Ok
    np.testing.assert_allclose(K, o["K"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sd, o["sd"], rtol=1e-12, atol=1e-14)
    assert np.all(sd ** 2 <= 0.9 + 1e-12)


def test_const_accel_parametrizer_on_device(ta, golden):
    """K3: time stamps bit-exact (sequential sum like the reference); q/qd/qdd at sample times."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    inst = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"], parametrizer="ParametrizeConstAccel")
    traj = inst.compute_trajectory(0, 0)
    assert _eq(traj._ts, g["ca_tgrid"]) and _eq(traj._us, g["ca_us"]) and traj.duration == float(g["ca_duration"])
    # evaluation: same formulas, the reference evaluates q(s) with scipy at s computed in numpy -> rounding-level
    np.testing.assert_allclose(traj(g["ca_ts"]), g["ca_q"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(traj(g["ca_ts"], 1), g["ca_qd"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(traj(g["ca_ts"], 2), g["ca_qdd"], rtol=1e-10, atol=1e-10)
    assert traj(0.5).shape == (7,) and traj.dof == 7
    # batched form: durations of a whole batch in one launch
    gb = golden("cfg2_seeds1000")
    bpath = ta.BatchSplineInterpolator(gb["ss"], gb["way"])
    res = ta.BatchTOPPRA(_cons(ta, gb), bpath, gb["grid"]).compute_parameterization(0, 0)
    bp = ta.BatchParametrizeConstAccel(bpath, gb["grid"], res.sd)
    dur = bp.durations.cpu().numpy()
    ref = np.array([np.sum(2 * np.diff(gb["grid"]) / (gb["sd"][b][1:] + gb["sd"][b][:-1])) for b in range(len(dur))])
    np.testing.assert_allclose(dur, ref, rtol=1e-13)
    q_end = bp(np.stack([np.array([0.0, d]) for d in dur]), 0).cpu().numpy()
    np.testing.assert_allclose(q_end[:, 0], gb["way"][:, 0], atol=1e-12)
    np.testing.assert_allclose(q_end[:, 1], gb["way"][:, -1], atol=1e-9)


def test_toppra_sd_and_reachable_sets(ta, golden):
    """TOPPRAsd (fastest/slowest passes on the GPU + the reference's bisection) and
    compute_reachable_sets (per-stage LPs through solve_stagewise_optim), bit-exact vs the reference."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    inst = ta.algorithm.TOPPRAsd(_cons(ta, g), path, gridpoints=g["grid"], solver_wrapper="seidel")
    for tag, dur in (("sd5", 5.0), ("sd_fast", 1.0), ("sd_slow", 1e9)):
        inst.set_desired_duration(dur)
        sdd, sd, v, K = inst.compute_parameterization(0, 0, return_data=True)
        assert _eq(sd, g[tag + "_sd"]) and _eq(sdd, g[tag + "_sdd"]) and _eq(K, g["K"]), tag
    inst.set_desired_duration(5.0)
    traj = inst.compute_trajectory(0, 0)
    assert abs(traj.duration - 5.0) < 1e-3
    L = ta.algorithm.TOPPRA(_cons(ta, g), path, gridpoints=g["grid"]).compute_reachable_sets(0.0, 0.5)
    assert _eq(L, g["L_0_05"])


def test_velocity_constraint_varying(ta, golden):
    """JointVelocityConstraintVarying (limits as a function of s), bit-exact vs the reference."""
    g = golden("cfg1_seed9")
    path = ta.SplineInterpolator(g["ss"], g["way"])
    vlim = g["vlim"]
    pc_var = ta.constraint.JointVelocityConstraintVarying(lambda s: vlim * (0.05 + 0.5 * s))
    assert pc_var.get_dof() == 7
    assert _eq(pc_var.compute_constraint_params(path, g["grid"])[-1], g["var_xbound"])
    inst = ta.algorithm.TOPPRA([pc_var, ta.constraint.JointAccelerationConstraint(g["alim"])], path,
                               gridpoints=g["grid"], solver_wrapper="seidel")
    _, sd, _, K = inst.compute_parameterization(0, 0, return_data=True)
    assert _eq(sd, g["var_sd"]) and _eq(K, g["var_K"])


def test_forward_retry_rule(ta, golden):
    """The cold retry branch of the forward pass (reachability_algorithm.py:315-343) against reference outputs."""
    import torch
    g = golden("retry_after_slack_start")
    path = ta.BatchSplineInterpolator(g["ss"], g["way"])
    inst = ta.BatchTOPPRA([ta.constraint.JointVelocityConstraint(g["vlim"]),
                           ta.constraint.JointAccelerationConstraint(g["alim"])], path, g["grid"])
    res = inst.compute_parameterization(g["sd_start"], 0.0, counters=True)
    h = res.to_host()
    assert _eq(h["status"], g["status"]) and set(g["status"].tolist()) == {0, 1}
    assert _eq(h["K"], g["K"]) and _eq(h["sd"], g["sd"]) and _eq(h["sdd"], g["sdd"])
    assert (res.counters[:, 3].cpu().numpy() > 0).all()           # every path went through the retry rule
    r = golden("retry_row_problems")
    for i in range(int(r["n"])):
        t = "c%d_" % i
        rows, xb, grid = r[t + "rows"], r[t + "xb"], r[t + "grid"]
        G, _, R = rows.shape
        W = ta.engine.record_doubles(R)
        rec = np.zeros((1, G, W))
        rec[0, :, :3 * R] = rows.reshape(G, 3 * R)
        rec[0, :, 3 * R:3 * R + 2] = xb
        dev = ta.engine.default_device()
        out = ta.engine.scan(ta.engine.as_device(rec, dev), R, ta.engine.as_device(grid, dev),
                             ta.engine.as_device(np.array([float(r[t + "sd_start"])]), dev), None)
        assert int(out["status"][0]) == int(r[t + "status"]), i
        assert _eq(out["K"][0].cpu().numpy(), r[t + "K"]) and _eq(out["sd"][0].cpu().numpy(), r[t + "sd"]), i
        assert _eq(out["u"][0].cpu().numpy(), r[t + "sdd"]), i
