"""CPU (-m "not gpu"): the Python host side end to end, with the kernels replaced by the oracle-backed test double
tests/cpu_engine.py.  The GPU parity tests of tests/test_gpu_parity.py are replayed unchanged (same assertions against
the reference's golden vectors), so a regression in the host logic — constraint classes, record assembly, solver
wrapper, algorithms, chunking, error behaviour — shows up without a GPU.  What the numbers prove here is only that the
host code feeds and reads the engine correctly; the CUDA kernels themselves are checked by the -m gpu run."""
import inspect

import numpy as np
import pytest

import cpu_engine
import test_gpu_parity as T
import test_gpu_scale as S
import test_robust as Rb
import test_zz_ppoly_in as P
from conftest import BATCH_CASES  # noqa: F401

REPLAYED = [
    "test_lp1d_kats", "test_lp2d_kats", "test_lp2d_random100_one_launch", "test_lp2d_many_rows_vs_oracle",
    "test_spline_fit_all_boundary_conditions", "test_spline_interpolator_api", "test_constraint_params_contract",
    "test_batch_cases_bit_exact", "test_single_path_api_bit_exact", "test_cpp_2dof_collocation_golden",
    "test_stagewise_plugin_interface", "test_robustness_suite", "test_torque_second_order",
    "test_joint_torque_constraint", "test_cartesian_velocity_norm", "test_errors",
    "test_custom_linear_constraint_generic_rows", "test_forward_retry_rule", "test_const_accel_parametrizer_on_device",
    "test_toppra_sd_and_reachable_sets", "test_velocity_constraint_varying", "test_cfg1_example_and_trajectory",
]


def _expand(fn):
    """[(id, kwargs)] from the function's own @pytest.mark.parametrize marks."""
    cases = [("", {})]
    for mark in getattr(fn, "pytestmark", []):
        if mark.name != "parametrize":
            continue
        names = [n.strip() for n in mark.args[0].split(",")]
        new = []
        for ident, kw in cases:
            for values in mark.args[1]:
                values = values if len(names) > 1 else (values,)
                new.append((ident + "-" + "-".join(str(v) for v in values), dict(kw, **dict(zip(names, values)))))
        cases = new
    return cases


CASES = [(T, name, ident, kw) for name in REPLAYED for ident, kw in _expand(getattr(T, name))]
CASES += [(P, name, "", {}) for name in ("test_ppoly_path_single", "test_scalar_and_low_degree_pieces", "test_batch_from_ppoly",
                                             "test_simple_path_and_polynomial_path", "test_univariate_spline_interpolator")]
for _name in ("test_gpu_zero_ellipsoid_equals_lp_path", "test_gpu_robust_coefficients", "test_gpu_toppra_conic_api"):
    CASES += [(Rb, _name, ident, kw) for ident, kw in _expand(getattr(Rb, _name))]
CASES += [(S, "test_shapes_rows_per_lane_and_tiny_grids", ident, kw)
          for ident, kw in _expand(S.test_shapes_rows_per_lane_and_tiny_grids)]


@pytest.mark.parametrize("module,name,ident,kw", CASES, ids=[n + i for _, n, i, _ in CASES])
def test_replay_gpu_parity_test_on_cpu(module, name, ident, kw, monkeypatch, golden):
    ta = cpu_engine.install(monkeypatch)
    fn = getattr(module, name)
    from oracle import oracle
    avail = dict(ta=ta, golden=golden, orc=oracle, **kw)
    fn(**{p: avail[p] for p in inspect.signature(fn).parameters})


def test_chunked_batch_equals_single_launch_on_cpu(monkeypatch):
    """BatchTOPPRA chunking (max_record_bytes) and BatchSplineInterpolator.chunk: host logic only."""
    ta = cpu_engine.install(monkeypatch)
    from problems import make_batch
    B, G = 12, 40
    ss, way, vlim, alim = make_batch(B, 1000)
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    cons = [ta.constraint.JointVelocityConstraint(vlim), ta.constraint.JointAccelerationConstraint(alim)]
    one = ta.BatchTOPPRA(cons, path, grid).compute_parameterization(0.0, 0.0).to_host()
    W = ta.engine.record_doubles(28)
    many = ta.BatchTOPPRA(cons, path, grid, max_record_bytes=5 * G * W * 8, fused=False).compute_parameterization(0.0, 0.0).to_host()
    for k in ("K", "sd", "sdd", "status"):
        assert np.array_equal(one[k], many[k]), k
    assert not one["status"].any()


def test_batch_input_validation_on_cpu(monkeypatch):
    """Shapes and value ranges are checked on the host before raw pointers reach the kernels: a short limit
    array, a break array of another batch, a non-increasing knot vector, gridpoints that do not span the path interval
    and negative boundary velocities raise instead of reading out of bounds / silently extrapolating."""
    import torch
    ta = cpu_engine.install(monkeypatch)
    from problems import make_batch
    B, G = 6, 30
    ss, way, vlim, alim = make_batch(B, 1000)
    grid = np.linspace(0, 1, G)
    path = ta.BatchSplineInterpolator(ss, way)
    acc = ta.constraint.JointAccelerationConstraint(alim)
    with pytest.raises(ValueError):  # limits of a smaller batch
        ta.BatchTOPPRA([ta.constraint.JointVelocityConstraint(vlim[:B - 1]), acc], path, grid).compute_parameterization(0, 0)
    with pytest.raises(ValueError):  # wrong dof
        ta.BatchTOPPRA([ta.constraint.JointVelocityConstraint(vlim[:, :5]), acc], path, grid).compute_parameterization(0, 0)
    with pytest.raises(ValueError):  # per-path knots of another batch size
        ta.BatchSplineInterpolator(np.tile(ss, (B - 1, 1)), way)
    bad_ss = ss.copy()
    bad_ss[2] = bad_ss[1]
    with pytest.raises(ValueError):
        ta.BatchSplineInterpolator(bad_ss, way)
    with pytest.raises(ValueError):
        ta.BatchSplineInterpolator(torch.as_tensor(bad_ss), torch.as_tensor(way))
    cons = [ta.constraint.JointVelocityConstraint(vlim), acc]
    with pytest.raises(ValueError, match="Invalid manually supplied gridpoints"):
        ta.BatchTOPPRA(cons, path, np.linspace(0, 0.9, G))
    with pytest.raises(ValueError, match="Invalid manually supplied gridpoints"):
        ta.BatchTOPPRA(cons, path, torch.linspace(0.1, 1, G, dtype=torch.float64))
    with pytest.raises(ValueError, match="Bad input gridpoints"):
        g2 = grid.copy()
        g2[3] = g2[2]
        ta.BatchTOPPRA(cons, path, torch.as_tensor(g2))
    inst = ta.BatchTOPPRA(cons, path, grid)
    with pytest.raises(ta.exceptions.BadInputVelocities):
        inst.compute_parameterization(torch.full((B,), -0.1, dtype=torch.float64), 0.0)
    with pytest.raises(ValueError):
        inst.compute_parameterization(torch.zeros(B - 1, dtype=torch.float64), 0.0)
    with pytest.raises(ta.exceptions.BadInputVelocities):
        inst.compute_parameterization(-0.1, 0.0)


def test_plot_helpers_draw_through_matplotlib(monkeypatch):
    """`inspect()` (algorithm.py:196-213) and `plot_parametrization()` (parametrizer.py:131-158): matplotlib is imported
    lazily; a recording stand-in shows that the panels are drawn from the solve's data."""
    import sys
    import types
    calls = []

    class Rec(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return Rec(name)

        def __call__(self, *args, **kwargs):
            calls.append((self.__name__, args))
            return Rec("result")

    plt = Rec("pyplot")
    root = Rec("matplotlib")
    root.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", root)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    ta = cpu_engine.install(monkeypatch)
    path = ta.SplineInterpolator([0, 1, 2], [(0, 0), (1, 2), (2, 0)])
    traj = ta.ParametrizeConstAccel(path, [0, 0.5, 1, 1.5, 2], [1, 2, 2, 1, 0])
    traj.plot_parametrization(show=True)
    names = [n for n, _ in calls]
    assert names.count("subplot") == 4 and names.count("plot") == 6 and names[-1] == "show"
    s_of_t = [a for n, a in calls if n == "plot"][0]
    assert np.all(np.diff(s_of_t[1]) >= 0) and abs(s_of_t[1][-1] - 2.0) < 1e-12      # s(t) runs to the path end
    lim = np.array([[-1.0, 1.0], [-1.0, 1.0]])
    inst = ta.algorithm.TOPPRA([ta.constraint.JointVelocityConstraint(lim), ta.constraint.JointAccelerationConstraint(lim)], path)
    inst.compute_feasible_sets()
    inst.compute_trajectory(0, 0)
    del calls[:]
    inst.inspect()
    assert [n for n, _ in calls].count("plot") == 5                                    # X (2), K (2), sd^2
