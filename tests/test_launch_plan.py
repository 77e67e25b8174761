"""CPU: which C-ABI calls a solve makes — the LAUNCH PLAN of every BASELINE configuration, pinned without a GPU.

The product's host code (BatchSplineInterpolator, BatchTOPPRA, BatchTOPPRAsd, constraint classes, chunking) runs unchanged
on CPU tensors against the recording library of tests/test_engine_marshalling.py: no kernel runs and the outputs are
garbage, but the SEQUENCE of entry points, their batch sizes, flags and pointer offsets are exactly what a GPU run issues.
This pins, among others, bench.py's `gpu_launches` claim (3 launches per cfg-2 step), the absence of stage records on the
fused path, and the chunk arithmetic of large record batches."""
import ctypes

import numpy as np

import toppra_b200 as ta
from toppra_b200 import engine
from test_engine_marshalling import lib, named_args  # noqa: F401  (lib: fixture)

B, N, DOF, G = 8, 5, 6, 50
SIZE_QUERIES = ("tb_record_doubles", "tb_spline_fit_workspace_doubles")


def _launches(lib, start=0):  # noqa: F811
    return [name for name, _ in lib.calls[start:] if name not in SIZE_QUERIES]


def _last(lib, name):  # noqa: F811
    """The last call of `name` as {header parameter name: value}."""
    return named_args(name, lib.last(name))


def _all(lib, name, start):  # noqa: F811
    return [named_args(n, a) for n, a in lib.calls[start:] if n == name]


def _problem(dof=DOF):
    rng = np.random.RandomState(0)
    path = ta.BatchSplineInterpolator(np.linspace(0, 1, N), rng.randn(B, N, dof))
    lim = np.tile(np.array([[-1.0, 1.0]]), (dof, 1))
    return path, ta.constraint.JointVelocityConstraint(lim), ta.constraint.JointAccelerationConstraint(lim)


def test_cfg2_step_is_three_launches_and_builds_no_records(lib):  # noqa: F811
    """bench.py's timed step: fit + velocity bound + fused scan (gpu_launches = 3 per step)."""
    path, vel, acc = _problem(7)
    assert _launches(lib) == ["tb_spline_fit"]
    mark = len(lib.calls)
    inst = ta.BatchTOPPRA([vel, acc], path, gridpoints=np.linspace(0, 1, G))
    res = inst.compute_parameterization(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_xbound_velocity", "tb_scan_velacc_ragged"]
    assert inst.fused and tuple(res.sd.shape) == (B, G)
    args = _last(lib, "tb_scan_velacc_ragged")
    assert args["glen"] is None and args["flags"] == 0            # common grid length; exact mode: no flags
    mark = len(lib.calls)
    ta.BatchTOPPRA([vel, acc], path, gridpoints=np.linspace(0, 1, G), exact=False).compute_parameterization(0.0, 0.0)
    assert _last(lib, "tb_scan_velacc_ragged")["flags"] == engine.SCAN_FLAGS["fast_lower"]
    mark = len(lib.calls)
    inst.compute_controllable_sets(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_scan_velacc_ragged"]     # the velocity bound is kept from the first solve
    assert _last(lib, "tb_scan_velacc_ragged")["flags"] == engine.SCAN_FLAGS["backward_only"]


def test_records_path_cfg3_and_chunking(lib):  # noqa: F811
    path, vel, acc = _problem()
    taulim = np.tile(np.array([[-30.0, 30.0]]), (DOF, 1))
    torque = ta.constraint.SecondOrderConstraint.joint_torque_constraint(
        None, taulim, np.zeros(DOF), device_model=("coupled_cosine", [2.0, 0.3, 0.1, 4.9]))
    mark = len(lib.calls)
    inst = ta.BatchTOPPRA([vel, acc, torque], path, gridpoints=np.linspace(0, 1, G))
    inst.compute_parameterization(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_coeff_velacc", "tb_coeff_second_order", "tb_scan_ragged"]
    R = 4 * DOF + 4 * DOF                                          # interpolated acceleration + torque rows
    assert inst.R == R and _last(lib, "tb_scan_ragged")["R"] == R
    assert _last(lib, "tb_coeff_second_order")["row0"] == 4 * DOF  # torque rows start behind the acceleration rows
    # the same problem with a record buffer that holds 3 paths: chunks of 3, 3, 2 through ONE buffer
    per_path = 8 * engine.record_doubles(R) * G
    mark = len(lib.calls)
    small = ta.BatchTOPPRA([vel, acc, torque], path, gridpoints=np.linspace(0, 1, G), max_record_bytes=3 * per_path)
    assert small.chunk_size() == 3
    small.compute_parameterization(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_coeff_velacc", "tb_coeff_second_order", "tb_scan_ragged"] * 3
    scans, k1s = _all(lib, "tb_scan_ragged", mark), _all(lib, "tb_coeff_velacc", mark)
    assert [a["B"] for a in scans] == [3, 3, 2] and [a["B"] for a in k1s] == [3, 3, 2]
    assert len({a["records"].value for a in scans}) == 1           # one record buffer, reused
    ppoly0 = small.path.d_ppoly.data_ptr()
    stride = 8 * 4 * (N - 1) * DOF
    assert [a["ppoly"].value - ppoly0 for a in k1s] == [0, 3 * stride, 6 * stride]   # each chunk reads ITS paths' coefficients


def test_robust_and_f_rows_plans(lib):  # noqa: F811
    path, vel, acc = _problem()
    grid = np.linspace(0, 1, G)
    robust = ta.constraint.RobustLinearConstraint(acc, [0.1, 0.1, 0.1], 1)
    mark = len(lib.calls)
    ta.BatchTOPPRA([vel, robust], path, gridpoints=grid).compute_parameterization(0.0, 0.0)
    assert _launches(lib, mark)[-1] == "tb_scan_robust" and "tb_scan_ragged" not in _launches(lib, mark)
    args = _last(lib, "tb_scan_robust")
    assert (args["conic_row0"], args["conic_rows"]) == (0, 4 * DOF)   # the conic rows are the acceleration rows
    mark = len(lib.calls)
    inst = ta.BatchTOPPRA([vel, acc], path, gridpoints=grid)
    inst.compute_reachable_sets(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_coeff_velacc", "tb_reachable_sets"]
    mark = len(lib.calls)
    inst.compute_feasible_sets()
    assert _launches(lib, mark)[-1] == "tb_feasible_sets_ex"
    mark = len(lib.calls)
    sd = ta.BatchTOPPRAsd([vel, acc], path, gridpoints=grid)
    sd.set_desired_duration(np.full(B, 5.0))
    sd.compute_parameterization(0.0, 0.0)
    plan = _launches(lib, mark)
    assert plan == ["tb_xbound_velocity", "tb_scan_velacc_ragged", "tb_scan_velacc_ragged", "tb_sd_bisect"]
    flags = [a["flags"] for a in _all(lib, "tb_scan_velacc_ragged", mark)]
    assert flags == [engine.SCAN_FLAGS["sd_fast"], engine.SCAN_FLAGS["sd_slow"]]   # fastest and slowest profile, each a full scan


def _poke(pointer, array):
    ctypes.memmove(pointer.value, array.ctypes.data, array.nbytes)


def test_auto_gridpoints_plan_is_ragged(lib, monkeypatch):  # noqa: F811
    """gridpoints=None: one tb_propose_gridpoints launch, then ONE ragged fused scan over the padded grids."""
    lens = np.array([20, 35, 35, 28, 50, 31, 22, 47], dtype=np.int32)

    def propose(*args):             # what the kernel would leave behind: per-path lengths, status ok, padded grids
        max_points = args[10]
        grid = np.ones((B, max_points))
        for b, n in enumerate(lens):
            grid[b, :n] = np.linspace(0, 1, n)
        _poke(args[11], grid)
        _poke(args[13], lens)
        _poke(args[14], np.zeros(B, dtype=np.int32))
        return 0

    monkeypatch.setitem(lib.RESULTS, "tb_propose_gridpoints", propose)
    path, vel, acc = _problem()
    mark = len(lib.calls)
    inst = ta.BatchTOPPRA([vel, acc], path)
    res = inst.compute_parameterization(0.0, 0.0)
    assert _launches(lib, mark) == ["tb_propose_gridpoints", "tb_xbound_velocity", "tb_scan_velacc_ragged"]
    args = _last(lib, "tb_scan_velacc_ragged")
    assert args["G"] == int(lens.max()) and args["glen"] is not None and args["grid_shared"] == 0   # padded per-path grids
    assert tuple(res.sd.shape) == (B, int(lens.max())) and np.array_equal(inst.glen.numpy(), lens)


def test_single_path_api_plan(lib):  # noqa: F811
    """The reference-shaped single-path classes: one fit at construction; TOPPRA builds stage records once (K1) and scans."""
    path = ta.SplineInterpolator(np.linspace(0, 1, N), np.random.RandomState(1).randn(N, DOF))
    lim = np.tile(np.array([[-1.0, 1.0]]), (DOF, 1))
    mark = len(lib.calls)
    inst = ta.algorithm.TOPPRA([ta.constraint.JointVelocityConstraint(lim), ta.constraint.JointAccelerationConstraint(lim)],
                               path, gridpoints=np.linspace(0, 1, G), solver_wrapper="seidel")
    assert _launches(lib, mark) == ["tb_coeff_velacc"]
    mark = len(lib.calls)
    try:
        inst.compute_parameterization(0, 0)      # status / sd are uninitialised memory here: the outcome is not the point
    except Exception:
        pass
    assert _launches(lib, mark)[0] == "tb_scan_ragged" and _last(lib, "tb_scan_ragged")["B"] == 1
