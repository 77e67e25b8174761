"""CPU: the argument validation of the C-ABI itself (the real libtoppra_b200.so, no GPU needed).

Every entry point checks its arguments on the host before it launches anything and reports through the return code +
tb_last_error() (include/toppra_b200.h: TB_ERR_ARG / TB_ERR_UNSUPPORTED / TB_ERR_ALIGN); errors never cross the boundary
as exceptions or crashes.  All calls below are REJECTED by that validation, so no kernel is launched and the dummy
"device" pointers (host buffers) are never dereferenced."""
import ctypes

import numpy as np
import pytest

from toppra_b200 import _lib

TB_ERR_ARG, TB_ERR_UNSUPPORTED, TB_ERR_ALIGN = -1, -2, -3


def _gpu_present():
    import torch
    return torch.cuda.is_available()


# host addresses stand in for device arrays: a case that slipped through the validation would launch a kernel on them
pytestmark = pytest.mark.skipif(_gpu_present(), reason="validation-only calls with host addresses: run without a GPU")


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.fixture(scope="module")
def buf():
    """A 16-byte aligned dummy address standing in for device arrays."""
    mem = np.zeros(1 << 16)
    addr = mem.ctypes.data
    assert addr % 16 == 0
    return mem, ctypes.c_void_p(addr)


def _defaults(name, pointer):
    """One value per parameter of `name`: `pointer` for pointers, 1 for ints, 1.0 for doubles."""
    out = []
    for t in _lib._PROTOS[name][0]:
        if t is ctypes.c_int:
            out.append(1)
        elif t is ctypes.c_double:
            out.append(1.0)
        else:
            out.append(pointer)
    return out


POINTER_ENTRIES = sorted(n for n, (argtypes, _) in _lib._PROTOS.items()
                         if n not in ("tb_limits", "tb_last_error") and any(t is ctypes.c_void_p for t in argtypes))


@pytest.mark.parametrize("name", POINTER_ENTRIES)
def test_null_pointers_are_rejected(lib, name):
    rc = getattr(lib, name)(*_defaults(name, None))
    assert rc == TB_ERR_ARG, (name, rc, lib.tb_last_error())
    assert lib.tb_last_error(), name


@pytest.mark.parametrize("name", POINTER_ENTRIES)
def test_non_positive_batch_is_rejected(lib, buf, name):
    from test_engine_marshalling import header_prototypes, PARAM_NAMES
    header_prototypes()
    if "B" not in PARAM_NAMES[name]:
        pytest.skip("no batch size")
    args = _defaults(name, buf[1])
    args[PARAM_NAMES[name].index("B")] = 0
    assert getattr(lib, name)(*args) == TB_ERR_ARG, (name, lib.tb_last_error())


def _call(lib, name, buf, **named):
    """Call `name` with dummy-but-valid defaults, overriding parameters by their header names."""
    from test_engine_marshalling import header_prototypes, PARAM_NAMES
    header_prototypes()
    args = _defaults(name, buf[1])
    for key, value in named.items():
        args[PARAM_NAMES[name].index(key)] = value
    rc = getattr(lib, name)(*args)
    return rc, (lib.tb_last_error() or b"").decode()


def test_scan_limits_and_alignment(lib, buf):
    ok = dict(W=86, R=28, B=4, G=10, grid_shared=1, flags=0)
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, R=127, W=3 * 127 + 3))
    assert rc == TB_ERR_UNSUPPORTED and "rows" in msg
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, W=85))
    assert rc == TB_ERR_ALIGN and "even" in msg
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, W=80))
    assert rc == TB_ERR_ALIGN
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, records=ctypes.c_void_p(buf[1].value + 8)))
    assert rc == TB_ERR_ALIGN and "aligned" in msg
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, flags=64))                 # TB_SCAN_UBOUND without the slots
    assert rc == TB_ERR_ARG and "UBOUND" in msg
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, W=88, flags=64 | 32))      # u-bounds exclude the fast lower LP
    assert rc == TB_ERR_UNSUPPORTED
    rc, msg = _call(lib, "tb_scan_ragged", buf, **ok)                                # glen with a shared grid
    assert rc == TB_ERR_ARG and "ragged" in msg
    rc, msg = _call(lib, "tb_scan_ragged", buf, **dict(ok, glen=None, sd=None))       # forward pass wanted, no output
    assert rc == TB_ERR_ARG and "null output" in msg
    rc, msg = _call(lib, "tb_scan_velacc_ragged", buf, nseg=4, dof=7, B=4, G=10, grid_shared=1, breaks_shared=1,
                    lim_shared=1, interp=1, flags=0)
    assert rc == TB_ERR_ARG and "ragged" in msg
    rc, msg = _call(lib, "tb_scan_velacc_ragged", buf, nseg=4, dof=7, B=4, G=10, grid_shared=1, breaks_shared=1,
                    lim_shared=1, interp=1, flags=0, glen=None, xbound=ctypes.c_void_p(buf[1].value + 8))
    assert rc == TB_ERR_ALIGN
    rc, msg = _call(lib, "tb_feasible_sets_ex", buf, W=86, R=28, B=4, G=10, grid_shared=1, flags=64)
    assert rc == TB_ERR_ARG and "3R+4" in msg
    rc, msg = _call(lib, "tb_reachable_sets", buf, W=86, R=28, B=4, G=10, grid_shared=1, flags=64)
    assert rc == TB_ERR_ARG and "3R+4" in msg
    rc, msg = _call(lib, "tb_lp2d_batch", buf, B=2, n=1000)
    assert rc == TB_ERR_UNSUPPORTED and "rows" in msg


def test_spline_and_eval_arguments(lib, buf):
    ok = dict(ss_shared=1, B=2, n=5, dof=3, bc0_kind=0, bc1_kind=0)
    assert _call(lib, "tb_spline_fit", buf, **dict(ok, n=1))[0] == TB_ERR_ARG
    rc, msg = _call(lib, "tb_spline_fit", buf, **dict(ok, bc0_kind=4))
    assert rc == TB_ERR_ARG and "bc kind" in msg
    assert _call(lib, "tb_spline_fit", buf, **dict(ok, bc0_kind=3, bc1_kind=0))[0] == TB_ERR_ARG   # periodic: both ends
    assert lib.tb_spline_fit_workspace_doubles(0, 5, 3) == TB_ERR_ARG
    assert lib.tb_spline_fit_workspace_doubles(2, 5, 3) >= 0             # small n: the solve runs in registers
    rc, msg = _call(lib, "tb_ppoly_eval", buf, breaks_shared=1, B=2, nseg=4, dof=3, s_shared=1, G=7, order=3)
    assert rc == TB_ERR_ARG and "order" in msg
    rc, msg = _call(lib, "tb_constaccel_eval", buf, breaks_shared=1, nseg=4, dof=3, grid_shared=1, B=2, G=7, order=5)
    assert rc == TB_ERR_ARG and "order" in msg
    assert _call(lib, "tb_time_grid", buf, grid_shared=1, B=2, G=1)[0] == TB_ERR_ARG


def test_row_builder_arguments(lib, buf):
    ok = dict(breaks_shared=1, B=2, nseg=4, dof=3, grid_shared=1, G=7, lim_shared=1, interp=1, W=3 * 12 + 2, R_total=12,
              row0=0, write_xbound=1)
    rc, msg = _call(lib, "tb_coeff_velacc", buf, **dict(ok, R_total=200, W=3 * 200 + 2))
    assert rc == TB_ERR_UNSUPPORTED
    assert _call(lib, "tb_coeff_velacc", buf, **dict(ok, row0=4))[0] == TB_ERR_ARG            # rows would not fit behind row0
    canlin = dict(F_mode=0, B=2, G=7, m=3, k=4, grid_shared=1, interp=0, W=3 * 8 + 2, R_total=8, row0=0)
    rc, msg = _call(lib, "tb_rows_canlinear", buf, **dict(canlin, F=None))
    assert rc == TB_ERR_ARG and "F is null" in msg
    rc, msg = _call(lib, "tb_rows_canlinear", buf, **dict(canlin, F_mode=2, k=5))
    assert rc == TB_ERR_ARG and "2m" in msg
    assert _call(lib, "tb_rows_canlinear", buf, **dict(canlin, row0=6))[0] == TB_ERR_ARG
    so = dict(model=0, nparams=4, breaks_shared=1, B=2, nseg=4, dof=6, grid_shared=1, G=7, lim_shared=1, interp=1,
              W=3 * 24 + 2, R_total=24, row0=0)
    rc, msg = _call(lib, "tb_coeff_second_order", buf, **dict(so, model=9))
    assert rc == TB_ERR_UNSUPPORTED and "unknown device model" in msg
    rc, msg = _call(lib, "tb_coeff_second_order", buf, **dict(so, nparams=3))
    assert rc == TB_ERR_ARG and "parameters" in msg
    rc, msg = _call(lib, "tb_coeff_second_order", buf, **dict(so, dof=200, W=3 * 800 + 2, R_total=800))
    assert rc == TB_ERR_UNSUPPORTED
    assert _call(lib, "tb_init_bounds", buf, B=2, G=7, W=10, R_total=8)[0] == TB_ERR_ARG


def test_robust_scan_arguments(lib, buf):
    ell = np.array([0.1, 0.1, 0.1])
    ok = dict(W=86, R=28, conic_row0=0, conic_rows=28, ellipsoid_host3=ctypes.c_void_p(ell.ctypes.data), grid_shared=1, B=2,
              G=9, flags=0)
    rc, msg = _call(lib, "tb_scan_robust", buf, **dict(ok, conic_rows=29))
    assert rc == TB_ERR_ARG and "conic" in msg
    neg = np.array([0.1, -0.1, 0.1])
    rc, msg = _call(lib, "tb_scan_robust", buf, **dict(ok, ellipsoid_host3=ctypes.c_void_p(neg.ctypes.data)))
    assert rc == TB_ERR_ARG and "ellipsoid" in msg
    assert _call(lib, "tb_scan_robust", buf, **dict(ok, W=80))[0] == TB_ERR_ALIGN
    assert _call(lib, "tb_scan_robust", buf, **dict(ok, R=127, W=3 * 127 + 3, conic_rows=1))[0] == TB_ERR_UNSUPPORTED


def test_f_row_arguments(lib, buf):
    rc, msg = _call(lib, "tb_spline_time_stamps", buf, grid_shared=1, B=2, G=9)                   # glen + shared grid
    assert rc == TB_ERR_ARG and "ragged" in msg
    rc, msg = _call(lib, "tb_sd_bisect", buf, grid_shared=1, B=2, G=1 << 20, atol=1e-5, max_iter=10)
    assert rc == TB_ERR_UNSUPPORTED and "too large" in msg
