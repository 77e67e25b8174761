"""CPU: the C-ABI library loads and exports every symbol include/toppra_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "toppra_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ("tb_spline_fit", "tb_ppoly_eval", "tb_coeff_velacc", "tb_rows_canlinear", "tb_scan", "tb_scan_ex",
                 "tb_feasible_sets", "tb_solve_velacc_host", "tb_lp2d_batch", "tb_lp1d_batch", "tb_version"):
        assert name in syms


def test_library_exports_every_declared_symbol():
    from toppra_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build() / make -C toppra_b200/csrc"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name
    assert _lib.load().tb_version() == 100
    # every prototype the Python host binds is declared in the header
    assert set(_lib.exported_symbols()) <= set(declared_symbols())


def test_argument_errors_do_not_need_a_gpu():
    from toppra_b200 import _lib
    lib = _lib.load()
    assert lib.tb_record_doubles(28) == 86
    assert lib.tb_record_doubles(3) == 12  # odd R: padded to even
    rc = lib.tb_scan(None, 86, 28, None, 1, 4, 10, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"bad argument" in lib.tb_last_error()
    mr, mk = ctypes.c_int(), ctypes.c_int()
    assert lib.tb_limits(ctypes.byref(mr), ctypes.byref(mk)) == 0 and mr.value >= 62 and mk.value >= 5


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import toppra_b200 as ta
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ta.SplineInterpolator([0, 0.5, 1], [[0.0, 1], [1, 2], [2, 0]])


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under toppra_b200/ may reference it."""
    pkg = os.path.join(ROOT, "toppra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "_never_", (dirpath, f)
                # nor the oracle-backed engine double the CPU tests install by monkeypatching (tests/cpu_engine.py)
                assert "cpu_engine" not in text, (dirpath, f)


def _build_c_example(tmp):
    import subprocess
    exe = os.path.join(tmp, "solve_host")
    lib_dir = os.path.join(ROOT, "toppra_b200")
    subprocess.check_call(["gcc", os.path.join(ROOT, "examples", "solve_host.c"), "-I", os.path.join(ROOT, "include"),
                           "-L", lib_dir, "-ltoppra_b200", "-Wl,-rpath," + lib_dir, "-lm", "-o", exe])
    return exe


def test_plain_c_program_links_against_the_cabi(tmp_path):
    """A torch-free C program compiles and links against include/toppra_b200.h + libtoppra_b200.so."""
    assert os.path.exists(_build_c_example(str(tmp_path)))


@pytest.mark.gpu
def test_plain_c_program_runs(tmp_path):
    import subprocess
    out = subprocess.run([_build_c_example(str(tmp_path))], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("status 0") == 8
