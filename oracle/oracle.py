"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/liboracle.so (the plain-C restatement
of the reference TOPP-RA hot path, oracle/toppra_oracle.c).  Parity status: pinned against the
reference build (oracle/_ref) and tests/golden/*.npz, see tests/test_oracle_vs_reference.py."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lp = ctypes.POINTER(ctypes.c_long)

STATUS_NAMES = ["Ok", "ErrUnknown", "ErrShortPath", "FailUncontrollable", "ErrForwardPassFail"]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("toppra_oracle.c", "toppra_robust_oracle.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so", "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _setup(ctypes.CDLL(build()))
    return _LIB


def _setup(L):
    L.orc_wrapper_create.restype = ctypes.c_void_p
    for name in ("orc_wrapper_a", "orc_wrapper_b", "orc_wrapper_c", "orc_wrapper_low", "orc_wrapper_high"):
        getattr(L, name).restype = _dp
        getattr(L, name).argtypes = [ctypes.c_void_p]
    return L


class shortcut_model:
    """Context manager: route this module through liboracle_shortcuts.so (oracle/shortcut_model.c), whose wrapper
    checks the scalar model of the CUDA kernel's Seidel shortcuts on every 2-variable LP.  `stats()` returns
    dict(lps, mismatches, resolves_ref, resolves_model, a_used, a_declined, b_used, b_declined)."""
    _lib = None

    def __enter__(self):
        global _LIB
        if shortcut_model._lib is None:
            so = os.path.join(_HERE, "liboracle_shortcuts.so")
            subprocess.check_call(["make", "-C", _HERE, "liboracle_shortcuts.so"], stdout=subprocess.DEVNULL)
            shortcut_model._lib = _setup(ctypes.CDLL(so))
        self._saved = _LIB
        _LIB = shortcut_model._lib
        _LIB.orc_shortcut_model_enable(1)
        self.stats(reset=True)
        return self

    def __exit__(self, *exc):
        global _LIB
        shortcut_model._lib.orc_shortcut_model_enable(0)
        _LIB = self._saved

    def stats(self, reset=False):
        out = np.zeros(8, dtype=np.int64)
        shortcut_model._lib.orc_shortcut_model_stats(out.ctypes.data_as(_lp), 1 if reset else 0)
        keys = ("lps", "mismatches", "resolves_ref", "resolves_model", "a_used", "a_declined", "b_used", "b_declined")
        return dict(zip(keys, (int(v) for v in out)))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def cubic_spline_fit(x, y, bc_type="not-a-knot"):
    """scipy.interpolate.CubicSpline(x, y, bc_type).c restated; y: [n, dof] -> c [4, n-1, dof]."""
    x, xp = _d(x)
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    y, yp = _d(y)
    n, dof = y.shape
    kinds = {"not-a-knot": 0, "clamped": 1, "natural": 2, "periodic": 3}
    if isinstance(bc_type, str):
        k0 = k1 = kinds[bc_type]
        v0 = v1 = np.zeros(dof)
    else:
        (k0, v0), (k1, v1) = bc_type
        v0 = np.broadcast_to(np.asarray(v0, dtype=np.float64), (dof,)).copy()
        v1 = np.broadcast_to(np.asarray(v1, dtype=np.float64), (dof,)).copy()
    v0, v0p = _d(v0)
    v1, v1p = _d(v1)
    c = np.zeros((4, n - 1, dof))
    rc = lib().orc_cubic_spline_fit(xp, yp, n, dof, int(k0), v0p, int(k1), v1p, c.ctypes.data_as(_dp))
    if rc != 0:
        raise ValueError("orc_cubic_spline_fit failed")
    return c


def ppoly_eval(c, x, s, order):
    c, cp = _d(c)
    x, xp = _d(x)
    s, sp = _d(np.atleast_1d(s))
    nseg, dof = c.shape[1], c.shape[2]
    out = np.zeros((s.shape[0], dof))
    lib().orc_ppoly_eval(cp, xp, nseg, dof, sp, s.shape[0], int(order), out.ctypes.data_as(_dp))
    return out


def velocity_xbound(qs, vlim):
    qs, qp = _d(qs)
    vlim, vp = _d(vlim)
    G, dof = qs.shape
    out = np.zeros((G, 2))
    lib().orc_velocity_xbound(qp, vp, G, dof, out.ctypes.data_as(_dp))
    return out


def lp1d(v, a, b, low, high):
    v, vp = _d(v)
    a, ap = _d(a)
    b, bp = _d(b)
    res, act = ctypes.c_int(), ctypes.c_int()
    optval, optvar = ctypes.c_double(), ctypes.c_double()
    lib().orc_lp1d(vp, len(a), ap, bp, ctypes.c_double(low), ctypes.c_double(high), ctypes.byref(res),
                   ctypes.byref(optval), ctypes.byref(optvar), ctypes.byref(act))
    return res.value, optval.value, optvar.value, act.value


def lp2d(v, a, b, c, low, high, active_c):
    v, vp = _d(v)
    a, ap = _d(a)
    b, bp = _d(b)
    c, cp = _d(c)
    low, lop = _d(low)
    high, hip = _d(high)
    act_in = np.ascontiguousarray(active_c, dtype=np.int64)
    res = ctypes.c_int()
    optval = ctypes.c_double()
    optvar = np.zeros(2)
    act_out = np.zeros(2, dtype=np.int32)
    lib().orc_lp2d(vp, len(a), ap, bp, cp, lop, hip, act_in.ctypes.data_as(_lp), ctypes.byref(res),
                   ctypes.byref(optval), optvar.ctypes.data_as(_dp), act_out.ctypes.data_as(_ip))
    return res.value, optval.value, optvar, act_out


class Wrapper:
    """Mirror of the reference `seidelWrapper` (cy_seidel_solverwrapper.pyx:392) over explicit rows.

    rows: [G, 3, R] (a, b, c of the static rows, i.e. a_arr[:, 2:] ...), xbound/ubound: [G, 2] or None."""

    def __init__(self, grid, rows, xbound=None, ubound=None, solve_lp1d=True):
        grid, gp = _d(grid)
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        G, _, R = rows.shape
        self.G, self.R, self.nC = G, R, R + 2
        self._w = ctypes.c_void_p(lib().orc_wrapper_create(gp, G, self.nC))
        for name, k in (("orc_wrapper_a", 0), ("orc_wrapper_b", 1), ("orc_wrapper_c", 2)):
            arr = np.ctypeslib.as_array(getattr(lib(), name)(self._w), shape=(G, self.nC))
            arr[:, 2:] = rows[:, k, :]
        if xbound is not None:
            xb, xbp = _d(xbound)
            lib().orc_wrapper_add_xbound(self._w, xbp)
        if ubound is not None:
            ub, ubp = _d(ubound)
            lib().orc_wrapper_add_ubound(self._w, ubp)
        lib().orc_wrapper_set_solve_lp1d(self._w, 1 if solve_lp1d else 0)

    def __del__(self):
        if getattr(self, "_w", None):
            lib().orc_wrapper_free(self._w)
            self._w = None

    def arrays(self):
        G, nC = self.G, self.nC
        out = {}
        for name, key in (("orc_wrapper_a", "a"), ("orc_wrapper_b", "b"), ("orc_wrapper_c", "c")):
            out[key] = np.ctypeslib.as_array(getattr(lib(), name)(self._w), shape=(G, nC)).copy()
        out["low"] = np.ctypeslib.as_array(lib().orc_wrapper_low(self._w), shape=(G, 2)).copy()
        out["high"] = np.ctypeslib.as_array(lib().orc_wrapper_high(self._w), shape=(G, 2)).copy()
        return out

    def solve_stagewise_optim(self, i, H, g, x_min, x_max, x_next_min, x_next_max):
        g, gp = _d(g)
        var = np.zeros(2)
        lib().orc_wrapper_solve_stagewise_optim(self._w, int(i), gp, ctypes.c_double(x_min), ctypes.c_double(x_max),
                                                ctypes.c_double(x_next_min), ctypes.c_double(x_next_max),
                                                var.ctypes.data_as(_dp))
        return var

    def compute_controllable_sets(self, sdmin, sdmax):
        K = np.zeros((self.G, 2))
        lib().orc_compute_controllable_sets(self._w, ctypes.c_double(sdmin), ctypes.c_double(sdmax),
                                            K.ctypes.data_as(_dp))
        return K

    def compute_feasible_sets(self):
        X = np.zeros((self.G, 2))
        lib().orc_compute_feasible_sets(self._w, X.ctypes.data_as(_dp))
        return X

    def compute_parameterization(self, sd_start, sd_end):
        K = np.zeros((self.G, 2))
        sd = np.zeros(self.G)
        u = np.zeros(self.G - 1)
        nre = ctypes.c_int()
        st = lib().orc_compute_parameterization(self._w, ctypes.c_double(sd_start), ctypes.c_double(sd_end),
                                                K.ctypes.data_as(_dp), sd.ctypes.data_as(_dp),
                                                u.ctypes.data_as(_dp), ctypes.byref(nre))
        return dict(K=K, sd=sd, u=u, status=int(st), retries=nre.value)

    def counters(self):
        out = np.zeros(3, dtype=np.int64)
        lib().orc_wrapper_counters(self._w, out.ctypes.data_as(_lp))
        return dict(lp2d=int(out[0]), lp1d=int(out[1]), resolves=int(out[2]))


def solve_velacc(c, x, grid, vlim, alim, interp=True, sd_start=0.0, sd_end=0.0, want_rows=False):
    """One path: PPoly (c [4,nseg,dof], breaks x) -> vel xbound + accel rows -> K, sd, u, status."""
    c, cp = _d(c)
    x, xp = _d(x)
    grid, gp = _d(grid)
    alim, alp = _d(alim)
    nseg, dof = c.shape[1], c.shape[2]
    G = grid.shape[0]
    R = (4 if interp else 2) * dof
    vp = None
    if vlim is not None:
        vlim, vp = _d(vlim)
    K = np.zeros((G, 2))
    sd = np.zeros(G)
    u = np.zeros(G - 1)
    rows = np.zeros((G, 3, R)) if want_rows else None
    xb = np.zeros((G, 2)) if want_rows else None
    cnt = np.zeros(3, dtype=np.int64)
    st = lib().orc_solve_velacc(cp, xp, nseg, dof, gp, G, vp, alp, 1 if interp else 0, ctypes.c_double(sd_start),
                                ctypes.c_double(sd_end), K.ctypes.data_as(_dp), sd.ctypes.data_as(_dp),
                                u.ctypes.data_as(_dp), rows.ctypes.data_as(_dp) if want_rows else None,
                                xb.ctypes.data_as(_dp) if want_rows else None, cnt.ctypes.data_as(_lp))
    out = dict(K=K, sd=sd, u=u, status=int(st), counters=dict(lp2d=int(cnt[0]), lp1d=int(cnt[1]), resolves=int(cnt[2])))
    if want_rows:
        out["rows"] = rows
        out["xbound"] = xb
    return out


def solve_rows(rows, xbound, grid, sd_start=0.0, sd_end=0.0, ubound=None):
    rows, rp = _d(rows)
    grid, gp = _d(grid)
    G, _, R = rows.shape
    xp = up = None
    if xbound is not None:
        xbound, xp = _d(xbound)
    if ubound is not None:
        ubound, up = _d(ubound)
    K = np.zeros((G, 2))
    sd = np.zeros(G)
    u = np.zeros(G - 1)
    st = lib().orc_solve_rows(rp, xp, up, gp, G, R, ctypes.c_double(sd_start), ctypes.c_double(sd_end),
                              K.ctypes.data_as(_dp), sd.ctypes.data_as(_dp), u.ctypes.data_as(_dp), None)
    return dict(K=K, sd=sd, u=u, status=int(st))


def solve_velacc_batch(c, x, grid, vlim, alim, interp=True, sd_start=None, sd_end=None, nthreads=1):
    """Batch: c [B,4,nseg,dof], x [B,nseg+1], shared grid [G]; vlim/alim [dof,2] shared or [B,dof,2]."""
    c, cp = _d(c)
    x, xp = _d(x)
    grid, gp = _d(grid)
    B, _, nseg, dof = c.shape
    G = grid.shape[0]
    alim, alp = _d(alim)
    astr = dof * 2 if alim.ndim == 3 else 0
    vp, vstr = None, 0
    if vlim is not None:
        vlim, vp = _d(vlim)
        vstr = dof * 2 if vlim.ndim == 3 else 0
    s0p = s1p = None
    if sd_start is not None:
        sd_start, s0p = _d(np.broadcast_to(sd_start, (B,)))
    if sd_end is not None:
        sd_end, s1p = _d(np.broadcast_to(sd_end, (B,)))
    K = np.zeros((B, G, 2))
    sd = np.zeros((B, G))
    u = np.zeros((B, G - 1))
    status = np.zeros(B, dtype=np.int32)
    lib().orc_solve_velacc_batch(cp, xp, B, nseg, dof, gp, G, vp, ctypes.c_long(vstr), alp, ctypes.c_long(astr),
                                 1 if interp else 0, s0p, s1p, K.ctypes.data_as(_dp), sd.ctypes.data_as(_dp),
                                 u.ctypes.data_as(_dp), status.ctypes.data_as(_ip), int(nthreads))
    return dict(K=K, sd=sd, u=u, status=status)


def solve_rows_robust(rows, xbound, grid, conic_row0, conic_rows, ellipsoid, sd_start=0.0, sd_end=0.0):
    """Robust (conic) TOPP-RA over explicit rows — oracle/toppra_robust_oracle.c (parity UNPINNED: problem
    definition only, see that file's header)."""
    rows, rp = _d(rows)
    grid, gp = _d(grid)
    G, _, R = rows.shape
    xp = None
    if xbound is not None:
        xbound, xp = _d(xbound)
    ell, ep = _d(ellipsoid)
    K = np.zeros((G, 2))
    sd = np.zeros(G)
    u = np.zeros(G - 1)
    nev = ctypes.c_long()
    st = lib().orc_solve_rows_robust(rp, xp, gp, G, R, int(conic_row0), int(conic_rows), ep, ctypes.c_double(sd_start),
                                     ctypes.c_double(sd_end), K.ctypes.data_as(_dp), sd.ctypes.data_as(_dp),
                                     u.ctypes.data_as(_dp), ctypes.byref(nev))
    return dict(K=K, sd=sd, u=u, status=int(st), n_eval=nev.value)


def feasible_rows_robust(rows, xbound, grid, conic_row0, conic_rows, ellipsoid):
    """Feasible sets X [G, 2] of the robust problem (every stage on its own, x_next boxed to +-1e4) —
    oracle/toppra_robust_oracle.c orc_feasible_rows_robust; NaN marks an infeasible stage."""
    rows, rp = _d(rows)
    grid, gp = _d(grid)
    G, _, R = rows.shape
    xp = None
    if xbound is not None:
        xbound, xp = _d(xbound)
    ell, ep = _d(ellipsoid)
    X = np.zeros((G, 2))
    fn = lib().orc_feasible_rows_robust
    fn.restype = None
    fn(rp, xp, gp, G, R, int(conic_row0), int(conic_rows), ep, X.ctypes.data_as(_dp))
    return X
