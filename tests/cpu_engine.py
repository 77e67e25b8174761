"""TEST DOUBLE (tests/ only, never imported by the product): a CPU stand-in of `toppra_b200.engine` built on the oracle.

`install(monkeypatch)` replaces the tensor-level kernel wrappers of toppra_b200/engine.py by functions that do the same
job on CPU torch tensors through the plain-C restatement (oracle/), and lets `_lib.require_cuda()` pass.  With it the
whole Python host side — SplineInterpolator, constraints, record assembly, solver wrapper, TOPPRA, BatchTOPPRA chunking,
parametrizers, error paths — runs under `-m "not gpu"` (tests/test_host_pipeline_cpu.py replays the GPU parity tests
through it).  It checks HOST LOGIC only: the numbers are the oracle's, the CUDA kernels are checked by the `-m gpu`
tests.  The product has no CPU path; this module lives outside the package on purpose."""
import ctypes

import numpy as np
import torch

from oracle import oracle as orc

VAR_MIN, VAR_MAX = -1e8, 1e8


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _per_path(arr, b):
    """arr is shared ([n]) or per path ([B, n])."""
    return arr if arr.ndim == 1 else arr[b]


# ---- K0 ---------------------------------------------------------------------------------------------------------------
def spline_fit(ss, wp, bc=((0, None), (0, None))):
    ssn, wpn = _np(ss), _np(wp)
    B, n, dof = wpn.shape
    (k0, v0), (k1, v1) = bc
    out = np.empty((B, 4, n - 1, dof))
    for b in range(B):
        vb0 = np.zeros(dof) if v0 is None else _np(v0)[b]
        vb1 = np.zeros(dof) if v1 is None else _np(v1)[b]
        out[b] = orc.cubic_spline_fit(_per_path(ssn, b), wpn[b], ((k0, vb0), (k1, vb1)))
    return torch.from_numpy(out)


def ppoly_eval(ppoly, breaks, s, order):
    pp, br, sn = _np(ppoly), _np(breaks), _np(s)
    B = pp.shape[0]
    return torch.from_numpy(np.stack([orc.ppoly_eval(pp[b], _per_path(br, b), _per_path(sn, b), order) for b in range(B)]))


# ---- records ------------------------------------------------------------------------------------------------------------
def record_doubles(R):
    return (3 * int(R) + 2 + 1) & ~1


def alloc_records(B, G, R, device):
    W = record_doubles(R)
    return torch.full((B, G, W), float("nan"), dtype=torch.float64), W


def init_bounds(records, R):
    records[:, :, 3 * R] = VAR_MIN
    records[:, :, 3 * R + 1] = VAR_MAX
    records[:, :, 3 * R + 2:] = 0.0


def _write_xbound(rec, R_total, xb, mode):
    lo, hi = rec[:, :, 3 * R_total], rec[:, :, 3 * R_total + 1]
    xb = torch.from_numpy(np.ascontiguousarray(xb))
    if mode == 1:
        lo.copy_(torch.clamp(xb[..., 0], min=VAR_MIN))
        hi.copy_(torch.clamp(xb[..., 1], max=VAR_MAX))
    elif mode == 2:
        lo.copy_(xb[..., 0])
        hi.copy_(xb[..., 1])
    elif mode == 3:
        lo.copy_(torch.maximum(lo, xb[..., 0]))
        hi.copy_(torch.minimum(hi, xb[..., 1]))


def coeff_velacc(ppoly, breaks, grid, vlim, alim, interp, records, R_total, row0=0, write_xbound=1):
    pp, br, gr = _np(ppoly), _np(breaks), _np(grid)
    vl, al = _np(vlim), _np(alim)
    B, _, nseg, dof = pp.shape
    G = gr.shape[-1]
    rec = records
    xb = np.empty((B, G, 2))
    L = orc.lib()
    for b in range(B):
        g = np.ascontiguousarray(_per_path(gr, b))
        qs = orc.ppoly_eval(pp[b], _per_path(br, b), g, 1)
        qss = orc.ppoly_eval(pp[b], _per_path(br, b), g, 2)
        if vl is None:
            xb[b, :, 0], xb[b, :, 1] = VAR_MIN, VAR_MAX
        else:
            xb[b] = orc.velocity_xbound(qs, vl if vl.ndim == 2 else vl[b])
        if al is not None:
            R = (4 if interp else 2) * dof
            a, bb, c = (np.zeros((G, R)) for _ in range(3))
            alb = np.ascontiguousarray(al if al.ndim == 2 else al[b])
            dp = ctypes.POINTER(ctypes.c_double)
            L.orc_accel_rows(np.ascontiguousarray(qs).ctypes.data_as(dp), np.ascontiguousarray(qss).ctypes.data_as(dp),
                             alb.ctypes.data_as(dp), g.ctypes.data_as(dp), G, dof, 1 if interp else 0, R, 0,
                             a.ctypes.data_as(dp), bb.ctypes.data_as(dp), c.ctypes.data_as(dp))
            rec[b, :, row0:row0 + R] = torch.from_numpy(a)
            rec[b, :, R_total + row0:R_total + row0 + R] = torch.from_numpy(bb)
            rec[b, :, 2 * R_total + row0:2 * R_total + row0 + R] = torch.from_numpy(c)
    if write_xbound:
        _write_xbound(rec, R_total, xb, write_xbound)
        rec[:, :, 3 * R_total + 2:] = 0.0


def rows_canlinear(a, b, c, F, g, F_mode, grid, interp, records, R_total, row0):
    """Port of rows_canlinear_kernel (csrc/tb_coeff.cu): same term order (sum over q = 0..m-1, left to right)."""
    an, bn, cn, Fn, gn, gr = _np(a), _np(b), _np(c), _np(F), _np(g), _np(grid)
    B, G, m = an.shape
    k = Fn.shape[0] if F_mode == 0 else (Fn.shape[2] if F_mode == 1 else 2 * m)
    nrows = 2 * k if interp else k
    N = G - 1
    out = np.zeros((B, G, 3, nrows))
    for p in range(B):
        gp = _per_path(gr, p)
        for gi in range(G):
            for r in range(nrows):
                second = r >= k
                j = r - k if second else r
                src = gi + 1 if (second and gi < N) else gi
                lift = second and gi < N
                two_delta = 2 * (gp[gi + 1] - gp[gi]) if lift else 0.0
                av = an[p, src] + two_delta * bn[p, src] if lift else an[p, src]
                if F_mode >= 2:
                    col, sgn = (j, 1.0) if j < m else (j - m, -1.0)
                    ta, tb, tc = sgn * av[col], sgn * bn[p, src, col], sgn * cn[p, src, col]
                    gv = gn[p, j] if F_mode == 3 else gn[j]
                else:
                    Fr = Fn[j] if F_mode == 0 else Fn[p, src, j]
                    ta = tb = tc = 0.0
                    for q in range(m):
                        ta = ta + Fr[q] * av[q]
                        tb = tb + Fr[q] * bn[p, src, q]
                        tc = tc + Fr[q] * cn[p, src, q]
                    gv = gn[j] if F_mode == 0 else gn[p, src, j]
                out[p, gi, :, r] = (ta, tb, tc - gv)
    t = torch.from_numpy(out)
    records[:, :, row0:row0 + nrows] = t[:, :, 0]
    records[:, :, R_total + row0:R_total + row0 + nrows] = t[:, :, 1]
    records[:, :, 2 * R_total + row0:2 * R_total + row0 + nrows] = t[:, :, 2]
    return nrows


def xbound_varying(ppoly, breaks, grid, vlim_grid, records, R_total, write_xbound):
    """_create_velocity_constraint_varying (_CythonUtils.pyx:61-101): the constant-limit formula with vlim_grid[i]."""
    pp, br, gr, vg = _np(ppoly), _np(breaks), _np(grid), _np(vlim_grid)
    B, G = pp.shape[0], gr.shape[-1]
    xb = np.empty((B, G, 2))
    for b in range(B):
        qs = orc.ppoly_eval(pp[b], _per_path(br, b), _per_path(gr, b), 1)
        lim = vg if vg.ndim == 3 else vg[b]
        for i in range(G):
            xb[b, i] = orc.velocity_xbound(qs[i:i + 1], lim[i])[0]
    _write_xbound(records, R_total, xb, write_xbound)


# ---- K2 -----------------------------------------------------------------------------------------------------------------
def _rows_of(records, R, b):
    rec = _np(records[b])
    rows = np.stack((rec[:, 0:R], rec[:, R:2 * R], rec[:, 2 * R:3 * R]), axis=1)
    return rows, rec[:, 3 * R:3 * R + 2]


def _scalar(t, b, default=0.0):
    return default if t is None else float(_np(t).reshape(-1)[b])


def scan(records, R, grid, sd_start=None, sd_end=None, sd_end_hi=None, backward_only=False, counters=False,
         sd_forward=None, forward_from=None, fast_lower=False):
    B, G, W = records.shape
    if sd_forward is not None:
        return _scan_sd(records, R, grid, sd_start, sd_end, sd_forward == "slow")
    gr = _np(grid)
    K = np.zeros((B, G, 2))
    sd = np.full((B, G), np.nan)
    u = np.full((B, max(G - 1, 0)), np.nan)
    status = np.zeros(B, dtype=np.int32)
    fail = np.full(B, -1, dtype=np.int32)
    cnt = np.zeros((B, 4), dtype=np.int32)
    for b in range(B):
        rows, xb = _rows_of(records, R, b)
        w = orc.Wrapper(_per_path(gr, b), rows, xb)
        s0, s1 = _scalar(sd_start, b), _scalar(sd_end, b)
        if backward_only:
            K[b] = w.compute_controllable_sets(s1, _scalar(sd_end_hi, b, s1))
            bad = np.isnan(K[b]).any(axis=1)
            if bad.any():
                status[b], fail[b] = 3, int(np.nonzero(bad)[0].max())
            continue
        o = w.compute_parameterization(s0, s1)
        K[b], status[b] = o["K"], o["status"]
        if o["status"] == 3:
            bad = np.isnan(K[b]).any(axis=1)
            fail[b] = int(np.nonzero(bad)[0].max()) if bad.any() else 0
        else:
            sd[b], u[b] = o["sd"], o["u"]
            if o["status"] != 0:
                fail[b] = int(np.nonzero(np.isnan(o["sd"]))[0].min()) - 1
        c = w.counters() if hasattr(w, "counters") else None
        if c is not None:
            cnt[b, :3] = [c.get("lp2d", 0), c.get("lp1d", 0), c.get("resolves", 0)]
            cnt[b, 3] = o.get("retries", 0)
    out = dict(K=torch.from_numpy(K), status=torch.from_numpy(status), fail_stage=torch.from_numpy(fail),
               sd=None if backward_only else torch.from_numpy(sd), u=None if backward_only else torch.from_numpy(u))
    if counters:
        out["counters"] = torch.from_numpy(cnt)
    return out


def coeff_second_order(model, params, ppoly, breaks, grid, taulim, friction, interp, records, R_total, row0):
    """Double of tb_coeff_second_order: the registry's models in numpy, rows assembled by the rows_canlinear double."""
    pp, br, gr = _np(ppoly), _np(breaks), _np(grid)
    B, _, nseg, dof = pp.shape
    G = gr.shape[-1]
    prm = np.asarray(params, dtype=np.float64).reshape(-1)
    a, b, c = np.empty((B, G, dof)), np.empty((B, G, dof)), np.empty((B, G, dof))
    for p_ in range(B):
        g = np.ascontiguousarray(_per_path(gr, p_))
        q, qd, qdd = (orc.ppoly_eval(pp[p_], _per_path(br, p_), g, o) for o in (0, 1, 2))
        sq, cq = np.sin(q), np.cos(q)
        if model == "coupled_cosine":
            m0, m1, h, g0 = prm
            c[p_] = g0 * sq
            a[p_] = m0 * qd + m1 * (cq * (cq * qd).sum(-1, keepdims=True) + sq * (sq * qd).sum(-1, keepdims=True))
            b[p_] = (m0 * qdd + m1 * (cq * (cq * qdd).sum(-1, keepdims=True) + sq * (sq * qdd).sum(-1, keepdims=True))
                     + h * sq * (qd * qd).sum(-1, keepdims=True))
        else:
            c[p_] = prm[1::2] * sq
            a[p_] = prm[0::2] * qd
            b[p_] = prm[0::2] * qdd
        if friction is not None:
            c[p_] = c[p_] + np.sign(qd) * _np(friction)
    tl = _np(taulim)
    gaug = np.concatenate((tl[..., 1], -tl[..., 0]), axis=-1)
    return rows_canlinear(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(c), None, torch.from_numpy(gaug),
                          2 if gaug.ndim == 1 else 3, grid, interp, records, R_total, row0)


def xbound_constant(ppoly, breaks, grid, vlim, records, R_total, write_xbound):
    coeff_velacc(ppoly, breaks, grid, vlim, None, False, records, R_total, 0, write_xbound)


def scan_velacc(ppoly, breaks, grid, alim, interp, xbound, sd_start=None, sd_end=None, sd_end_hi=None,
                backward_only=False, counters=False, sd_forward=None, forward_from=None, fast_lower=False):
    """Double of the fused vel+acc scan: materialise the acceleration rows with the K1 double, take the velocity bound
    from `xbound`, run the record scan."""
    B, _, nseg, dof = ppoly.shape
    G = grid.shape[-1]
    R = (4 if interp else 2) * dof
    rec, _ = alloc_records(B, G, R, None)
    coeff_velacc(ppoly, breaks, grid, None, alim, interp, rec, R, 0, 0)
    rec[:, :, 3 * R:3 * R + 2] = xbound
    return scan(rec, R, grid, sd_start, sd_end, sd_end_hi, backward_only, counters, sd_forward, forward_from, fast_lower)


def _scan_sd(records, R, grid, sd_start, sd_end, slow):
    """TOPPRAsd passes (desired_duration_algorithm.py:42-121, 207-234): controllable sets, then a forward pass with no
    retry rule and x_next = clip(x + 2 delta u - 1e-5, K[i+1]); fastest: g = (-2 delta, -1), slowest: g = (2 delta, 1).
    The `sd` output holds x = sd^2 like the kernel's TB_SCAN_SD_FORWARD mode."""
    B, G, W = records.shape
    gr = _np(grid)
    K = np.zeros((B, G, 2))
    xs, us = np.full((B, G), np.nan), np.full((B, G - 1), np.nan)
    status, fail = np.zeros(B, dtype=np.int32), np.full(B, -1, dtype=np.int32)
    for b in range(B):
        rows, xb = _rows_of(records, R, b)
        g = _per_path(gr, b)
        w = orc.Wrapper(g, rows, xb)
        s0, s1 = _scalar(sd_start, b), _scalar(sd_end, b)
        K[b] = w.compute_controllable_sets(s1, s1)
        x0 = s0 * s0
        if np.isnan(K[b]).any() or x0 + 1e-5 < K[b, 0, 0] or K[b, 0, 1] + 1e-5 < x0:
            status[b] = 3
            fail[b] = int(np.nonzero(np.isnan(K[b]).any(axis=1))[0].max()) if np.isnan(K[b]).any() else 0
            continue
        xs[b, 0] = x0
        for i in range(G - 1):
            delta = g[i + 1] - g[i]
            obj = [2 * delta, 1.0] if slow else [-2 * delta, -1.0]
            u = w.solve_stagewise_optim(i, None, obj, xs[b, i], xs[b, i], K[b, i + 1, 0], K[b, i + 1, 1])[0]
            if np.isnan(u):
                status[b], fail[b] = 1, i
                break
            us[b, i] = u
            xs[b, i + 1] = min(K[b, i + 1, 1], max(K[b, i + 1, 0], xs[b, i] + 2 * delta * u - 1e-5))
    return dict(K=torch.from_numpy(K), sd=torch.from_numpy(xs), u=torch.from_numpy(us), status=torch.from_numpy(status),
                fail_stage=torch.from_numpy(fail))


def scan_robust(records, R, conic_row0, conic_rows, ellipsoid, grid, sd_start=None, sd_end=None, backward_only=False,
                counters=False, feasible_sets=False):
    if backward_only or feasible_sets:
        raise NotImplementedError("cpu_engine test double: robust controllable / feasible sets are gpu-only here")
    B, G, W = records.shape
    gr = _np(grid)
    K, sd, u = np.zeros((B, G, 2)), np.full((B, G), np.nan), np.full((B, max(G - 1, 1)), np.nan)
    status = np.zeros(B, dtype=np.int32)
    for b in range(B):
        rows, xb = _rows_of(records, R, b)
        o = orc.solve_rows_robust(rows, xb, _per_path(gr, b), conic_row0, conic_rows, ellipsoid, _scalar(sd_start, b),
                                  _scalar(sd_end, b))
        K[b], status[b] = o["K"], o["status"]
        if o["status"] == 0:
            sd[b], u[b, :G - 1] = o["sd"], o["u"]
    out = dict(K=torch.from_numpy(K), sd=torch.from_numpy(sd), u=torch.from_numpy(u[:, :max(G - 1, 0)]),
               status=torch.from_numpy(status), fail_stage=torch.full((B,), -1, dtype=torch.int32))
    if counters:
        out["counters"] = torch.zeros((B, 4), dtype=torch.int32)
    return out


def feasible_sets(records, R, grid):
    B, G, W = records.shape
    gr = _np(grid)
    X = np.empty((B, G, 2))
    for b in range(B):
        rows, xb = _rows_of(records, R, b)
        X[b] = orc.Wrapper(_per_path(gr, b), rows, xb).compute_feasible_sets()
    return torch.from_numpy(X)


# ---- LP shims -------------------------------------------------------------------------------------------------------------
def lp2d_batch(v, a, b, c, low, high, active_c=None):
    v, low, high = (np.asarray(_np(t) if isinstance(t, torch.Tensor) else t, dtype=np.float64) for t in (v, low, high))
    B = v.shape[0]
    a, b, c = (np.asarray(_np(t) if isinstance(t, torch.Tensor) else t, dtype=np.float64).reshape(B, -1) for t in (a, b, c))
    act = np.zeros((B, 2), dtype=np.int64) if active_c is None else np.asarray(active_c, dtype=np.int64).reshape(B, 2)
    res, val, var, out_act = np.zeros(B, np.int32), np.zeros(B), np.zeros((B, 2)), np.zeros((B, 2), np.int32)
    for i in range(B):
        res[i], val[i], var[i], out_act[i] = orc.lp2d(v[i], a[i], b[i], c[i], low[i], high[i], act[i])
    return res, val, var, out_act


def lp1d_batch(v, a, b, low, high):
    v, low, high = (np.asarray(_np(t) if isinstance(t, torch.Tensor) else t, dtype=np.float64) for t in (v, low, high))
    B = v.shape[0]
    a, b = (np.asarray(_np(t) if isinstance(t, torch.Tensor) else t, dtype=np.float64).reshape(B, -1) for t in (a, b))
    res, val, var, act = np.zeros(B, np.int32), np.zeros(B), np.zeros(B), np.zeros(B, np.int32)
    for i in range(B):
        res[i], val[i], var[i], act[i] = orc.lp1d(v[i], a[i], b[i], float(low[i]), float(high[i]))
    return res, val, var, act


# ---- K3: ParametrizeConstAccel (parametrizer.py:52-129) -------------------------------------------------------------------
def time_grid(sd, grid):
    sdn, gr = _np(sd), _np(grid)
    B, G = sdn.shape
    t, us = np.zeros((B, G)), np.zeros((B, G - 1))
    for b in range(B):
        s, v = _per_path(gr, b), sdn[b]
        x = v ** 2
        for i in range(G - 1):
            us[b, i] = 0.5 * (x[i + 1] - x[i]) / (s[i + 1] - s[i])
            t[b, i + 1] = t[b, i] + 2 * (s[i + 1] - s[i]) / (v[i] + v[i + 1])
    return torch.from_numpy(t), torch.from_numpy(us)


def constaccel_eval(ppoly, breaks, grid, sd, t_grid, us, ts, order):
    pp, br, gr, sdn, tg, un, tn = (_np(t) for t in (ppoly, breaks, grid, sd, t_grid, us, ts))
    B, dof = pp.shape[0], pp.shape[3]
    M = tn.shape[-1]
    out = np.zeros((B, M, dof))
    for b in range(B):
        s, t = _per_path(gr, b), _per_path(tn, b)
        idx = np.searchsorted(tg[b], t, side="right") - 1
        idx = np.where(idx == un.shape[1], idx - 1, idx)
        dt = t - tg[b][idx]
        u = un[b][idx]
        v = sdn[b][idx] + dt * u
        pos = s[idx] + dt * sdn[b][idx] + 0.5 * dt ** 2 * u
        ev = lambda k: orc.ppoly_eval(pp[b], _per_path(br, b), pos, k)  # noqa: E731
        if order == 0:
            out[b] = ev(0)
        elif order == 1:
            out[b] = ev(1) * v[:, None]
        else:
            out[b] = ev(2) * v[:, None] ** 2 + ev(1) * u[:, None]
    return torch.from_numpy(out)


class _NoStream(object):
    def synchronize(self):
        pass


PATCHED = ("spline_fit", "ppoly_eval", "record_doubles", "alloc_records", "init_bounds", "coeff_velacc",
           "rows_canlinear", "coeff_second_order", "xbound_varying", "xbound_constant", "scan", "scan_velacc", "scan_robust", "feasible_sets", "lp2d_batch", "lp1d_batch",
           "time_grid", "constaccel_eval")


def install(monkeypatch):
    """Route toppra_b200 through this module for the duration of one test."""
    import toppra_b200  # noqa: F401
    from toppra_b200 import _lib, engine
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch)
    monkeypatch.setattr(engine, "default_device", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    for name in PATCHED:
        assert hasattr(engine, name), name
        monkeypatch.setattr(engine, name, globals()[name])
    return toppra_b200


def install_plain():
    """The same routing without pytest's monkeypatch: for spawned worker processes (tests/test_distributed_gloo.py),
    which end with the test anyway."""
    class _Setter(object):
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)
    return install(_Setter)
