"""Reference-side binding of libtoppra_b200.so — the file a maintainer of hungpham2511/toppra would add as
`toppra/solverwrapper/b200_solverwrapper.py`.  Pure ctypes + numpy: no torch, nothing of the
toppra_b200 Python package.

The reference picks its solver wrapper by name in `ReachabilityAlgorithm.__init__`
(toppra/algorithm/reachabilitybased/reachability_algorithm.py:85-129) and then calls it 3N times per path
(`solve_stagewise_optim`).  That granularity is wrong for a GPU, so the name "b200" routes the two whole passes of
`compute_parameterization` (:240-376) — compute_controllable_sets + the forward loop — through ONE C-ABI call,
`tb_solve_velacc_host`, which fits the spline, builds the velocity bound, and runs the backward and forward scans on
the device with host buffers.  Everything else of the reference (gridpoint proposal, problem_data, return codes,
parametrizers, compute_trajectory) runs unchanged on the arrays that call returns.

`install(toppra)` patches the loaded reference package in memory (what the two-line edit of reachability_algorithm.py
quoted in INTEGRATION.md does in the source): it accepts solver_wrapper="b200" for problems made of a
JointVelocityConstraint and a JointAccelerationConstraint on a SplineInterpolator path, and refuses anything else."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("TOPPRA_B200_LIB",
                           os.path.join(_HERE, "..", "..", "toppra_b200", "libtoppra_b200.so"))
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_LIB_PATH)
        dp, ci = ctypes.c_void_p, ctypes.c_int
        lib.tb_solve_velacc_host.argtypes = [ci, dp, dp, ci, ci, ci, dp, ci, dp, dp, ci, ci, dp, dp, dp, dp, dp, dp]
        lib.tb_solve_velacc_host.restype = ci
        lib.tb_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def solve_velacc(ss, waypoints, gridpoints, vlim, alim, sd_start=0.0, sd_end=0.0, interpolation=True, device=0):
    """B paths at once: waypoints [B, n, dof]; vlim / alim [dof, 2] (shared) or [B, dof, 2].  Replaces, per path,
    TOPPRA([JointVelocityConstraint(vlim), JointAccelerationConstraint(alim)], SplineInterpolator(ss, wp), gridpoints,
    solver_wrapper='seidel').compute_parameterization(sd_start, sd_end, return_data=True).
    Returns (u [B, G-1], sd [B, G], K [B, G, 2], status [B]); status follows ParameterizationReturnCode's order."""
    lib = _load()
    wp = np.ascontiguousarray(waypoints, dtype=np.float64)
    B, n, dof = wp.shape
    ss = np.ascontiguousarray(ss, dtype=np.float64)
    grid = np.ascontiguousarray(gridpoints, dtype=np.float64)
    vl = np.ascontiguousarray(vlim, dtype=np.float64)
    al = np.ascontiguousarray(alim, dtype=np.float64)
    shared = al.ndim == 2
    if (vl.ndim == 2) != shared:
        vl = np.ascontiguousarray(np.broadcast_to(vl, (B, dof, 2)))
        al = np.ascontiguousarray(np.broadcast_to(al, (B, dof, 2)))
        shared = False
    G = len(grid)
    K = np.empty((B, G, 2))
    sd = np.empty((B, G))
    u = np.empty((B, max(G - 1, 1)))
    status = np.empty(B, dtype=np.int32)
    s0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sd_start, dtype=np.float64), (B,)))
    s1 = np.ascontiguousarray(np.broadcast_to(np.asarray(sd_end, dtype=np.float64), (B,)))
    p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    rc = lib.tb_solve_velacc_host(int(device), p(ss), p(wp), B, n, dof, p(grid), G, p(vl), p(al), 1 if shared else 0,
                                  1 if interpolation else 0, p(s0), p(s1), p(K), p(sd), p(u), p(status))
    if rc != 0:
        raise RuntimeError("tb_solve_velacc_host failed (rc=%d): %s" % (rc, lib.tb_last_error().decode()))
    return u[:, :G - 1], sd, K, status


def install(toppra):
    """Teach the loaded reference package the solver wrapper name "b200" (in memory; the reference tree is not edited)."""
    import toppra.algorithm.reachabilitybased.reachability_algorithm as ra
    from toppra.algorithm.algorithm import ParameterizationReturnCode
    from toppra.constraint import DiscretizationType, JointAccelerationConstraint, JointVelocityConstraint
    from toppra.interpolator import SplineInterpolator
    RA = ra.ReachabilityAlgorithm
    if getattr(RA, "_b200_installed", False):
        return
    orig_init, orig_cp = RA.__init__, RA.compute_parameterization
    codes = list(ParameterizationReturnCode)

    def __init__(self, constraint_list, path, gridpoints=None, solver_wrapper=None, **kwargs):
        use = isinstance(solver_wrapper, str) and solver_wrapper.lower() == "b200"
        # the reference builds its own seidel wrapper as well: it stays available for the per-stage methods
        # (compute_feasible_sets, compute_reachable_sets) the GPU call does not replace here
        orig_init(self, constraint_list, path, gridpoints=gridpoints,
                  solver_wrapper="seidel" if use else solver_wrapper, **kwargs)
        self._b200 = None
        if use:
            vel = [c for c in constraint_list if type(c) is JointVelocityConstraint]
            acc = [c for c in constraint_list if type(c) is JointAccelerationConstraint]
            if (len(vel) != 1 or len(acc) != 1 or len(constraint_list) != 2 or not isinstance(path, SplineInterpolator)
                    or getattr(path, "bc_type", "not-a-knot") != "not-a-knot"):
                raise NotImplementedError('solver_wrapper="b200": JointVelocityConstraint + JointAccelerationConstraint on a '
                                          "not-a-knot SplineInterpolator path")
            self._b200 = dict(ss=np.asarray(path.ss_waypoints, dtype=np.float64),
                              wp=np.asarray(path.waypoints[1], dtype=np.float64)[None], vlim=vel[0].vlim, alim=acc[0].alim,
                              interp=acc[0].discretization_type == DiscretizationType.Interpolation)

    def compute_parameterization(self, sd_start, sd_end, return_data=False):
        if getattr(self, "_b200", None) is None:
            return orig_cp(self, sd_start, sd_end, return_data)
        if sd_end < 0 or sd_start < 0:
            raise toppra.exceptions.BadInputVelocities(
                "Negative path velocities: path velocities must be positive: (%s, %s)" % (sd_start, sd_end))
        p = self._b200
        u, sd, K, status = solve_velacc(p["ss"], p["wp"], self.gridpoints, p["vlim"], p["alim"], sd_start, sd_end, p["interp"])
        K, code = K[0], codes[int(status[0])]
        self._problem_data.return_code = code
        if code == ParameterizationReturnCode.FailUncontrollable:  # reachability_algorithm.py:278-301
            return (None, None, None, K) if return_data else (None, None, None)
        self._problem_data.K = K
        sd_vec, sdd_vec = sd[0], u[0]
        v_vec = np.zeros((self.solver_wrapper.get_no_stages(), self.solver_wrapper.get_no_vars() - 2))
        self._problem_data.sd_vec, self._problem_data.sdd_vec = sd_vec, sdd_vec
        return (sdd_vec, sd_vec, v_vec, K) if return_data else (sdd_vec, sd_vec, v_vec)

    RA.__init__ = __init__
    RA.compute_parameterization = compute_parameterization
    RA._b200_installed = True
