"""Solver-wrapper strategy interface — same surface as the reference
`toppra/solverwrapper/solverwrapper.py:9-166` + `cy_seidel_solverwrapper.pyx:392-697` (seidelWrapper).

`B200SolverWrapper` holds the stage records on the device (built once at construction, like
seidelWrapper.__init__) and exposes
  * the reference per-stage method `solve_stagewise_optim(i, H, g, x_min, x_max, x_next_min, x_next_max)`
    (one tiny launch per call — for API parity and stage-level tests, not for speed);
  * whole-pass methods used by the algorithms: `parameterize`, `controllable_sets`, `feasible_sets`
    (one launch each, csrc/tb_scan.cu)."""
import numpy as np

from .. import engine
from ..constraint import RecordContext


def available_solvers(output_msg=True):
    """Solver availability in the reference's format; only the Seidel path exists here."""
    solver_availability = (("seidel", True), ("b200", True), ("hotqpoases", False), ("qpoases", False),
                           ("ecos", False), ("cvxpy", False))
    if output_msg:
        print(solver_availability)
    return solver_availability


def check_solver_availability(solver):
    return any(sname == solver and avail for sname, avail in available_solvers(False))


class SolverWrapper(object):
    """Base class of solver wrappers (reference solverwrapper.py:49-166)."""

    def __init__(self, constraint_list, path, path_discretization):
        self.constraints = constraint_list
        self.path = path
        self.path_discretization = np.array(path_discretization)
        self.N = len(path_discretization) - 1
        self.deltas = self.path_discretization[1:] - self.path_discretization[:-1]
        assert path.path_interval[0] == path_discretization[0]
        assert path.path_interval[1] == path_discretization[-1]
        for i in range(self.N):
            assert path_discretization[i + 1] > path_discretization[i]
        self.nV = 2 + sum([c.get_no_extra_vars() for c in constraint_list])

    def get_no_stages(self):
        return self.N

    def get_no_vars(self):
        return self.nV

    def get_deltas(self):
        return self.deltas

    def solve_stagewise_optim(self, i, H, g, x_min, x_max, x_next_min, x_next_max):
        raise NotImplementedError

    def setup_solver(self):
        pass

    def close_solver(self):
        pass


class B200SolverWrapper(SolverWrapper):
    """GPU implementation of the reference `seidelWrapper` (Seidel's LP, solve_lp1d=True)."""

    def __init__(self, constraint_list, path, path_discretization, solve_lp1d=True):
        from ..batch import build_records
        super(B200SolverWrapper, self).__init__(constraint_list, path, path_discretization)
        bpath = path.as_batch()
        grid = np.ascontiguousarray(self.path_discretization, dtype=np.float64)
        self.ctx = RecordContext(bpath, engine.as_device(grid, bpath.device), grid, path)
        self.records, self.R = build_records(self.ctx, constraint_list)
        self.nC = self.R + 2
        self._solve_lp1d = solve_lp1d
        self._params = None

    @property
    def params(self):
        """Per-constraint 7-tuples (a, b, c, F, g, ubound, xbound), like seidelWrapper.params."""
        if self._params is None:
            self._params = [c.compute_constraint_params(self.path, self.path_discretization)
                            for c in self.constraints]
        return self._params

    def rows(self):
        """LP row arrays like seidelWrapper's a_arr, b_arr, c_arr [(N+1), nC] (rows 0,1 zero) and low/high."""
        rec = self.records[0].cpu().numpy()
        R, G = self.R, rec.shape[0]
        out = {}
        for idx, key in enumerate(("a", "b", "c")):
            arr = np.zeros((G, self.nC))
            arr[:, 2:] = rec[:, idx * R:(idx + 1) * R]
            out[key] = arr
        out["low"] = np.stack((np.full(G, -1e8), rec[:, 3 * R]), axis=1)
        out["high"] = np.stack((np.full(G, 1e8), rec[:, 3 * R + 1]), axis=1)
        return out

    # ---- whole-pass entry points ---------------------------------------------------------------------
    def _scalar(self, v):
        return engine.as_device(np.array([float(v)]), self.ctx.device)

    def parameterize(self, sd_start, sd_end, counters=False):
        out = engine.scan(self.records, self.R, self.ctx.d_grid, self._scalar(sd_start), self._scalar(sd_end),
                          counters=counters)
        res = dict(K=out["K"][0].cpu().numpy(), sd=out["sd"][0].cpu().numpy(), u=out["u"][0].cpu().numpy(),
                   status=int(out["status"][0].item()), fail_stage=int(out["fail_stage"][0].item()))
        if counters:
            res["counters"] = out["counters"][0].cpu().numpy()
        return res

    def controllable_sets(self, sdmin, sdmax):
        out = engine.scan(self.records, self.R, self.ctx.d_grid, None, self._scalar(sdmin), self._scalar(sdmax),
                          backward_only=True)
        return out["K"][0].cpu().numpy(), int(out["status"][0].item())

    def feasible_sets(self):
        return engine.feasible_sets(self.records, self.R, self.ctx.d_grid)[0].cpu().numpy()
