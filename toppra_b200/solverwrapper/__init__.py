"""Solver wrappers — the reference's strategy interface (`toppra/solverwrapper/solverwrapper.py:49-166`)
with one implementation: the GPU Seidel solver."""
from .solverwrapper import available_solvers, check_solver_availability, SolverWrapper, B200SolverWrapper
from .cy_seidel_solverwrapper import seidelWrapper   # the reference's name and constructor defaults (solve_lp1d=0)

__all__ = ["available_solvers", "check_solver_availability", "SolverWrapper", "B200SolverWrapper", "seidelWrapper"]
