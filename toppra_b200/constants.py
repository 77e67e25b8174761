"""Numerical constants of the TOPP-RA path — same names and values as the reference
`toppra/constants.py:14-47` (algorithm layer) and `cy_seidel_solverwrapper.pyx:17-29` (LP layer;
those live in csrc/tb_common.cuh)."""
SUPERTINY = 1e-10
TINY = 1e-8
SMALL = 1e-5
LARGE = 1000.0
VERYLARGE = 1e8
INFTY = 1e16

# Number of times xs[i] is lowered during the forward pass (reachability_algorithm.py:315-343).
MAX_TRIES = 10

MAXU = 10000
MAXX = 10000
MAXSD = 100

JVEL_MAXSD = 1e8
JACC_MAXU = 1e16

QPOASES_INFTY = 1e16
CVXPY_MAXX = 10000
CVXPY_MAXU = 10000
ECOS_MAXX = 10000
ECOS_INFTY = 1000

FOUND_OPENRAVE = False
