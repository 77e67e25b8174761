/*
 * oracle/toppra_robust_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the robust (conic) TOPP-RA stage problems solved by K2r (toppra_b200/csrc/tb_robust.cu).
 *
 * Parity status: UNPINNED.  The reference solves these problems with ECOS (third-party interior-point solver,
 * unpinned version, NOT installed here) through ecosWrapper.solve_stagewise_optim
 * (toppra/solverwrapper/ecos_solverwrapper.py:90-207); no reference output can be generated in this
 * environment and the reference's own tests for this path only check sanity bounds
 * (tests/tests/retime/test_retime_wconic_constraints.py:31-48).  What this file restates is the PROBLEM
 * DEFINITION:
 *   rows (conic_constraint.py:95-124)  a u + b x + c + || diag(ru, rx, rc) [u, x, 1] ||_2 <= 0
 *   x box: NaN bounds -> -/+1000 (ECOS_INFTY), x <= min(1e4, xbound_hi), x >= xbound_lo  (ecos_solverwrapper.py:112-172)
 *   x_next rows: x_next_min <= x + 2 delta u <= x_next_max
 *   driver: reachability_algorithm.py:166-376.
 * It is validated by (i) zero ellipsoid == linear Seidel results, (ii) feasibility residuals of the returned
 * points, (iii) monotonicity in the ellipsoid size (tests/test_robust.py), and it pins the CUDA kernel bit-for-bit.
 */
#include <math.h>
#include <stdlib.h>

#define LP_TINY 1e-10
#define LP_SMALL 1e-8
#define VAR_MIN (-100000000.0)
#define VAR_MAX (100000000.0)
#define ECOS_INFTY 1000.0
#define ECOS_MAXX 10000.0
#define ALG_TINY 1e-8
#define ALG_SMALL 1e-5
#define MAX_TRIES 10

typedef struct {
  int nC;                 /* rows incl. the two x_next rows at index 0,1 */
  const double *a, *b, *c;
  const unsigned char *conic;
  double ru, rx, rc;
  long n_eval;
} stage_t;

static void row_u_bounds(int conic, double a, double b, double c, double ru, double rx, double rc, double x,
                         double *lo, double *hi, int *bad) {
  double beta = b * x + c;
  double gamma2 = 0.0;
  if (conic) gamma2 = rx * rx * (x * x) + rc * rc;
  if (!conic || ru == 0.0) {
    if (conic) beta = beta + sqrt(gamma2);
    if (a > LP_TINY) { double t = -beta / a; if (t < *hi) *hi = t; }
    else if (a < -LP_TINY) { double t = -beta / a; if (t > *lo) *lo = t; }
    else if (beta > LP_SMALL) *bad = 1;
    return;
  }
  double A = a * a - ru * ru;
  double D = a * a * gamma2 + ru * ru * (beta * beta - gamma2);
  double p = -a * beta;
  if (A > 0.0) {
    double sq = sqrt(D > 0.0 ? D : 0.0);
    double s = (a > 0.0) ? 1.0 : -1.0;
    double root = (s * p <= 0.0) ? (p - s * sq) / A : (beta * beta - gamma2) / (p + s * sq);
    if (a > 0.0) { if (root < *hi) *hi = root; } else { if (root > *lo) *lo = root; }
  } else if (A < 0.0) {
    if (D < 0.0 || beta > 0.0) { *bad = 1; return; }
    double sq = sqrt(D);
    double q = p + ((p >= 0.0) ? sq : -sq);
    double r1, r2;
    if (q != 0.0) { r1 = q / A; r2 = (beta * beta - gamma2) / q; } else { r1 = 0.0; r2 = 0.0; }
    double rl = (r1 < r2) ? r1 : r2, rh = (r1 < r2) ? r2 : r1;
    if (rl > *lo) *lo = rl;
    if (rh < *hi) *hi = rh;
  } else {
    if (beta >= 0.0) { *bad = 1; return; }
    double root = (gamma2 - beta * beta) / (2 * a * beta);
    if (a > 0.0) { if (root < *hi) *hi = root; } else { if (root > *lo) *lo = root; }
  }
}

static double u_interval(stage_t *st, double x, double *uhi) {
  double lo = VAR_MIN, hi = VAR_MAX;
  int bad = 0;
  for (int r = 0; r < st->nC; ++r)
    row_u_bounds(st->conic[r], st->a[r], st->b[r], st->c[r], st->ru, st->rx, st->rc, x, &lo, &hi, &bad);
  st->n_eval++;
  *uhi = hi;
  if (bad) return -INFINITY;
  return hi - lo;
}

static int extreme_x(stage_t *st, int dir, double xl, double xh, double *xout, double hint) {
  if (xl > xh) return 0;
  double uh;
  double xgoal = (dir > 0) ? xh : xl, xother = (dir > 0) ? xl : xh;
  double wg = u_interval(st, xgoal, &uh);
  if (wg >= 0.0) { *xout = xgoal; return 1; }
  double wo = u_interval(st, xother, &uh);
  double xf = xother, wf = wo;
  double xb0 = xgoal, wb0 = wg;
  if (wo >= 0.0 && hint > xl && hint < xh) {
    double h1 = hint, h2 = (dir > 0) ? fmin(xh, hint * 1.25 + 1e-9) : fmax(xl, hint * 0.8 - 1e-9);
    double w1 = u_interval(st, h1, &uh);
    if (w1 >= 0.0) {
      xf = h1; wf = w1;
      if (h2 != xgoal) {
        double w2 = u_interval(st, h2, &uh);
        if (w2 >= 0.0) { xf = h2; wf = w2; } else { xb0 = h2; wb0 = w2; }
      }
    } else {
      xb0 = h1; wb0 = w1;
    }
  }
  if (!(wo >= 0.0)) {
    const double invphi = 0.6180339887498949;
    double lo = xl, hi = xh;
    double x1 = hi - invphi * (hi - lo), x2 = lo + invphi * (hi - lo);
    double w1 = u_interval(st, x1, &uh);
    double w2 = u_interval(st, x2, &uh);
    int found = 0;
    for (int it = 0; it < 80; ++it) {
      if (w1 >= 0.0) { xf = x1; wf = w1; found = 1; break; }
      if (w2 >= 0.0) { xf = x2; wf = w2; found = 1; break; }
      if (!(hi - lo > 1e-15 * (fabs(hi) + fabs(lo)) + 1e-300)) break;
      if (w1 > w2) { hi = x2; x2 = x1; w2 = w1; x1 = hi - invphi * (hi - lo); w1 = u_interval(st, x1, &uh); }
      else { lo = x1; x1 = x2; w1 = w2; x2 = lo + invphi * (hi - lo); w2 = u_interval(st, x2, &uh); }
    }
    if (!found) return 0;
  }
  double xb = xb0, wb = wb0;
  for (int it = 0; it < 200; ++it) {
    double width = fabs(xb - xf);
    if (!(width > 2.3e-16 * (fabs(xb) + fabs(xf)) + 1e-300)) break;
    double t;
    int finite = wb > -1e300;
    if (finite && (it % 3) != 2) {
      double frac = wf / (wf - wb);
      frac = (frac < 0.02) ? 0.02 : ((frac > 0.98) ? 0.98 : frac);
      t = xf + (xb - xf) * frac;
    } else {
      t = 0.5 * (xf + xb);
    }
    if (t == xf || t == xb) break;
    double wt = u_interval(st, t, &uh);
    if (wt >= 0.0) { xf = t; wf = wt; } else { xb = t; wb = wt; }
  }
  *xout = xf;
  return 1;
}

/* rows: [G][3][R] (a,b,c), xbound [G][2]; rows [conic0, conic0+conicn) are robust.  Returns status. */
int orc_solve_rows_robust(const double *rows, const double *xbound, const double *grid, int G, int R, int conic0,
                          int conicn, const double *ell, double sd_start, double sd_end, double *K, double *sd,
                          double *u, long *n_eval_out) {
  int N = G - 1, nC = R + 2;
  double *a = (double *)malloc(sizeof(double) * nC), *b = (double *)malloc(sizeof(double) * nC);
  double *c = (double *)malloc(sizeof(double) * nC);
  unsigned char *conic = (unsigned char *)calloc(nC, 1);
  for (int r = 0; r < R; ++r) conic[2 + r] = (r >= conic0 && r < conic0 + conicn);
  stage_t st = {nC, a, b, c, conic, ell[0], ell[1], ell[2], 0};
  for (int i = 0; i < 2 * G; ++i) K[i] = 0.0;
  for (int i = 0; i < G; ++i) sd[i] = NAN;
  for (int i = 0; i < N; ++i) u[i] = NAN;
  K[2 * N] = sd_end * sd_end; K[2 * N + 1] = sd_end * sd_end;
  int status = 0;
  for (int i = N - 1; i >= 0; --i) {
    for (int r = 0; r < R; ++r) {
      a[2 + r] = rows[((size_t)i * 3 + 0) * R + r]; b[2 + r] = rows[((size_t)i * 3 + 1) * R + r];
      c[2 + r] = rows[((size_t)i * 3 + 2) * R + r];
    }
    double xlo_b = xbound ? fmax(VAR_MIN, xbound[i * 2]) : VAR_MIN, xhi_b = xbound ? fmin(VAR_MAX, xbound[i * 2 + 1]) : VAR_MAX;
    double xl = fmax(-ECOS_INFTY, xlo_b), xh = fmin(ECOS_INFTY, fmin(ECOS_MAXX, xhi_b));
    double delta = grid[i + 1] - grid[i];
    a[0] = -2 * delta; b[0] = -1.0; c[0] = K[2 * (i + 1)];
    a[1] = 2 * delta; b[1] = 1.0; c[1] = -K[2 * (i + 1) + 1];
    double x_upper = NAN, x_lower = NAN;
    int ok_hi = extreme_x(&st, +1, xl, xh, &x_upper, K[2 * (i + 1) + 1]);
    int ok_lo = ok_hi && extreme_x(&st, -1, xl, xh, &x_lower, K[2 * (i + 1)]);
    if (!ok_hi) x_upper = NAN;
    if (!ok_lo) x_lower = NAN;
    if (x_lower < 0) x_lower = 0;
    K[2 * i] = x_lower; K[2 * i + 1] = x_upper;
    if (!(ok_hi && ok_lo)) { status = 3; break; }
  }
  double x_start = sd_start * sd_start;
  if (status == 0 && (x_start + ALG_SMALL < K[0] || K[1] + ALG_SMALL < x_start)) status = 3;
  if (status == 0) {
    double *xs = (double *)calloc(G, sizeof(double));
    for (int i = 0; i < N; ++i) u[i] = 0.0;
    xs[0] = x_start;
    for (int i = 0; i < N; ++i) {
      for (int r = 0; r < R; ++r) {
        a[2 + r] = rows[((size_t)i * 3 + 0) * R + r]; b[2 + r] = rows[((size_t)i * 3 + 1) * R + r];
        c[2 + r] = rows[((size_t)i * 3 + 2) * R + r];
      }
      double delta = grid[i + 1] - grid[i];
      double k0 = K[2 * (i + 1)], k1 = K[2 * (i + 1) + 1];
      a[0] = -2 * delta; b[0] = -1.0; c[0] = k0;
      a[1] = 2 * delta; b[1] = 1.0; c[1] = -k1;
      int tries = 0, ok;
      double uh = 0.0, x = xs[i];
      while (1) {
        double w = u_interval(&st, x, &uh);
        ok = w >= 0.0;
        if (ok || tries >= MAX_TRIES) break;
        x = fmax(x - ALG_TINY, 0.999 * x);
        ++tries;
      }
      xs[i] = x;
      if (!ok) { for (int j = i + 1; j < G; ++j) xs[j] = NAN; status = 1; break; }
      u[i] = uh;
      double x_next = x + 2 * delta * uh;
      x_next = fmax(x_next - ALG_TINY, 0.9999 * x_next);
      xs[i + 1] = fmin(k1, fmax(k0, x_next));
    }
    for (int j = 0; j < G; ++j) sd[j] = sqrt(xs[j]);
    free(xs);
  }
  if (n_eval_out) *n_eval_out = st.n_eval;
  free(a); free(b); free(c); free(conic);
  return status;
}

/* Feasible sets of the robust problem: every stage on its own, x and x_next boxed to [-1e4, 1e4] (the reference's
 * compute_feasible_sets, reachability_algorithm.py:131-164, passes x_next bounds of +-CVXPY_MAXX = 1e4, constants.py:40).
 * rows: [G][3][R], xbound [G][2], X out [G][2] (NaN = stage infeasible).  No hint: both searches start from the box. */
#define CVXPY_MAXX 10000.0
void orc_feasible_rows_robust(const double *rows, const double *xbound, const double *grid, int G, int R, int conic0,
                              int conicn, const double *ell, double *X) {
  int N = G - 1, nC = R + 2;
  double *a = (double *)malloc(sizeof(double) * nC), *b = (double *)malloc(sizeof(double) * nC);
  double *c = (double *)malloc(sizeof(double) * nC);
  unsigned char *conic = (unsigned char *)calloc(nC, 1);
  for (int r = 0; r < R; ++r) conic[2 + r] = (r >= conic0 && r < conic0 + conicn);
  stage_t st = {nC, a, b, c, conic, ell[0], ell[1], ell[2], 0};
  for (int i = 0; i <= N; ++i) {
    for (int r = 0; r < R; ++r) {
      a[2 + r] = rows[((size_t)i * 3 + 0) * R + r]; b[2 + r] = rows[((size_t)i * 3 + 1) * R + r];
      c[2 + r] = rows[((size_t)i * 3 + 2) * R + r];
    }
    double xlo_b = xbound ? xbound[i * 2] : VAR_MIN, xhi_b = xbound ? xbound[i * 2 + 1] : VAR_MAX;
    double xl = fmax(-CVXPY_MAXX, xlo_b), xh = fmin(CVXPY_MAXX, fmin(ECOS_MAXX, xhi_b));
    a[0] = 0.0; b[0] = 0.0; c[0] = -1.0;
    a[1] = 0.0; b[1] = 0.0; c[1] = -1.0;
    if (i < N) {
      double delta = grid[i + 1] - grid[i];
      a[0] = -2 * delta; b[0] = -1.0; c[0] = -CVXPY_MAXX;
      a[1] = 2 * delta; b[1] = 1.0; c[1] = -CVXPY_MAXX;
    }
    double x0 = NAN, x1 = NAN;
    if (!extreme_x(&st, -1, xl, xh, &x0, NAN)) x0 = NAN;
    if (!extreme_x(&st, +1, xl, xh, &x1, NAN)) x1 = NAN;
    if (x0 < 0) x0 = 0;
    X[2 * i] = x0; X[2 * i + 1] = x1;
  }
  free(a); free(b); free(c); free(conic);
}
