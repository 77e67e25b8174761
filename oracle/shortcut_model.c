/*
 * oracle/shortcut_model.c — TEST INFRASTRUCTURE ONLY.
 *
 * Scalar model of the two Seidel shortcuts of the CUDA scan kernel (toppra_b200/csrc/tb_scan.cu, lp2d_impl:
 * "Shortcut A" and "Shortcut B") with the same decision rules and margins, plus a checker that runs the model
 * beside the sequential restatement (toppra_oracle.c: lp2d, cy_seidel_solverwrapper.pyx:149-390) on every
 * 2-variable LP the wrapper solves and counts results that differ in any bit (optimum, active pair, feasibility).
 * tests/test_shortcut_model.py drives it on CPU over benchmark-like, velocity-limited, badly scaled and
 * near-degenerate problems; the GPU tests then check the kernel itself against the oracle.
 *
 * Why the shortcuts are exact: every Seidel visit (pyx:276-384) recomputes the point from scratch on the line of
 * the visited row over ALL earlier rows, so the state after the loop depends only on the LAST visited row.  A
 * shortcut names a row the reference is certain to visit (decisions far above its 1e-10 threshold, no skipped
 * visit able to end the solve early) and lets the ordinary, exact walk continue from an exact re-solve on it.
 *
 * Built as its own library (liboracle_shortcuts.so) that contains a private copy of the restatement.
 */
#include "toppra_oracle.c"

static long sm_lp = 0, sm_resolve_ref = 0, sm_resolve_model = 0, sm_mismatch = 0;
static long sm_a_used = 0, sm_a_abort = 0, sm_b_used = 0, sm_b_abort = 0;
static double sm_gap = 1e-7;  /* tb_scan.cu: SKIP_GAP */
static double sm_tmax = 90.0; /* tb_scan.cu: SKIP_TMAX */

/* one exact re-solve on the row at position k of the order (pyx:276-384) */
static int sm_resolve(const double v[3], int k, const long *order, const double *a, const double *b, const double *c,
                      const double low[2], const double high[2], double *a_1d, double *b_1d, double cur[2],
                      int act[2]) {
  const long i = order[k];
  const double nrm = a[i] * a[i] + b[i] * b[i];
  const double z0 = -a[i] * c[i] / nrm, z1 = -b[i] * c[i] / nrm;
  const double d0 = -b[i], d1 = a[i];
  double v_1d[2] = {d0 * v[0] + d1 * v[1], 0};
  const int n1 = 4 + k;
  for (int j = 0; j < n1; ++j) {
    double aj, bj, cj;
    if (j == k) { aj = -1; bj = 0; cj = low[0]; }
    else if (j == k + 1) { aj = 1; bj = 0; cj = -high[0]; }
    else if (j == k + 2) { aj = 0; bj = -1; cj = low[1]; }
    else if (j == k + 3) { aj = 0; bj = 1; cj = -high[1]; }
    else { aj = a[order[j]]; bj = b[order[j]]; cj = c[order[j]]; }
    const double denom = d0 * aj + d1 * bj;
    if (denom > LP_TINY) { a_1d[j] = 1.0; b_1d[j] = -(-(cj + z1 * bj + z0 * aj) / denom); }
    else if (denom < -LP_TINY) { a_1d[j] = -1.0; b_1d[j] = -(cj + z1 * bj + z0 * aj) / denom; }
    else { if (cj + z1 * bj + z0 * aj > LP_SMALL) return 0; a_1d[j] = 0; b_1d[j] = -1.0; }
  }
  const LpSol s1 = lp1d(v_1d, n1, a_1d, b_1d, -LP_INF, LP_INF);
  if (s1.result == 0) return 0;
  cur[0] = z0 + s1.optvar[0] * d0;
  cur[1] = z1 + s1.optvar[0] * d1;
  act[0] = (int)i;
  if (s1.active_c[0] >= 0 && s1.active_c[0] < k) act[1] = (int)order[s1.active_c[0]];
  else if (s1.active_c[0] >= k && s1.active_c[0] <= k + 3) act[1] = -1 - (s1.active_c[0] - k);
  else return 0;
  return 1;
}

/* Shortcut A: no valid warm-start pair (natural order), objective (+-1e-9, -+1).  Returns the target row or -1. */
static int sm_shortcut_a(const double v[3], int nrows, const double *a, const double *b, const double *c,
                         const double low[2], const double high[2], const double cur[2]) {
  const double sg = v[0] > 0 ? 1.0 : -1.0; /* mirrored variable ua = sg * u: the walk only lowers ua */
  const double x = cur[1], u0 = sg * cur[0];
  int m = -1;
  double um = 0, second = u0;
  for (int j = 0; j < nrows; ++j)
    if (sg * a[j] > LP_TINY) {
      const double uo = sg * (-(b[j] * x + c[j]) / a[j]);
      if (m < 0 || uo < um) { m = j; um = uo; }
    }
  if (m < 0) return -1;
  for (int j = 0; j < nrows; ++j)
    if (sg * a[j] > LP_TINY && j != m) {
      const double uo = sg * (-(b[j] * x + c[j]) / a[j]);
      if (uo < second) second = uo;
    }
  { /* the tightest OTHER bound (or the start value) must violate row m far above the TINY threshold */
    const double au = a[m] * (sg * second), val = au + (b[m] * x + c[m]);
    if (!(val >= sm_gap * (1.0 + fabs(au) + fabs(b[m] * x) + fabs(c[m])))) return -1;
  }
  for (int j = 0; j < nrows; ++j) {
    const double bxc = b[j] * x + c[j];
    if (sg * a[j] > LP_TINY) {
      const double v1d = (-b[j]) * v[0] + a[j] * v[1], ur = -bxc / a[j];
      if (!((fabs(v1d) < LP_TINY) || v1d < 0)) return -1; /* a visit must pick the low end of its line */
      if (!(fabs(x * a[j] - ur * b[j]) < sm_tmax * (a[j] * a[j] + b[j] * b[j]))) return -1; /* line parameter */
    } else if (j < m) {
      if (sg * a[j] < -LP_TINY) { if (sg * (-bxc / a[j]) > um - 1e-9 * (1 + fabs(um))) return -1; }
      else if (bxc > -1e-9 || a[j] != 0.0) return -1;
    }
  }
  if (sg * um < low[0] + 1.0 || sg * um > high[0] - 1.0) return -1;
  return m;
}

/* Shortcut B: valid warm-start pair, order = (p, k, rest), row p violated at the start vertex.  Returns 1 when the
 * optimum of line p inside the box clearly violates row k (then the reference re-solves on position 1 next). */
static int sm_shortcut_b(const double v[3], const double *a, const double *b, const double *c, const double low[2],
                         const double high[2], long p, long k) {
  const double ap = a[p], bp = b[p], cp = c[p];
  if (!(fabs(ap) > 1e-6)) return 0;
  const double ia = 1.0 / ap;
  if (!(low[1] <= high[1] - 1e-7 * (1 + fabs(low[1]) + fabs(high[1])))) return 0;
  const double slope = v[1] - v[0] * bp * ia;
  if (fabs(slope) < 1e-6) return 0;
  /* optimum of line p inside the box = the x bound the objective points to, if u stays well inside its bounds there */
  const double sx = slope > 0 ? high[1] : low[1], su = -(bp * sx + cp) * ia;
  if (!(su >= low[0] + 1.0 && su <= high[0] - 1.0)) return 0;
  if (!(fabs(sx * ap - su * bp) < 1e9 * (ap * ap + bp * bp))) return 0; /* clear of the +-1e10 sentinel */
  const double t1 = a[k] * su, t2 = b[k] * sx, val = t1 + t2 + c[k];
  return val >= sm_gap * (1.0 + fabs(t1) + fabs(t2) + fabs(c[k]));
}

static LpSol sm_lp2d(const double v[3], int nrows, const double *a, const double *b, const double *c,
                     const double low[2], const double high[2], const long active_c[2], long *order, double *a_1d,
                     double *b_1d, long *nres) {
  LpSol sol;
  memset(&sol, 0, sizeof(sol));
  double cur[2];
  int act[2];
  for (int i = 0; i < 2; ++i) {
    if (low[i] > high[i]) return sol;
    if (v[i] > LP_TINY) { cur[i] = high[i]; act[i] = (i == 0) ? -2 : -4; }
    else { cur[i] = low[i]; act[i] = (i == 0) ? -1 : -3; }
  }
  const int valid = active_c[0] >= 0 && active_c[0] < nrows && active_c[1] >= 0 && active_c[1] < nrows &&
                    active_c[0] != active_c[1];
  if (valid) {
    int n = 2;
    order[0] = active_c[1];
    order[1] = active_c[0];
    for (int i = 0; i < nrows; ++i) if (i != active_c[0] && i != active_c[1]) order[n++] = i;
  } else {
    for (int i = 0; i < nrows; ++i) order[i] = i;
  }
  const int objective_ok = (v[0] > LP_TINY && v[1] < 0) || (v[0] < -LP_TINY && v[1] > 0);
  int k = 0, first = 1;
  while (k < nrows) {
    int kk = -1;
    for (int j = k; j < nrows; ++j) {
      const long i = order[j];
      if (!(a[i] * cur[0] + b[i] * cur[1] + c[i] < LP_TINY)) { kk = j; break; }
    }
    if (kk < 0) break;
    int target = kk;
    if (first && objective_ok) {
      if (!valid) {
        const int m = sm_shortcut_a(v, nrows, a, b, c, low, high, cur);
        if (m >= 0) { target = m; ++sm_a_used; } else ++sm_a_abort;
      } else if (kk == 0) {
        if (sm_shortcut_b(v, a, b, c, low, high, order[0], order[1])) { target = 1; ++sm_b_used; } else ++sm_b_abort;
      }
    }
    first = 0;
    ++*nres;
    if (!sm_resolve(v, target, order, a, b, c, low, high, a_1d, b_1d, cur, act)) return sol;
    k = target + 1;
  }
  sol.result = 1;
  sol.optvar[0] = cur[0];
  sol.optvar[1] = cur[1];
  sol.active_c[0] = act[0];
  sol.active_c[1] = act[1];
  return sol;
}

static void sm_hook(const double v[3], int nrows, const double *a, const double *b, const double *c,
                    const double low[2], const double high[2], const long active_in[2], int result,
                    const double optvar[2], const int active_out[2]) {
  long *order = (long *)malloc(sizeof(long) * (nrows + 1));
  double *a1 = (double *)malloc(sizeof(double) * (nrows + 4)), *b1 = (double *)malloc(sizeof(double) * (nrows + 4));
  long nref = 0, nmodel = 0;
  { /* re-solve count of the reference order, for the statistics only */
    long *im = (long *)malloc(sizeof(long) * (nrows + 1));
    (void)lp2d(v, nrows, a, b, c, low, high, active_in, im, a1, b1, &nref);
    free(im);
  }
  const LpSol t = sm_lp2d(v, nrows, a, b, c, low, high, active_in, order, a1, b1, &nmodel);
  free(order); free(a1); free(b1);
  ++sm_lp;
  sm_resolve_ref += nref;
  sm_resolve_model += nmodel;
  if (t.result != result ||
      (result && (memcmp(t.optvar, optvar, 2 * sizeof(double)) || t.active_c[0] != active_out[0] ||
                  t.active_c[1] != active_out[1])))
    ++sm_mismatch;
}

void orc_shortcut_model_enable(int on) { orc_lp2d_hook = on ? sm_hook : NULL; }
void orc_shortcut_model_margins(double gap, double tmax) { sm_gap = gap; sm_tmax = tmax; }
/* out: LPs, mismatches, re-solves (reference), re-solves (model), A used, A declined, B used, B declined */
void orc_shortcut_model_stats(long *out, int reset) {
  out[0] = sm_lp; out[1] = sm_mismatch; out[2] = sm_resolve_ref; out[3] = sm_resolve_model;
  out[4] = sm_a_used; out[5] = sm_a_abort; out[6] = sm_b_used; out[7] = sm_b_abort;
  if (reset) sm_lp = sm_mismatch = sm_resolve_ref = sm_resolve_model = sm_a_used = sm_a_abort = sm_b_used = sm_b_abort = 0;
}
