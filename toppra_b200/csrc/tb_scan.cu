// tb_scan.cu — K2: backward controllable sets + forward parameterisation (TOPP-RA) as a sub-warp cooperative scan.
//
// Replaces (reference, hungpham2511/toppra v0.6.2):
//   ReachabilityAlgorithm.compute_controllable_sets / _one_step   reachability_algorithm.py:166-238
//   ReachabilityAlgorithm.compute_parameterization                reachability_algorithm.py:240-376
//   TOPPRA._forward_step                                          time_optimal_algorithm.py:55-92
//   seidelWrapper.solve_stagewise_optim                           cy_seidel_solverwrapper.pyx:549-697
//   cy_solve_lp2d / cy_solve_lp1d                                 cy_seidel_solverwrapper.pyx:149-390 / 93-144
//
// Design (B200).  The stages of one path are strictly sequential (K[i] <- K[i+1], x[i+1] <- x[i]); the parallelism is
// across paths and across the LP rows of a stage.  A group of LPP lanes (8, 16 or 32) owns one path; a warp carries
// P = 32/LPP paths in lockstep through the stages; every lane holds RPL item slots of its path's stage problem
// (item r = lane_in_group + LPP*s: LP rows 0..nC-1, then the four box rows of the 2-variable LP, then padding).
// The instruction stream that is uniform per path (row search, broadcasts, reductions, shortcut tests, ring
// bookkeeping) is issued once per warp for P paths, and the per-item arithmetic (one IEEE division per projected
// item) runs RPL-deep per lane, which gives each warp instruction-level parallelism instead of idle lanes.
// Round 1 ran one warp per path (tb_scan_v1.cu): 241 k issued instructions per path, issue-bound.
//
// Seidel's incremental 2-variable LP keeps its exact row order (including the reference's warm-start permutation), so
// results are bit-identical to the Cython solver:
//   * "first violated row in order"      -> per-lane order key + group min (xor-shuffle butterfly / redux.sync)
//   * projection of earlier rows + box   -> one fp64 division per item slot
//   * 1-D LP (min of upper / max of lower limits) -> group min of doubles (exact, order independent)
// Control flow is warp-uniform (every collective is executed by all 32 lanes); what differs between the paths of a
// warp is carried in per-group predicates.  The per-stage record (3R+2 doubles) is streamed HBM -> shared memory with
// cp.async.bulk (TMA bulk copy, mbarrier complete_tx) into a ring per path, three stages ahead of the solve; the row
// a re-solve sits on is read back from that shared-memory copy.
// Compiled with -fmad=false: no FMA contraction, same roundings as the x86-64 reference.
#include <limits.h>
#include <stdlib.h>

#include "tb_common.cuh"

extern "C" int tb_scan_ex_v1(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                             const double *sd_start, const double *sd_end, const double *sd_end_hi, int flags, double *K,
                             double *sd, double *u, int *status, int *fail_stage, int *counters, void *stream);

namespace tb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void *src, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_s(const uint32_t addr, unsigned parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok) __nanosleep(64);  // the copy is still in flight: do not burn issue slots other warps could use
  } while (!ok);
}

// Warp-wide min of doubles (no NaNs) with two 32-bit redux.sync: order-preserving map double -> (khi, klo), reduce the
// high words, then the low words among the lanes that tie on the high word.
__device__ __forceinline__ double warp_min(double v) {
  const int hi = __double2hiint(v), lo = __double2loint(v);
  const int m = hi >> 31;  // 0 or -1
  const unsigned khi = (unsigned)(hi ^ (m | (int)0x80000000)), klo = (unsigned)(lo ^ m);
  const unsigned mh = __reduce_min_sync(FULL, khi);
  const unsigned ml = __reduce_min_sync(FULL, khi == mh ? klo : 0xffffffffu);
  const int m2 = ((int)~mh) >> 31;  // -1 if the winner is negative
  return __hiloint2double((int)(mh ^ (unsigned)(m2 | (int)0x80000000)), (int)(ml ^ (unsigned)m2));
}
__device__ __forceinline__ double warp_max(double v) { return -warp_min(-v); }

// ---- group collectives: a group = LPP consecutive lanes (one path); every lane of the WARP executes them -------------
template <int LPP>
__device__ __forceinline__ int grp_min_int(int v) {
  if constexpr (LPP == 32) {
    return __reduce_min_sync(FULL, v);
  } else {
#pragma unroll
    for (int off = LPP / 2; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(FULL, v, off));
    return v;
  }
}
// exact min of doubles (no NaNs).  After the butterfly the lanes of a group agree up to the sign of a zero; the copy of
// the group's first lane is broadcast so that they agree bit for bit.
template <int LPP>
__device__ __forceinline__ double grp_min_f64(double v, const int gbase) {
  if constexpr (LPP == 32) {
    return warp_min(v);
  } else {
#pragma unroll
    for (int off = LPP / 2; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(FULL, v, off);
      v = (o < v) ? o : v;
    }
    return __shfl_sync(FULL, v, gbase);
  }
}
template <int LPP>
__device__ __forceinline__ bool grp_any(const bool p, const unsigned gmask) {
  if constexpr (LPP == 32) {
    return __any_sync(FULL, p);
  } else {
    return (__ballot_sync(FULL, p) & gmask) != 0u;
  }
}

// Identity the optimiser cannot see through: keeps a sanitised division operand from being folded back into the
// original one when the quotient is later replaced by a select (the compiler would divide the raw value again).
__device__ __forceinline__ double opaque(double v) {
  asm volatile("" : "+d"(v));
  return v;
}

// Python's builtin max(a, b) / min(a, b) on floats (reachability_algorithm.py:324-354): a unless b compares beyond it
__device__ __forceinline__ double py_max(const double a, const double b) { return (b > a) ? b : a; }
__device__ __forceinline__ double py_min(const double a, const double b) { return (b < a) ? b : a; }

constexpr int BOXBASE = 1 << 20;
constexpr double SKIP_GAP = 1e-7;   // shortcuts A/B: required violation, relative to the terms' magnitudes (TINY = 1e-10)
constexpr double SKIP_BIG = 1e300;
constexpr double SKIP_TMAX = 90.0;  // shortcut A: largest line parameter of a skipped visit (see lp2d_group)
#ifndef TB_SCAN_NBUF
#define TB_SCAN_NBUF 4
#endif
constexpr int SCAN_NBUF = TB_SCAN_NBUF;  // record buffers per path (NBUF-1 stages of look-ahead)

// Which items this lane's slots hold.  Item r = l + LPP*s: r < nC LP row r; nC <= r < nC+4 box row r-nC; else padding.
// With LPP >= 4 a lane holds at most one box row.
struct SlotMap {
  unsigned rowbits;  // bit s: slot s is an LP row
  int boxslot;       // slot of this lane's box row, -1 if none
  int boxm;          // which box row: 0: low0 <= u, 1: u <= high0, 2: low1 <= x, 3: x <= high1   (pyx:300-318)
};
template <int LPP, int RPL>
__device__ __forceinline__ SlotMap make_slotmap(const int l, const int nC) {
  SlotMap sm{0u, -1, 0};
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const int r = l + LPP * s;
    if (r < nC) sm.rowbits |= 1u << s;
    else if (r < nC + 4) { sm.boxslot = s; sm.boxm = r - nC; }
  }
  return sm;
}
// (a, b, c) of the lane's box row; padding slots are (0, 0, -1)
template <int RPL>
__device__ __forceinline__ void set_box_row(const SlotMap sm, const double low0, const double high0, const double low1,
                                            const double high1, double (&a)[RPL], double (&b)[RPL], double (&c)[RPL]) {
  const int m = sm.boxm;
  const double ba = __hiloint2double((m == 0) ? (int)0xBFF00000 : ((m == 1) ? 0x3FF00000 : 0), 0);
  const double bb = __hiloint2double((m == 2) ? (int)0xBFF00000 : ((m == 3) ? 0x3FF00000 : 0), 0);
  const double bc = (m < 2) ? ((m == 0) ? low0 : -high0) : ((m == 2) ? low1 : -high1);
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const bool isbox = (s == sm.boxslot);
    a[s] = isbox ? ba : a[s];
    b[s] = isbox ? bb : b[s];
    c[s] = isbox ? bc : c[s];
  }
}

// One projected constraint of the 1-D sub-problem (pyx:326-347): its limit on t as an upper bound `thi` (denom >
// TINY) or a lower bound `tlo` (denom < -TINY); +-LP_INF = no limit of that kind (the 1-D LP's own bounds);
// bad: parallel & infeasible.  The divisor of unused slots is replaced by 1 so that the IEEE division never
// leaves its fast path for a value that is thrown away (x/0 would take the slow-path subroutine).
__device__ __forceinline__ void project_item(const bool part, const double aj, const double bj, const double cj,
                                             const double dt0, const double dt1, const double z0, const double z1,
                                             double &thi, double &tlo, bool &bad) {
  const double denom = dt0 * aj + dt1 * bj;
  const double num = cj + z1 * bj + z0 * aj;
  const bool up = part && (denom > LP_TINY), dn = part && (denom < -LP_TINY);
  const double t = -num / ((up || dn) ? denom : 1.0);
  // `cur_x < cur_max` / `cur_x > cur_min` (pyx:115-124): a limit at or beyond the sentinel, or NaN, never wins
  thi = (up && t < LP_INF) ? t : LP_INF;
  tlo = (dn && t > -LP_INF) ? t : -LP_INF;
  bad = bad || (part && !(up || dn) && (num > LP_SMALL));
}

// cy_solve_lp2d (pyx:149-390) on one lane group; the P groups of a warp solve their LPs in lockstep.
//   maximise v0*u + v1*x  s.t.  a u + b x + c <= 0 (rows),  low <= (u, x) <= high (the lane's box-row slot must hold
//   the box row, see set_box_row).  `act`: this group takes part (its outputs are untouched otherwise).
//   ac0/ac1: in = warm-start pair (active_c of the previous solve of this slot), out = new active pair (updated only
//   when feasible, like pyx:673-676,690-691).  Returns false when infeasible (or !act).
//   fetch(r, ak, bk, ck): coefficients of LP row r (group-uniform r) — the re-solve line.
//
// Order: a valid warm-start pair puts active_c[1] first, active_c[0] second, then the remaining rows ascending
// (pyx:252-264) -> order key 0, 1, r + 2.  Per violated row k (one "re-solve") the earlier rows and the four box rows
// are projected onto line k, one item per slot.
// SKIP = the caller is the backward pass of the scan: the shortcuts A / B below may name the first row to re-solve on
// (bit-identical; a scalar model of the rules is checked by tests/test_shortcut_model.py); all else walks the rows in
// order.
template <int LPP, int RPL, bool SKIP, class Fetch>
__device__ __forceinline__ bool lp2d_group(const double v0, const double v1, const double (&a)[RPL],
                                           const double (&b)[RPL], const double (&c)[RPL], const SlotMap sm,
                                           const int l, const int gbase, const unsigned gmask, const int nC,
                                           const double low0, const double high0, const double low1,
                                           const double high1, int &ac0, int &ac1, const bool act, double &out_u,
                                           double &out_x, int &n_resolve, Fetch fetch) {
  bool feas = act && !(low0 > high0 || low1 > high1);  // pyx:233-235
  const bool valid = ac0 >= 0 && ac0 < nC && ac1 >= 0 && ac1 < nC && ac0 != ac1;  // group-uniform
  int pos[RPL];
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const int r = l + LPP * s;
    const bool isrow = (sm.rowbits >> s) & 1u;
    pos[s] = isrow ? ((valid && r == ac1) ? 0 : ((valid && r == ac0) ? 1 : r + 2)) : INT_MAX;
  }
  const int kr0 = ac1, kr1 = ac0;
  auto key_row = [&](const int key) { return (key == 0) ? kr0 : ((key == 1) ? kr1 : key - 2); };
  double p0 = (v0 > LP_TINY) ? high0 : low0;  // pyx:236-247
  double p1 = (v1 > LP_TINY) ? high1 : low1;
  int nac0 = (v0 > LP_TINY) ? -2 : -1;
  int nac1 = (v1 > LP_TINY) ? -4 : -3;
  int kpos = -1, knew = INT_MAX;
  bool running = feas;
  // shortcuts A/B below: only for the two objectives of the backward pass (min x, max x)
  const bool skip_ok = SKIP && (((v0 > LP_TINY) && (v1 < 0)) || ((v0 < -LP_TINY) && (v1 > 0)));
  if constexpr (SKIP) {
    const bool wantA = running && !valid && skip_ok;
    if (__any_sync(FULL, wantA)) {
      // Shortcut A (natural order).  Start vertex = (high0, low1) for the min-x LP, (low0,
      // high1) for the max-x LP.  In mirrored variables (ua = sg*u) every visit of the reference's walk sits on a
      // row that bounds ua from above, lands on x = its box bound and only lowers ua; each visit recomputes the
      // point from scratch over ALL earlier rows, so the final state depends only on the LAST visited row, and
      // that is the row m with the smallest own bound at this x.  The reference is certain to visit m when the
      // smallest bound among the OTHER rows (and the start value) violates row m far above the TINY threshold;
      // one exact re-solve on m then reproduces the reference's state bit for bit, and the exact walk goes on
      // from there.  Rows before m that bound ua from below (or not at all) must hold at the final point with
      // a margin, and every upper row must pick the low end of its line (the exact path's v1d test).  Any doubt
      // -> ordinary walk.  A scalar model of these rules is checked by tests/test_shortcut_model.py.
      const double sg = (v0 > 0) ? 1.0 : -1.0;
      const double x = p1, u0m = sg * p0;
      double uo[RPL], bxc[RPL];
      bool upr[RPL], lor[RPL];
      double lmin = SKIP_BIG;
      bool bad = false;
#pragma unroll
      for (int s = 0; s < RPL; ++s) {
        const bool real = pos[s] != INT_MAX;
        const double sa = sg * a[s];
        bxc[s] = b[s] * x + c[s];
        upr[s] = real && (sa > LP_TINY);
        lor[s] = real && (sa < -LP_TINY);
        // a zero numerator (row 0 at x = 0 with K_lo = 0: every stage) would send the IEEE division through its
        // slow-path subroutine; this value only feeds the margin tests, so 0 is substituted directly
        const bool zn = (bxc[s] == 0.0);
        const double qd = -opaque(zn ? 1.0 : bxc[s]) / ((upr[s] || lor[s]) ? a[s] : 1.0);
        uo[s] = zn ? 0.0 : sg * qd;
        const double v1d_own = (-b[s]) * v0 + a[s] * v1;  // the exact path's v1d if this row were visited
        bad = bad || (upr[s] && !((fabs(v1d_own) < LP_TINY) || (v1d_own < 0)));
        // line parameter t of this row's landing point (own bound, x): a skipped visit must neither end on the
        // +-1e10 sentinel of the 1-D LP nor be far enough from the foot point for a "parallel" row (|denom| <=
        // TINY although the lines cross) to fail the LP_SMALL test there: |t| * TINY stays far below LP_SMALL
        bad = bad || (upr[s] && !(fabs(x * a[s] - (sg * uo[s]) * b[s]) < SKIP_TMAX * (a[s] * a[s] + b[s] * b[s])));
        lmin = (upr[s] && uo[s] < lmin) ? uo[s] : lmin;
      }
      const double um = grp_min_f64<LPP>(lmin, gbase);
      int mp = INT_MAX;
#pragma unroll
      for (int s = 0; s < RPL; ++s) mp = (upr[s] && uo[s] == um) ? min(mp, pos[s]) : mp;
      const int m = grp_min_int<LPP>(mp);
      double l2 = SKIP_BIG;
#pragma unroll
      for (int s = 0; s < RPL; ++s) l2 = (upr[s] && pos[s] != m && uo[s] < l2) ? uo[s] : l2;
      double second = grp_min_f64<LPP>(l2, gbase);
      second = (u0m < second) ? u0m : second;
#pragma unroll
      for (int s = 0; s < RPL; ++s) {
        const double au = a[s] * (sg * second);
        const double val = au + bxc[s];
        const bool bad_m = !(val >= SKIP_GAP * (1.0 + fabs(au) + fabs(b[s] * x) + fabs(c[s])));
        const bool bad_before = (lor[s] && (uo[s] > um - 1e-9 * (1.0 + fabs(um)))) ||
                                (!upr[s] && !lor[s] && ((bxc[s] > -1e-9) || (a[s] != 0.0)));
        bad = bad || ((pos[s] == m) ? bad_m : ((pos[s] < m) && bad_before));
      }
      const double ur = sg * um;
      bad = bad || (ur < low0 + 1.0) || (ur > high0 - 1.0);
      const bool anybad = grp_any<LPP>(bad, gmask);
      knew = (wantA && m != INT_MAX && !anybad) ? m : knew;
    }
  }
  bool first = true;
  while (true) {
    // first row (in order) violated at the current point, pyx:269-275.  NaN counts as violated (not `< TINY`).
    const bool need = running && (knew == INT_MAX);
    if (__any_sync(FULL, need)) {
      int mypos = INT_MAX;
#pragma unroll
      for (int s = 0; s < RPL; ++s) {
        const double val = a[s] * p0 + b[s] * p1 + c[s];
        const bool cand = !(val < LP_TINY) && (pos[s] > kpos) && (pos[s] != INT_MAX);
        mypos = cand ? min(mypos, pos[s]) : mypos;
      }
      const int found = grp_min_int<LPP>(mypos);
      knew = need ? found : knew;
      running = running && !(need && found == INT_MAX);  // no violated row left: this group's LP is solved
    }
    if constexpr (SKIP) {
      if (first) {
        const bool wantB = running && valid && skip_ok && (knew == 0);
        if (__any_sync(FULL, wantB)) {
          // Shortcut B (valid warm-start pair; order = row p = ac1, row k = ac0, the rest).  Row p is violated at
          // the start vertex, so the reference re-solves on it against the box only and holds the optimum of line p
          // inside the box next.  That point is cheap to compute with plain arithmetic; if row k is violated there
          // far above the TINY threshold the reference is certain to re-solve on position 1 next, and that re-solve
          // (row p + the box, recomputed from scratch) does not depend on the skipped one.
          double ap, bp, cp;
          fetch(valid ? kr0 : 0, ap, bp, cp);
          bool okb = fabs(ap) > 1e-6;
          const double ia = 1.0 / (okb ? ap : 1.0);
          okb = okb && (low1 <= high1 - 1e-7 * (1.0 + fabs(low1) + fabs(high1)));
          const double slp = v1 - v0 * bp * ia;  // d objective / dx along line p
          okb = okb && !(fabs(slp) < 1e-6);
          // optimum of line p inside the box: the x bound the objective points to, provided u stays well inside its
          // own bounds there (otherwise the u bounds clip the line first: left to the exact path)
          const double sx = (slp > 0) ? high1 : low1;
          const double su = -(bp * sx + cp) * ia;
          okb = okb && (su >= low0 + 1.0) && (su <= high0 - 1.0);
          // the skipped 1-D optimum must stay clear of the +-1e10 sentinel (pyx:376-383 would report infeasible)
          okb = okb && (fabs(sx * ap - su * bp) < 1e9 * (ap * ap + bp * bp));
          bool kviol = false;
#pragma unroll
          for (int s = 0; s < RPL; ++s) {
            const double t1 = a[s] * su, t2 = b[s] * sx;
            const double val = t1 + t2 + c[s];
            kviol = kviol || ((pos[s] == 1) && (val >= SKIP_GAP * (1.0 + fabs(t1) + fabs(t2) + fabs(c[s]))));
          }
          const bool anyk = grp_any<LPP>(kviol, gmask);
          knew = (wantB && anyk && okb) ? 1 : knew;
        }
      }
    }
    first = false;
    if (!__any_sync(FULL, running)) break;
    // ---- one re-solve on row `knew` for every running group (the others ride along with their results masked) ----
    kpos = running ? knew : kpos;
    const int krow = running ? key_row(knew) : 0;
    n_resolve += running ? 1 : 0;
    double ak, bk, ck;
    fetch(krow, ak, bk, ck);
    // project the origin onto line k, pyx:290-295: z = (-a c, -b c) / (a^2 + b^2).  One division sequence for
    // both components: odd lanes divide the second numerator.
    const double nrm = ak * ak + bk * bk;
    const double zq = ((l & 1) ? (-bk * ck) : (-ak * ck)) / nrm;
    const double z0 = __shfl_sync(FULL, zq, gbase);
    const double z1 = __shfl_sync(FULL, zq, gbase + 1);
    const double dt0 = -bk, dt1 = ak;
    const double v1d = dt0 * v0 + dt1 * v1;
    // project the earlier rows and the four box rows onto the line, pyx:298-347
    double thi[RPL], tlo[RPL];
    int key[RPL];
    bool bad = false;
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
      const bool isbox = (s == sm.boxslot);
      key[s] = isbox ? BOXBASE + sm.boxm : pos[s];
      project_item(isbox || (pos[s] < kpos), a[s], b[s], c[s], dt0, dt1, z0, z1, thi[s], tlo[s], bad);
    }
    // 1-D LP on the line with bounds +-INF, pyx:350 -> cy_solve_lp1d pyx:93-144
    double my_hi = thi[0], my_lo = tlo[0];
#pragma unroll
    for (int s = 1; s < RPL; ++s) {
      my_hi = (thi[s] < my_hi) ? thi[s] : my_hi;
      my_lo = (tlo[s] > my_lo) ? tlo[s] : my_lo;
    }
    // The objective's sign decides which end of [cur_min, cur_max] is the optimum (pyx:130-143); only that end is
    // reduced exactly (max lo = -min(-lo)); "cur_min > cur_max" (pyx:126-128) is a vote against the other side.
    const bool pick_min = (fabs(v1d) < LP_TINY) || (v1d < 0);
    const double red = grp_min_f64<LPP>(pick_min ? -my_lo : my_hi, gbase);
    const double tstar = pick_min ? -red : red;
    const bool cross = pick_min ? (my_hi < tstar) : (my_lo > tstar);
    bool infeas = grp_any<LPP>(bad || cross, gmask);
    // optimum on the +-INF sentinel (1-D active index -1/-2) counts as infeasible, pyx:376-383
    infeas = infeas || (tstar == (pick_min ? -LP_INF : LP_INF));
    // active item = first (lowest key) item that attains the optimum; tstar is finite here, sentinels never match
    int mykey = INT_MAX;
#pragma unroll
    for (int s = 0; s < RPL; ++s) mykey = ((pick_min ? tlo[s] : thi[s]) == tstar) ? min(mykey, key[s]) : mykey;
    const int akey = grp_min_int<LPP>(mykey);
    const int n1 = (akey >= BOXBASE) ? (-1 - (akey - BOXBASE)) : key_row(akey);
    const bool upd = running && !infeas;
    p0 = upd ? (z0 + tstar * dt0) : p0;  // pyx:362-363
    p1 = upd ? (z1 + tstar * dt1) : p1;
    nac0 = upd ? krow : nac0;
    nac1 = upd ? n1 : nac1;
    feas = feas && !(running && infeas);
    running = running && !infeas;
    knew = INT_MAX;
  }
  ac0 = feas ? nac0 : ac0;
  ac1 = feas ? nac1 : ac1;
  out_u = feas ? p0 : out_u;
  out_x = feas ? p1 : out_x;
  return feas;
}

// cy_solve_lp1d (pyx:93-144) as used by the x_min == x_max branch of solve_stagewise_optim (pyx:631-650):
// rows a*u + (b*x + c) <= 0 over ALL nC rows, u in [low0, high0]; objective v0*u.  False if infeasible.
template <int LPP, int RPL>
__device__ __forceinline__ bool lp1d_fixed_x_group(const double v0, const double x, const double (&a)[RPL],
                                                   const double (&b)[RPL], const double (&c)[RPL], const SlotMap sm,
                                                   const int gbase, const unsigned gmask, const double low0,
                                                   const double high0, double &out_u) {
  double my_hi = high0, my_lo = low0;
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const bool isrow = (sm.rowbits >> s) & 1u;
    const double bxc = b[s] * x + c[s];
    const bool up = isrow && (a[s] > LP_TINY), dn = isrow && (a[s] < -LP_TINY);
    // unused slots divide by 1 and a zero numerator (row 0 at x = 0 with K_lo = 0) is not divided at all: both would
    // leave the IEEE division's fast path.  (-bxc) * a is the quotient's correctly signed zero.
    const double den = (up || dn) ? a[s] : 1.0;
    const bool zn = (bxc == 0.0);
    const double q = -opaque(zn ? 1.0 : bxc) / den;
    const double t = zn ? (-bxc) * den : q;
    my_hi = (up && t < my_hi) ? t : my_hi;
    my_lo = (dn && t > my_lo) ? t : my_lo;
  }
  // exact reduction of the optimal end only; infeasibility (cur_min > cur_max) as a vote against the other side
  const bool pick_min = (fabs(v0) < LP_TINY) || (v0 < 0);
  const double red = grp_min_f64<LPP>(pick_min ? -my_lo : my_hi, gbase);
  const double ustar = pick_min ? -red : red;
  const bool infeas = grp_any<LPP>(pick_min ? (my_hi < ustar) : (my_lo > ustar), gmask);
  out_u = ustar;
  return !infeas;
}

// Load this lane's LP rows of one stage record (shared memory) into registers: LP row r >= 2 is static row r-2; rows
// 0, 1 (the x_next rows) are filled by the caller; box / padding slots keep what they hold.
template <int LPP, int RPL>
__device__ __forceinline__ void load_rows(const double *rec, const int R, const int nC, const int l,
                                          double (&a)[RPL], double (&b)[RPL], double (&c)[RPL]) {
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const int r = l + LPP * s;
    const bool in = (r >= 2) && (r < nC);
    const int j = in ? r - 2 : 0;  // always a valid slot of the record: load, then select
    const double va = rec[j], vb = rec[R + j], vc = rec[2 * R + j];
    a[s] = in ? va : a[s];
    b[s] = in ? vb : b[s];
    c[s] = in ? vc : c[s];
  }
}

template <int RPL>
__device__ __forceinline__ void set_xnext_rows(const int l, const double delta, const double xn_min,
                                               const double xn_max, double (&a)[RPL], double (&b)[RPL],
                                               double (&c)[RPL]) {
  // pyx:604-620: row0 = (-2 delta, -1, x_next_min), row1 = (2 delta, 1, -x_next_max); selects, no branches
  const bool xr = l < 2, first = l == 0;
  const double sgn = first ? -1.0 : 1.0;
  a[0] = xr ? sgn * (2 * delta) : a[0];  // -(2 delta) == -2 * delta bit for bit
  b[0] = xr ? sgn : b[0];
  c[0] = xr ? (first ? xn_min : -xn_max) : c[0];
}

// Row fetch for the scan: LP row r of the current stage (group-uniform r) from the shared-memory record, rows 0/1
// synthesised like set_xnext_rows.
struct StageFetch {
  const double *rec;
  int R;
  double two_delta, xn_min, xn_max;
  __device__ __forceinline__ void operator()(const int r, double &ak, double &bk, double &ck) const {
    const int j = (r >= 2) ? r - 2 : 0;
    const double va = rec[j], vb = rec[R + j], vc = rec[2 * R + j];
    const bool xr = r < 2, first = r == 0;
    const double sgn = first ? -1.0 : 1.0;
    ak = xr ? sgn * two_delta : va;
    bk = xr ? sgn : vb;
    ck = xr ? (first ? xn_min : -xn_max) : vc;
  }
};

// CFLAGS >= 0: the scan-mode bits of `flags` (backward-only, forward-only, TOPPRAsd rules) are this compile-time
// constant (the argument's mode bits are ignored), so the unused passes and rules fold away; -1: run time.
// TB_SCAN_FAST_LOWER is always a run-time bit.
template <int LPP, int RPL, int CFLAGS>
__global__ void __launch_bounds__(32)
scan_kernel(const double *__restrict__ records, const int W, const int R, const double *__restrict__ grid,
            const int grid_shared, const int B, const int G, const double *__restrict__ sd_start,
            const double *__restrict__ sd_end, const double *__restrict__ sd_end_hi, const int flags_arg,
            double *__restrict__ Kout, double *__restrict__ sdout, double *__restrict__ uout,
            int *__restrict__ status, int *__restrict__ fail_stage, int *__restrict__ counters) {
  constexpr int P = 32 / LPP;  // paths per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = (int)threadIdx.x;
  const int g = lane / LPP, l = lane % LPP, gbase = g * LPP;
  const unsigned gmask = (LPP == 32) ? FULL : (((1u << (LPP & 31)) - 1u) << gbase);
  const long path_raw = (long)blockIdx.x * P + g;
  const bool wr = path_raw < B;                 // groups past the end of the batch shadow the last path, writing nothing
  const long path = wr ? path_raw : (long)B - 1;
  double *bufs = reinterpret_cast<double *>(smem_raw) + (size_t)g * SCAN_NBUF * W;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)P * SCAN_NBUF * W * sizeof(double)) + g * SCAN_NBUF;
  const int N = G - 1, nC = R + 2;
  const unsigned rec_bytes = (unsigned)(W * sizeof(double));
  const double *rec_path = records + (size_t)path * G * W;
  const double *gp = grid + (grid_shared ? 0 : (size_t)path * G);
  double *Kp = Kout + (size_t)path * G * 2;
  const int flags = (CFLAGS >= 0) ? CFLAGS : flags_arg;
  const bool backward_only = (flags & TB_SCAN_BACKWARD_ONLY) != 0;  // compute_controllable_sets(sdmin, sdmax) alone
  const bool forward_only = (flags & TB_SCAN_FORWARD_ONLY) != 0;    // K and status come from an earlier backward-only launch
  const bool fast_lower = (flags_arg & TB_SCAN_FAST_LOWER) != 0;    // opt-in shortcut for the min-x LP (not bit-identical)
  const bool sd_mode = (flags & TB_SCAN_SD_FORWARD) != 0;           // TOPPRAsd forward-pass rules (no retry, x_next - 1e-5 clip)
  const bool sd_slow = (flags & TB_SCAN_SD_SLOW) != 0;              // TOPPRAsd slowest pass: minimise the next velocity
  double *sdp = backward_only ? nullptr : sdout + (size_t)path * G;
  double *up = backward_only ? nullptr : uout + (size_t)path * (G > 1 ? G - 1 : 0);

  if (l == 0) {
#pragma unroll
    for (int q = 0; q < SCAN_NBUF; ++q) mbar_init(&bars[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  // Ring of SCAN_NBUF record buffers per path: up to SCAN_NBUF-1 bulk copies in flight per group.  The groups of a
  // warp move through the stages in lockstep, so the ring counters are warp-uniform.
  unsigned n_issued = 0, n_waited = 0;
  const uint32_t bufs_s = smem_u32(bufs), bars_s = smem_u32(bars);
  auto issue = [&](int stage) {
    if (l == 0) {
      const unsigned q = n_issued % SCAN_NBUF;
      mbar_expect_tx_s(bars_s + q * 8u, rec_bytes);
      bulk_g2s_s(bufs_s + q * rec_bytes, rec_path + (size_t)stage * W, rec_bytes, bars_s + q * 8u);
    }
    ++n_issued;
  };
  auto acquire = [&]() -> const double * {
    const unsigned q = n_waited % SCAN_NBUF;
    mbar_wait_s(bars_s + q * 8u, (n_waited / SCAN_NBUF) & 1);
    __syncwarp();
    ++n_waited;
    return bufs + (size_t)q * W;
  };
  constexpr int AHEAD = SCAN_NBUF - 1;

  // instrumentation: projected re-solves, retries, fast-mode stages; the LP counts are derived from the stage counts
  int n_resolve = 0, n_retry = 0, n_fast = 0;
  const SlotMap sm = make_slotmap<LPP, RPL>(l, nC);
  double a[RPL], b[RPL], c[RPL];
#pragma unroll
  for (int s = 0; s < RPL; ++s) { a[s] = 0.0; b[s] = 0.0; c[s] = -1.0; }  // padding rows are (0, 0, -1)
  const double nan_d = __longlong_as_double(0x7ff8000000000000LL);

  // ---------------- backward pass: controllable sets, reachability_algorithm.py:166-238 ----------------
  const double sde = sd_end ? sd_end[path] : 0.0;
  const double sds = sd_start ? sd_start[path] : 0.0;
  const double sdeh = sd_end_hi ? sd_end_hi[path] : sde;
  double kn0 = sde * sde, kn1 = sdeh * sdeh;  // K[N] = [sdmin^2, sdmax^2], reachability_algorithm.py:185
  if (l == 0 && wr && !forward_only) { Kp[2 * N] = kn0; Kp[2 * N + 1] = kn1; }
  int st = TB_STATUS_OK, fstage = -1;
  int up0 = 0, up1 = 0, dn0 = 0, dn1 = 0;  // active_c_up / active_c_down, initialised to zeros (pyx:526-527)
  if (forward_only) {
    st = status[path];
    fstage = fail_stage ? fail_stage[path] : -1;
    kn0 = Kp[0];
    kn1 = Kp[1];
  }
  bool alive = !forward_only;  // this group is still in the backward recursion
  for (int q = 0; !forward_only && q < AHEAD && N - 1 - q >= 0; ++q) issue(N - 1 - q);
  for (int i = forward_only ? -1 : N - 1; i >= 0; --i) {
    const double *rec = acquire();
    load_rows<LPP, RPL>(rec, R, nC, l, a, b, c);
    const double xlo = rec[3 * R], xhi = rec[3 * R + 1];
    if (i - AHEAD >= 0) issue(i - AHEAD);  // into the buffer of stage i+1: every lane passed acquire()'s __syncwarp after its last read
    const double delta = gp[i + 1] - gp[i];
    set_xnext_rows<RPL>(l, delta, kn0, kn1, a, b, c);
    set_box_row<RPL>(sm, VAR_MIN, VAR_MAX, xlo, xhi, a, b, c);  // low/high: pyx:587-601 with x_min = x_max = NaN
    const StageFetch fetch{rec, R, 2 * delta, kn0, kn1};
    // The two LPs of the stage share ONE copy of the solver code (the unrolled slot loops make it large; two inlined
    // copies overflow the instruction cache): which = 0: x_upper, g = (1e-9, -1) -> v = (-1e-9, 1), slot
    // active_c_down (g[1] <= 0), reachability_algorithm.py:229-233; which = 1: x_lower, g = (-1e-9, 1) ->
    // v = (1e-9, -1), slot active_c_up, :234-236.
    bool fast_hit = false;
    if (fast_lower) {
      // TB_SCAN_FAST_LOWER: some u is feasible at x = xlo, so min x IS xlo.  The reference reaches the same vertex
      // through ~4 projected re-solves and returns xlo plus rounding noise of its projection arithmetic
      // (|noise| <= ~1e-16, 5 % of the stages): this shortcut is exact for the LP, not bit-identical to that noise.
      double ufeas;
      const bool okf = lp1d_fixed_x_group<LPP, RPL>(1.0, xlo, a, b, c, sm, gbase, gmask, VAR_MIN, VAR_MAX, ufeas);
      fast_hit = alive && okf && (xlo <= xhi);
      n_fast += fast_hit ? 1 : 0;
    }
    bool ok_hi = false, ok_lo2 = false;
    double x_upper = nan_d, x_low2 = nan_d;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const double sgnv = which ? -1.0 : 1.0;
      int w0 = which ? up0 : dn0, w1 = which ? up1 : dn1;
      double uu = 0.0, xx = 0.0;
      const bool okw = lp2d_group<LPP, RPL, true>(sgnv * -1e-9, sgnv, a, b, c, sm, l, gbase, gmask, nC, VAR_MIN, VAR_MAX,
                                                  xlo, xhi, w0, w1, which ? (alive && !fast_hit) : alive, uu, xx,
                                                  n_resolve, fetch);
      if (which) { up0 = w0; up1 = w1; ok_lo2 = okw; x_low2 = okw ? xx : nan_d; }
      else { dn0 = w0; dn1 = w1; ok_hi = okw; x_upper = okw ? xx : nan_d; }
    }
    const bool ok_lo = fast_hit || ok_lo2;
    double x_lower = fast_hit ? xlo : x_low2;
    if (x_lower < 0) x_lower = 0;  // reachability_algorithm.py:190-191
    if (l == 0 && wr && alive) { Kp[2 * i] = x_lower; Kp[2 * i + 1] = x_upper; }
    if (alive && !(ok_hi && ok_lo)) {
      // reachability_algorithm.py:192-197: stop; the remaining K entries stay 0 (np.zeros)
      st = TB_STATUS_FAIL_UNCONTROLLABLE;
      fstage = i;
      if (wr) for (int j = l; j < 2 * i; j += LPP) Kp[j] = 0.0;
      alive = false;
    }
    kn0 = alive ? x_lower : kn0;
    kn1 = alive ? x_upper : kn1;
    if (!__any_sync(FULL, alive)) break;
  }
  if (counters && l == 0 && wr && !forward_only) {
    // backward stages entered: N, or N - fstage when stage fstage failed; 2 LPs each (fast mode: n_fast of them 1-variable)
    const int nb = (st == TB_STATUS_OK) ? N : N - fstage;
    counters[path * 4 + 0] = 2 * nb - n_fast;
    counters[path * 4 + 1] = n_fast;
  }
  // drain prefetches that were issued but not consumed (failure path), so the buffers can be reused
  while (n_waited < n_issued) (void)acquire();
  __syncwarp();

  const double x_start = sds * sds;
  if (backward_only) {
    if (l == 0 && wr) {
      status[path] = st;
      if (fail_stage) fail_stage[path] = fstage;
    }
    return;
  }
  if (st == TB_STATUS_OK) {
    // kn0,kn1 == K[0]; admissibility check reachability_algorithm.py:290-301
    if (x_start + ALG_SMALL < kn0 || kn1 + ALG_SMALL < x_start) { st = TB_STATUS_FAIL_UNCONTROLLABLE; fstage = 0; }
  }
  const bool entered = (st == TB_STATUS_OK);  // this group runs the forward pass
  if (!entered && wr) {
    for (int j = l; j < G; j += LPP) sdp[j] = nan_d;
    for (int j = l; j < N; j += LPP) up[j] = nan_d;
  }
  if (__any_sync(FULL, entered)) {
    // ---------------- forward pass, reachability_algorithm.py:303-364 ----------------
    // sd = sqrt(x) is applied in one sweep after the pass; until then sd[] holds x
    bool fw = entered;
    double x = x_start;
    if (l == 0 && wr && fw) sdp[0] = x;
    for (int q = 0; q < AHEAD && q < N; ++q) issue(q);
    for (int i = 0; i < N; ++i) {
      const double *rec = acquire();
      load_rows<LPP, RPL>(rec, R, nC, l, a, b, c);
      if (i + AHEAD < N) issue(i + AHEAD);
      const double delta = gp[i + 1] - gp[i];
      const double k0 = Kp[2 * (i + 1)], k1 = Kp[2 * (i + 1) + 1];
      set_xnext_rows<RPL>(l, delta, k0, k1, a, b, c);
      int tries = 0;
      bool ok = false, pending = fw;
      double uopt = 0.0;
      while (true) {
        // _forward_step: g = (-2 delta, -1), x_min = x_max = x -> 1-D branch, v0 = 2 delta (pyx:628-636);
        // TOPPRAsd's slowest pass uses g = (2 delta, 1) (desired_duration_algorithm.py:218-223)
        double ucand;
        const bool okk = lp1d_fixed_x_group<LPP, RPL>(sd_slow ? -(2 * delta) : -(-2 * delta), x, a, b, c, sm, gbase,
                                                      gmask, VAR_MIN, VAR_MAX, ucand);
        if (pending) {
          ok = okk;
          uopt = ucand;
          if (ok || sd_mode || tries >= MAX_TRIES) {  // TOPPRAsd has no retry rule
            pending = false;
          } else {
            x = py_max(x - ALG_TINY, 0.999 * x);  // reachability_algorithm.py:324-327
            ++tries;
            ++n_retry;
          }
        }
        if (!__any_sync(FULL, pending)) break;
      }
      if (fw && !ok) {
        // reachability_algorithm.py:337-342: xs[i+1:] = nan -> sd NaN -> ErrUnknown; us stay 0
        // (TOPPRAsd: us[i:] and xs[i+1:] become NaN, desired_duration_algorithm.py:106-111)
        st = TB_STATUS_ERR_UNKNOWN;
        fstage = i;
        if (wr) {
          if (l == 0) sdp[i] = x;
          for (int j = i + 1 + l; j < G; j += LPP) sdp[j] = nan_d;
          for (int j = i + l; j < N; j += LPP) up[j] = sd_mode ? nan_d : 0.0;
        }
        fw = false;
      }
      if (fw) {
        double x_next = x + 2 * delta * uopt;                       // reachability_algorithm.py:352
        if (sd_mode) {
          x_next = py_min(k1, py_max(k0, x_next - ALG_SMALL));        // desired_duration_algorithm.py:117
        } else {
          x_next = py_max(x_next - ALG_TINY, 0.9999 * x_next);      // :353
          x_next = py_min(k1, py_max(k0, x_next));                    // :354
        }
        if (l == 0 && wr) {
          up[i] = uopt;
          if (tries) sdp[i] = x;  // x was shrunk by the retry rule
          sdp[i + 1] = x_next;
        }
        x = x_next;
      }
      if (!__any_sync(FULL, fw)) break;
    }
    while (n_waited < n_issued) (void)acquire();
    __syncwarp();
    if (!sd_mode && entered && wr)  // TOPPRAsd combines the squared velocities: its passes return x = sd^2
      for (int j = l; j < G; j += LPP) sdp[j] = sqrt(sdp[j]);  // reachability_algorithm.py:365
  }
  if (l == 0 && wr) {
    status[path] = st;
    if (fail_stage) fail_stage[path] = fstage;
    if (counters) {
      // forward: one 1-variable LP per stage entered (N, or fstage + 1 when stage fstage failed) + one per retry;
      // none when the path failed before the forward pass
      const int n_fwd_stages = (st == TB_STATUS_OK) ? N : ((st == TB_STATUS_ERR_UNKNOWN) ? fstage + 1 : -1);
      if (forward_only) { counters[path * 4 + 0] = 0; counters[path * 4 + 1] = 0; }
      if (n_fwd_stages >= 0) counters[path * 4 + 1] += n_fwd_stages + n_retry;
      counters[path * 4 + 2] = n_resolve;
      counters[path * 4 + 3] = n_retry;
    }
  }
}

// Row fetch from plain global arrays a/b/c[n] (feasible sets read the record in HBM, the LP shim its own arrays).
struct ArrayFetch {
  const double *pa, *pb, *pc;
  __device__ __forceinline__ void operator()(const int r, double &ak, double &bk, double &ck) const {
    ak = pa[r]; bk = pb[r]; ck = pc[r];
  }
};
// feasible sets: LP rows 0/1 are the x_next rows for the whole x_next box, or (0,0,-1) at the last gridpoint
struct FeasFetch {
  const double *rec;
  int R;
  double two_delta;
  bool last;
  __device__ __forceinline__ void operator()(const int r, double &ak, double &bk, double &ck) const {
    const int j = (r >= 2) ? r - 2 : 0;
    const double va = rec[j], vb = rec[R + j], vc = rec[2 * R + j];
    const bool xr = r < 2, first = r == 0;
    const double sgn = first ? -1.0 : 1.0;
    ak = xr ? (last ? 0.0 : sgn * two_delta) : va;
    bk = xr ? (last ? 0.0 : sgn) : vb;
    ck = xr ? (last ? -1.0 : (first ? -CVXPY_MAXX : -CVXPY_MAXX)) : vc;
  }
};

// compute_feasible_sets, reachability_algorithm.py:131-164: X[i] = [min x, max x] over stage i alone
// (x in [-1e4, 1e4], x_next in [-1e4, 1e4]); warm-start slots chained over i like the reference.
template <int LPP, int RPL>
__global__ void __launch_bounds__(32)
feasible_kernel(const double *__restrict__ records, const int W, const int R, const double *__restrict__ grid,
                const int grid_shared, const int B, const int G, double *__restrict__ Xout) {
  constexpr int P = 32 / LPP;
  const int lane = (int)threadIdx.x;
  const int g = lane / LPP, l = lane % LPP, gbase = g * LPP;
  const unsigned gmask = (LPP == 32) ? FULL : (((1u << (LPP & 31)) - 1u) << gbase);
  const long path_raw = (long)blockIdx.x * P + g;
  const bool wr = path_raw < B;
  const long path = wr ? path_raw : (long)B - 1;
  const int N = G - 1, nC = R + 2;
  const double *rec_path = records + (size_t)path * G * W;
  const double *gp = grid + (grid_shared ? 0 : (size_t)path * G);
  double *Xp = Xout + (size_t)path * G * 2;
  const SlotMap sm = make_slotmap<LPP, RPL>(l, nC);
  double a[RPL], b[RPL], c[RPL];
#pragma unroll
  for (int s = 0; s < RPL; ++s) { a[s] = 0.0; b[s] = 0.0; c[s] = -1.0; }
  int up0 = 0, up1 = 0, dn0 = 0, dn1 = 0, n_resolve = 0;
  const double nan_d = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = 0; i <= N; ++i) {
    const double *rec = rec_path + (size_t)i * W;
    load_rows<LPP, RPL>(rec, R, nC, l, a, b, c);
    const double xlo = fmax(rec[3 * R], -CVXPY_MAXX), xhi = fmin(rec[3 * R + 1], CVXPY_MAXX);  // pyx:598-601
    const bool last = (i == N);
    const double delta = last ? 0.0 : gp[i + 1] - gp[i];
    if (!last) {
      set_xnext_rows<RPL>(l, delta, -CVXPY_MAXX, CVXPY_MAXX, a, b, c);
    } else if (l < 2) {  // i == N: rows 0,1 are (0,0,-1), pyx:621-625
      a[0] = 0.0; b[0] = 0.0; c[0] = -1.0;
    }
    set_box_row<RPL>(sm, VAR_MIN, VAR_MAX, xlo, xhi, a, b, c);
    const FeasFetch fetch{rec, R, 2 * delta, last};
    double uu = 0.0, xx = 0.0;
    // g_lower = (1e-9, 1): g[1] > 0 -> slot up; v = (-1e-9, -1)
    const bool ok0 = lp2d_group<LPP, RPL, false>(-1e-9, -1.0, a, b, c, sm, l, gbase, gmask, nC, VAR_MIN, VAR_MAX, xlo,
                                                 xhi, up0, up1, true, uu, xx, n_resolve, fetch);
    double x0 = ok0 ? xx : nan_d;
    const bool ok1 = lp2d_group<LPP, RPL, false>(1e-9, 1.0, a, b, c, sm, l, gbase, gmask, nC, VAR_MIN, VAR_MAX, xlo, xhi,
                                                 dn0, dn1, true, uu, xx, n_resolve, fetch);
    const double x1 = ok1 ? xx : nan_d;
    if (x0 < 0) x0 = 0;  // reachability_algorithm.py:160-162
    if (l == 0 && wr) { Xp[2 * i] = x0; Xp[2 * i + 1] = x1; }
  }
}

// Batched stand-alone LPs (one warp per LP): the device counterparts of the reference's Python shims
// solve_lp2d / solve_lp1d (cy_seidel_solverwrapper.pyx:42-87).  Used by B200SolverWrapper.solve_stagewise_optim
// and by the LP-level known-answer / differential tests.
template <int RPL>
__global__ void __launch_bounds__(32)
lp2d_batch_kernel(const double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ b,
                  const double *__restrict__ c, const double *__restrict__ low, const double *__restrict__ high,
                  const int *__restrict__ active_in, const int B, const int n, int *__restrict__ result,
                  double *__restrict__ optval, double *__restrict__ optvar, int *__restrict__ active_out) {
  const int lane = (int)threadIdx.x;
  const long p = (long)blockIdx.x;
  if (p >= B) return;
  const SlotMap sm = make_slotmap<32, RPL>(lane, n);
  double ra[RPL], rb[RPL], rc[RPL];
#pragma unroll
  for (int s = 0; s < RPL; ++s) {
    const int r = lane + 32 * s;
    if (r < n) { ra[s] = a[p * n + r]; rb[s] = b[p * n + r]; rc[s] = c[p * n + r]; }
    else { ra[s] = 0.0; rb[s] = 0.0; rc[s] = -1.0; }
  }
  const double l0 = low[p * 2], h0 = high[p * 2], l1 = low[p * 2 + 1], h1 = high[p * 2 + 1];
  set_box_row<RPL>(sm, l0, h0, l1, h1, ra, rb, rc);
  int ac0 = active_in ? active_in[p * 2] : 0, ac1 = active_in ? active_in[p * 2 + 1] : 0;
  int n_resolve = 0;
  double uu = 0.0, xx = 0.0;
  const double v0 = v[p * 3], v1 = v[p * 3 + 1], v2 = v[p * 3 + 2];
  const ArrayFetch fetch{a + p * n, b + p * n, c + p * n};
  const bool ok = lp2d_group<32, RPL, false>(v0, v1, ra, rb, rc, sm, lane, 0, FULL, n, l0, h0, l1, h1, ac0, ac1, true, uu,
                                             xx, n_resolve, fetch);
  if (lane == 0) {
    result[p] = ok ? 1 : 0;
    const double nan_d = __longlong_as_double(0x7ff8000000000000LL);
    optvar[p * 2] = ok ? uu : nan_d;
    optvar[p * 2 + 1] = ok ? xx : nan_d;
    optval[p] = ok ? (uu * v0 + xx * v1 + v2) : nan_d;  // pyx:389
    active_out[p * 2] = ac0;
    active_out[p * 2 + 1] = ac1;
  }
}

// cy_solve_lp1d with the active index (pyx:93-144): max v0 x + v1, a x + b <= 0, low <= x <= high.
__global__ void lp1d_batch_kernel(const double *__restrict__ v, const double *__restrict__ a,
                                  const double *__restrict__ b, const double *__restrict__ low,
                                  const double *__restrict__ high, const int B, const int n,
                                  int *__restrict__ result, double *__restrict__ optval,
                                  double *__restrict__ optvar, int *__restrict__ active_out) {
  const int lane = threadIdx.x & 31;
  const long p = (long)blockIdx.x;
  if (p >= B) return;
  double my_hi = high[p], my_lo = low[p];
  int hi_idx = INT_MAX, lo_idx = INT_MAX;  // first row index attaining the bound (strict improvement only)
  for (int r = lane; r < n; r += 32) {
    const double ar = a[p * n + r], br = b[p * n + r];
    if (ar > LP_TINY) {
      const double cx = -br / ar;
      if (cx < my_hi) { my_hi = cx; hi_idx = r; }
    } else if (ar < -LP_TINY) {
      const double cx = -br / ar;
      if (cx > my_lo) { my_lo = cx; lo_idx = r; }
    }
  }
  const double cur_max = warp_min(my_hi), cur_min = warp_max(my_lo);
  // sequential semantics: the first row (lowest index) that reaches the final value wins; -2/-1 if none improved
  int hk = (hi_idx != INT_MAX && my_hi == cur_max) ? hi_idx : INT_MAX;
  int lk = (lo_idx != INT_MAX && my_lo == cur_min) ? lo_idx : INT_MAX;
  hk = __reduce_min_sync(FULL, hk);
  lk = __reduce_min_sync(FULL, lk);
  if (lane == 0) {
    const double nan_d = __longlong_as_double(0x7ff8000000000000LL);
    const double v0 = v[p * 2], v1 = v[p * 2 + 1];
    if (cur_min > cur_max) {
      result[p] = 0; optval[p] = nan_d; optvar[p] = nan_d; active_out[p] = 0;
    } else if (fabs(v0) < LP_TINY || v0 < 0) {
      result[p] = 1; optvar[p] = cur_min; optval[p] = v0 * cur_min + v1; active_out[p] = (lk == INT_MAX) ? -1 : lk;
    } else {
      result[p] = 1; optvar[p] = cur_max; optval[p] = v0 * cur_max + v1; active_out[p] = (hk == INT_MAX) ? -2 : hk;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// launch: the group shape (LPP lanes per path x RPL item slots per lane) is chosen from the item count nC + 4
// ---------------------------------------------------------------------------------------------------------------
template <int LPP, int RPL, int CFLAGS>
int launch_scan_shape(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                      const double *sd_start, const double *sd_end, const double *sd_end_hi, int flags, double *K,
                      double *sd, double *u, int *status, int *fail_stage, int *counters, cudaStream_t stream) {
  constexpr int P = 32 / LPP;
  const size_t smem = (size_t)P * SCAN_NBUF * W * sizeof(double) + (size_t)P * SCAN_NBUF * sizeof(uint64_t);
  auto kern = scan_kernel<LPP, RPL, CFLAGS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tb_scan: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const int blocks = (B + P - 1) / P;
  kern<<<blocks, 32, smem, stream>>>(records, W, R, grid, grid_shared, B, G, sd_start, sd_end, sd_end_hi, flags, K, sd, u,
                                     status, fail_stage, counters);
  return check_launch("tb_scan");
}

#define TB_SCAN_ARGS records, W, R, grid, grid_shared, B, G, sd_start, sd_end, sd_end_hi, flags, K, sd, u, status, fail_stage, counters, stream

// the three launch kinds of the batched solver get their own instantiation (full scan, backward only, forward only)
template <int LPP, int RPL>
int launch_scan_modes(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                      const double *sd_start, const double *sd_end, const double *sd_end_hi, int flags, double *K,
                      double *sd, double *u, int *status, int *fail_stage, int *counters, cudaStream_t stream) {
  const int mode = flags & (TB_SCAN_BACKWARD_ONLY | TB_SCAN_SD_FORWARD | TB_SCAN_SD_SLOW | TB_SCAN_FORWARD_ONLY);
  if (mode == 0) return launch_scan_shape<LPP, RPL, 0>(TB_SCAN_ARGS);
  if (mode == TB_SCAN_BACKWARD_ONLY) return launch_scan_shape<LPP, RPL, TB_SCAN_BACKWARD_ONLY>(TB_SCAN_ARGS);
  if (mode == TB_SCAN_FORWARD_ONLY) return launch_scan_shape<LPP, RPL, TB_SCAN_FORWARD_ONLY>(TB_SCAN_ARGS);
  return launch_scan_shape<LPP, RPL, -1>(TB_SCAN_ARGS);
}

int check_scan_args(const char *fn, const void *records, int W, int R, const void *grid, int B, int G) {
  if (!records || !grid || B <= 0 || G <= 0 || R < 0) { set_error("%s: bad argument", fn); return TB_ERR_ARG; }
  if (R > MAX_ROWS) { set_error("%s: R=%d > %d rows", fn, R, MAX_ROWS); return TB_ERR_UNSUPPORTED; }
  if (W < 3 * R + 2 || (W & 1)) { set_error("%s: record stride W=%d must be even and >= 3R+2", fn, W); return TB_ERR_ALIGN; }
  if (((uintptr_t)records & 15) != 0) { set_error("%s: records not 16-byte aligned", fn); return TB_ERR_ALIGN; }
  return 0;
}

}  // namespace
}  // namespace tb

extern "C" int tb_scan_ex(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                          const double *sd_start, const double *sd_end, const double *sd_end_hi, int flags, double *K,
                          double *sd, double *u, int *status, int *fail_stage, int *counters, void *stream_) {
  using namespace tb;
  static const char *impl_env = getenv("TB_SCAN_IMPL");
  if (impl_env && impl_env[0] == 'v' && impl_env[1] == '1')
    return tb_scan_ex_v1(records, W, R, grid, grid_shared, B, G, sd_start, sd_end, sd_end_hi, flags, K, sd, u, status,
                         fail_stage, counters, stream_);
  int rc = check_scan_args("tb_scan", records, W, R, grid, B, G);
  if (rc) return rc;
  const bool backward_only = (flags & TB_SCAN_BACKWARD_ONLY) != 0;
  if (!K || !status || (!backward_only && (!sd || (G > 1 && !u)))) { set_error("tb_scan: null output"); return TB_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int items = R + 2 + 4;  // LP rows + the four box rows
  static const char *lpp_env = getenv("TB_SCAN_LPP");  // tuning override: lanes per path (8, 16, 32)
  const int want = lpp_env ? atoi(lpp_env) : 0;
#ifdef TB_SCAN_TUNE
  if (want == 16 && items <= 48) return launch_scan_modes<16, 3>(TB_SCAN_ARGS);
  if (want == 32 && items <= 64) return launch_scan_modes<32, 2>(TB_SCAN_ARGS);
  if (want == 4 && items <= 36) return launch_scan_modes<4, 9>(TB_SCAN_ARGS);
#endif
  (void)want;
  if (items <= 32) return launch_scan_modes<8, 4>(TB_SCAN_ARGS);
  if (items <= 40) return launch_scan_modes<8, 5>(TB_SCAN_ARGS);
  if (items <= 64) return launch_scan_modes<16, 4>(TB_SCAN_ARGS);
  if (items <= 80) return launch_scan_shape<16, 5, -1>(TB_SCAN_ARGS);
  if (items <= 128) return launch_scan_shape<32, 4, -1>(TB_SCAN_ARGS);
  return launch_scan_shape<32, 5, -1>(TB_SCAN_ARGS);
}

extern "C" int tb_scan(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                       const double *sd_start, const double *sd_end, double *K, double *sd, double *u, int *status,
                       int *fail_stage, void *stream) {
  return tb_scan_ex(records, W, R, grid, grid_shared, B, G, sd_start, sd_end, nullptr, 0, K, sd, u, status, fail_stage,
                    nullptr, stream);
}

extern "C" int tb_feasible_sets(const double *records, int W, int R, const double *grid, int grid_shared, int B, int G,
                                double *X, void *stream) {
  using namespace tb;
  int rc = check_scan_args("tb_feasible_sets", records, W, R, grid, B, G);
  if (rc) return rc;
  if (!X) { set_error("tb_feasible_sets: null output"); return TB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  const int items = R + 2 + 4;
#define TB_LAUNCH_FEAS(LPP, RPL) \
  feasible_kernel<LPP, RPL><<<(B + (32 / LPP) - 1) / (32 / LPP), 32, 0, s>>>(records, W, R, grid, grid_shared, B, G, X)
  if (items <= 40) TB_LAUNCH_FEAS(8, 5);
  else if (items <= 80) TB_LAUNCH_FEAS(16, 5);
  else TB_LAUNCH_FEAS(32, 5);
#undef TB_LAUNCH_FEAS
  return check_launch("tb_feasible_sets");
}

extern "C" int tb_lp2d_batch(const double *v, const double *a, const double *b, const double *c, const double *low,
                             const double *high, const int *active_in, int B, int n, int *result, double *optval,
                             double *optvar, int *active_out, void *stream) {
  using namespace tb;
  if (!v || !low || !high || !result || !optval || !optvar || !active_out || B <= 0 || n < 0 || (n > 0 && (!a || !b || !c))) {
    set_error("tb_lp2d_batch: bad argument");
    return TB_ERR_ARG;
  }
  if (n > MAX_ROWS + 2) { set_error("tb_lp2d_batch: n=%d > %d rows", n, MAX_ROWS + 2); return TB_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
#define TB_LAUNCH_LP2D(RPL) \
  lp2d_batch_kernel<RPL><<<B, 32, 0, s>>>(v, a, b, c, low, high, active_in, B, n, result, optval, optvar, active_out)
  if (n + 4 <= 64) TB_LAUNCH_LP2D(2);
  else if (n + 4 <= 96) TB_LAUNCH_LP2D(3);
  else if (n + 4 <= 128) TB_LAUNCH_LP2D(4);
  else TB_LAUNCH_LP2D(5);
#undef TB_LAUNCH_LP2D
  return check_launch("tb_lp2d_batch");
}

extern "C" int tb_lp1d_batch(const double *v, const double *a, const double *b, const double *low, const double *high,
                             int B, int n, int *result, double *optval, double *optvar, int *active_out,
                             void *stream) {
  using namespace tb;
  if (!v || !low || !high || !result || !optval || !optvar || !active_out || B <= 0 || n < 0 || (n > 0 && (!a || !b))) {
    set_error("tb_lp1d_batch: bad argument");
    return TB_ERR_ARG;
  }
  lp1d_batch_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(v, a, b, low, high, B, n, result, optval, optvar, active_out);
  return check_launch("tb_lp1d_batch");
}
