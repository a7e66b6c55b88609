// bp_lp.cuh -- K6: do two convex sets {A1 x <= b1}, {A2 x <= b2} intersect
// after shrinking every face by tol?
//
// Replaces BoundPlanner.set_intersection (BoundPlanner.py:774-787): a HiGHS
// feasibility LP with c = 0 over the stacked rows, called from add_edges with
// tol = 0.01 (:796-798).  The answer is the sign of
//     s* = min_x max_i (a_i . x - (b_i - tol))        (rows with a_i != 0)
// (intersects <=> s* <= 0; any positive row weights give the same sign, so the
// rows are NOT normalised).  s* is bracketed from both sides by a phase-I
// log-barrier Newton method in (x, s) in R^4:
//   upper bound: s_ub = max_i (a_i.x - c_i) at the current iterate (exact
//                evaluation; s_ub <= 0 is a feasible point -> "intersects");
//   lower bound: weak duality with the barrier multipliers, padded by
//                ||sum lam_i a_i|| * BP_LP_DIAMETER for the dual residual
//                (lb > 0 -> "disjoint").
// Most pairs leave through one of the two exits after a few Newton steps; only
// near-ties run until the duality gap is below BP_LP_GAP_TOL.
//
// One thread per set pair.  ALL CONTROL FLOW IS WARP-UNIFORM: the 32 lanes of a
// warp step through "one Newton iteration per trip" together, loops run to the
// warp-wide maximum trip count and finished lanes are predicated off.  (The
// first version let every lane run its own nested loops and measured 3.1
// active threads per warp in ncu, profiles/r01_baseline_mvie_pair_summary.txt.)
// On the host (tests/host_harness.cpp) the vote macros degenerate to the
// single-thread predicate, so the same code is the CPU-tested specification.
#pragma once
#include "bp_math.cuh"
#include "bp_mvie.cuh"   // bp_ldl_solve

#define BP_LP_DIAMETER 1.0e2     // bound (metres) on how far a feasible point can be from an iterate
#define BP_LP_GAP_TOL 1.0e-13
#define BP_LP_T_MULT 25.0
#define BP_LP_INNER_MAX 30
#define BP_LP_OUTER_MAX 14

#ifdef __CUDA_ARCH__
#define BP_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define BP_WARP_MAX_INT(v) __reduce_max_sync(0xffffffffu, (v))
#else
#define BP_WARP_ANY(p) (p)
#define BP_WARP_MAX_INT(v) (v)
#endif

// `active` = this lane has a pair to test (every lane of the warp must call).
// xout (optional): start point in, last iterate out.  Returns 1 = intersects.
template <class ROWS>
BP_HD int bp_pair_feasible(const ROWS& r1, int m1, const ROWS& r2, int m2, double tol, double* xout,
                           int* iters_out, bool active = true, double t0_scale = 0.0) {
  if (!active) { m1 = 0; m2 = 0; }
  const int m = m1 + m2;
  const int mw = BP_WARP_MAX_INT(m);             // warp-uniform row-loop trip count
#define BP_ROW(i, A0, A1, A2, C)                                   \
  double A0 = 0.0, A1 = 0.0, A2 = 0.0, C = 1.0;                    \
  if ((i) < m1) { A0 = r1.a((i), 0); A1 = r1.a((i), 1); A2 = r1.a((i), 2); C = r1.b((i)) - tol; } \
  else if ((i) < m) { A0 = r2.a((i) - m1, 0); A1 = r2.a((i) - m1, 1); A2 = r2.a((i) - m1, 2); C = r2.b((i) - m1) - tol; }

  double x[4] = {0.0, 0.0, 0.0, 0.0};
  if (xout && active) { x[0] = xout[0]; x[1] = xout[1]; x[2] = xout[2]; }
  bool done = !active;
  int result = 0;
  int mm = 0;                                    // rows with a != 0
  {
    double smax = -BP_INF;
    bool empty = false;
    for (int i = 0; i < mw; ++i) {
      BP_ROW(i, a0, a1, a2, c)
      if (i < m) {
        if (a0 != 0.0 || a1 != 0.0 || a2 != 0.0) {
          double v = a0 * x[0] + a1 * x[1] + a2 * x[2] - c;
          smax = v > smax ? v : smax;
          ++mm;
        } else if (c < 0.0) {
          empty = true;                          // 0 <= c violated
        }
      }
    }
    if (!done) {
      if (empty) { done = true; result = 0; }
      else if (mm == 0 || smax <= 0.0) { done = true; result = 1; }
      x[3] = smax + 1.0;
    }
  }
  // barrier parameter at entry: 1, or (rows / initial violation) * t0_scale so that the first
  // centering already resolves margins of the size of the initial violation
  double t = 1.0;
  if (t0_scale > 0.0 && !done) { t = t0_scale * mm / (x[3] - 1.0); if (!(t > 1.0)) t = 1.0; }
  int iters = 0, inner = 0, outer = 0;
  while (BP_WARP_ANY(!done)) {                    // one Newton iteration per trip
    if (!done) ++iters;
    double g[4] = {0.0, 0.0, 0.0, t};
    double H[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) H[k] = 0.0;
    double rn = 0.0;                              // sum r_i
    double minq = BP_INF;                         // min slack_i
    for (int i = 0; i < mw; ++i) {
      BP_ROW(i, a0, a1, a2, c)
      const bool nz = (i < m) && (a0 != 0.0 || a1 != 0.0 || a2 != 0.0);
      const double slack = c - (a0 * x[0] + a1 * x[1] + a2 * x[2]) + x[3];
      if (nz && !done) {
        const double r = 1.0 / slack;
        minq = slack < minq ? slack : minq;
        const double v[4] = {a0 * r, a1 * r, a2 * r, -r};
        rn += r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          g[j] += v[j];
#pragma unroll
          for (int k = 0; k <= j; ++k) H[j * (j + 1) / 2 + k] += v[j] * v[k];
        }
      }
    }
    if (!done) {
      // exits
      if (x[3] - minq <= 0.0) { done = true; result = 1; }
      else {
        const double irn = 1.0 / rn;
        const double rho0 = g[0] * irn, rho1 = g[1] * irn, rho2 = g[2] * irn;       // sum lam_i a_i
        const double lb = x[3] - mm * irn - sqrt(rho0 * rho0 + rho1 * rho1 + rho2 * rho2) * BP_LP_DIAMETER;
        if (lb > 0.0) { done = true; result = 0; }
      }
    }
    double dx[4] = {0.0, 0.0, 0.0, 0.0};
    double lam2 = 0.0;
    bool stage_end = false;
    if (!done) {
      if (!bp_ldl_solve<4>(H, g, dx)) { done = true; result = 0; }
      else {
        lam2 = -(g[0] * dx[0] + g[1] * dx[1] + g[2] * dx[2] + g[3] * dx[3]);
        if (!(lam2 > 0.0)) stage_end = true;
      }
    }
    // line search, in lock step: every trip evaluates one trial step for the lanes that still need one
    bool need = !done && !stage_end;
    bool accepted = false;
    double alpha = 1.0;
    int bt = 0;
    while (BP_WARP_ANY(need && !accepted && bt < 60)) {
      const bool mine = need && !accepted && bt < 60;
      bool ok = true;
      double prod = 1.0, logsum = 0.0;
      const bool want_armijo = lam2 >= 0.01;
      for (int i = 0; i < mw; ++i) {
        BP_ROW(i, a0, a1, a2, c)
        const bool nz = (i < m) && (a0 != 0.0 || a1 != 0.0 || a2 != 0.0);
        if (nz && mine) {
          const double slack = c - (a0 * x[0] + a1 * x[1] + a2 * x[2]) + x[3];
          const double dsl = -(a0 * dx[0] + a1 * dx[1] + a2 * dx[2]) + dx[3];
          if (!(slack + alpha * dsl > 0.0)) ok = false;
          if (want_armijo) {
            prod *= 1.0 + alpha * dsl / slack;
            if ((i & 7) == 7) { logsum += log(prod); prod = 1.0; }
          }
        }
      }
      if (mine) {
        if (ok) {
          if (!want_armijo) accepted = true;
          else {
            logsum += log(prod);
            const double dF = t * alpha * dx[3] - logsum;
            if (dF <= -0.25 * alpha * lam2) accepted = true;
          }
        }
        if (!accepted) { alpha *= 0.5; ++bt; }
      }
    }
    if (need) {
      if (accepted) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] += alpha * dx[k];
        ++inner;
        if (lam2 < 1e-4 || inner >= BP_LP_INNER_MAX) stage_end = true;
      } else {
        stage_end = true;                         // no representable progress
      }
    }
    if (!done && stage_end) {
      if (mm / t < BP_LP_GAP_TOL || outer + 1 >= BP_LP_OUTER_MAX) {
        done = true; result = 0;                  // |s*| below the resolvable gap: not strictly feasible
      } else {
        t *= BP_LP_T_MULT;
        inner = 0;
        ++outer;
      }
    }
  }
#undef BP_ROW
  if (xout && active) { xout[0] = x[0]; xout[1] = x[1]; xout[2] = x[2]; }
  if (iters_out && active) *iters_out = iters;
  return result;
}
