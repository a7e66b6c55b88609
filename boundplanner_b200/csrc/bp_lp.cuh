// bp_lp.cuh -- K6: do two convex sets {A1 x <= b1}, {A2 x <= b2} intersect
// after shrinking every face by tol?
//
// Replaces BoundPlanner.set_intersection (BoundPlanner.py:774-787): a HiGHS
// feasibility LP with c = 0 over the stacked rows, called from add_edges with
// tol = 0.01 (:796-798).  The answer is the sign of the margin
//     s* = min_x max_i (a_i . x - (b_i - tol)) / ||a_i||      (rows with a_i != 0)
// (intersects <=> s* <= 0).  It is bracketed from both sides by a phase-I
// log-barrier Newton method in (x, s) in R^4:
//   upper bound: s_ub = max_i (a_i.x - c_i)/n_i at the current iterate (exact
//                evaluation; s_ub <= 0 is a feasible point -> "intersects");
//   lower bound: weak duality with the barrier multipliers, padded by
//                ||sum lam_i a_i|| * BP_LP_DIAMETER for the dual residual
//                (lb > 0 -> "disjoint").
// Most pairs leave through one of the two exits after a few Newton steps; only
// near-ties run until the duality gap is below BP_LP_GAP_TOL.
// Thread-serial: one thread per set pair.
#pragma once
#include "bp_math.cuh"
#include "bp_mvie.cuh"   // bp_ldl_solve

#define BP_LP_DIAMETER 1.0e2     // bound (metres) on how far a feasible point can be from an iterate
#define BP_LP_GAP_TOL 1.0e-13
#define BP_LP_T_MULT 25.0

template <class ROWS>
BP_HD int bp_pair_feasible(const ROWS& r1, int m1, const ROWS& r2, int m2, double tol, double* xout,
                           int* iters_out) {
  const int m = m1 + m2;
#define BP_ROW(i, A0, A1, A2, C)                                   \
  double A0, A1, A2, C;                                            \
  if ((i) < m1) { A0 = r1.a((i), 0); A1 = r1.a((i), 1); A2 = r1.a((i), 2); C = r1.b((i)) - tol; } \
  else { A0 = r2.a((i) - m1, 0); A1 = r2.a((i) - m1, 1); A2 = r2.a((i) - m1, 2); C = r2.b((i) - m1) - tol; }

  double x[4] = {0.0, 0.0, 0.0, 0.0};
  if (xout) { x[0] = xout[0]; x[1] = xout[1]; x[2] = xout[2]; }   // caller-supplied start
  // s0 = s_ub(x0) + 1
  int mm = 0;
  {
    double smax = -BP_INF;
    for (int i = 0; i < m; ++i) {
      BP_ROW(i, a0, a1, a2, c)
      double n = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
      if (n > 0.0) {
        double v = (a0 * x[0] + a1 * x[1] + a2 * x[2] - c) / n;
        smax = v > smax ? v : smax;
        ++mm;
      } else if (c < 0.0) {
        if (iters_out) *iters_out = 0;
        return 0;                       // 0 <= c violated: empty
      }
    }
    if (mm == 0) { if (iters_out) *iters_out = 0; return 1; }
    if (smax <= 0.0) { if (iters_out) *iters_out = 0; return 1; }
    x[3] = smax + 1.0;
  }
  double t = 1.0;
  int iters = 0;
  int result = 0;
  for (int outer = 0; outer < 14; ++outer) {
    for (int inner = 0; inner < 30; ++inner) {
      ++iters;
      double g[4] = {0.0, 0.0, 0.0, t};
      double H[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) H[k] = 0.0;
      double rn = 0.0;                  // sum r_i n_i
      double minq = BP_INF;             // min slack_i / n_i
      for (int i = 0; i < m; ++i) {
        BP_ROW(i, a0, a1, a2, c)
        double n = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
        if (!(n > 0.0)) continue;
        double slack = c - (a0 * x[0] + a1 * x[1] + a2 * x[2]) + n * x[3];
        double r = 1.0 / slack;
        double q = slack / n;
        minq = q < minq ? q : minq;
        double v[4] = {a0 * r, a1 * r, a2 * r, -n * r};
        rn += n * r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          g[j] += v[j];
#pragma unroll
          for (int k = 0; k <= j; ++k) H[j * (j + 1) / 2 + k] += v[j] * v[k];
        }
      }
      // exits
      double s_ub = x[3] - minq;
      if (s_ub <= 0.0) { result = 1; goto done; }
      {
        double irn = 1.0 / rn;
        double rho0 = g[0] * irn, rho1 = g[1] * irn, rho2 = g[2] * irn;     // sum lam_i a_i
        double lb = x[3] - mm * irn - sqrt(rho0 * rho0 + rho1 * rho1 + rho2 * rho2) * BP_LP_DIAMETER;
        if (lb > 0.0) { result = 0; goto done; }
      }
      double dx[4];
      if (!bp_ldl_solve<4>(H, g, dx)) { result = 0; goto done; }
      double lam2 = -(g[0] * dx[0] + g[1] * dx[1] + g[2] * dx[2] + g[3] * dx[3]);
      if (!(lam2 > 0.0)) break;
      double alpha = 1.0;
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        bool ok = true;
        double prod = 1.0, logsum = 0.0;
        for (int i = 0; i < m; ++i) {
          BP_ROW(i, a0, a1, a2, c)
          double n = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
          if (!(n > 0.0)) continue;
          double slack = c - (a0 * x[0] + a1 * x[1] + a2 * x[2]) + n * x[3];
          double dsl = -(a0 * dx[0] + a1 * dx[1] + a2 * dx[2]) + n * dx[3];
          double rel = alpha * dsl / slack;
          if (!(rel > -1.0)) { ok = false; break; }
          prod *= 1.0 + rel;
          if ((i & 7) == 7) { logsum += log(prod); prod = 1.0; }
        }
        if (ok) {
          if (lam2 < 0.01) accepted = true;
          else {
            logsum += log(prod);
            double dF = t * alpha * dx[3] - logsum;
            if (dF <= -0.25 * alpha * lam2) accepted = true;
          }
          if (accepted) {
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] += alpha * dx[k];
            break;
          }
        }
        alpha *= 0.5;
      }
      if (!accepted) break;
      if (lam2 < 1e-4) break;
    }
    if (mm / t < BP_LP_GAP_TOL) break;
    t *= BP_LP_T_MULT;
  }
  result = 0;     // |s*| below the resolvable gap: not strictly feasible
done:
#undef BP_ROW
  if (xout) { xout[0] = x[0]; xout[1] = x[1]; xout[2] = x[2]; }
  if (iters_out) *iters_out = iters;
  return result;
}
