// bp_mvie_fixed_r.cuh -- K4b: inscribed ellipsoid with FIXED rotation and centre.
//
// Replaces mvie_socp_fixed_r (ConvexSetFinder.py:564-588, factory :650-680,
// cones :745-766), only reached from find_set_around_line (:242-307):
//   vars x = [x0,x1,x2,t1,t2,t3]; cones  || diag(x) R^T a_i || <= b_i - a_i . p_mid,
//   x0 >= a_lb (:667-669),  t1^2 <= x0 x1, t2^2 <= x1 x2, t3^2 <= t1 t2, maximise t3,
// i.e. maximise (x0 x1^2 x2)^(1/4): the middle semi-axis is weighted twice (quirk Q1).
// In y_k = x_k^2 the cone rows become LINEAR,  sum_k g_ik^2 y_k <= s_i^2  with
// g_i = R^T a_i, s_i = b_i - a_i . p_mid > 0, and the objective is (half of)
//   log y0 + 2 log y1 + log y2,        y0 >= a_lb^2,
// a 3-variable problem with a self-concordant objective solved by log-barrier
// path following (Newton in y, exact line search feasibility, Armijo on F_t).
//
// One implementation for both sides: RED distributes the rows and reduces over
// them -- BpSerialRed (host harness / one thread) or BpWarpRed (lanes own rows).
#pragma once
#include "bp_mvie.cuh"

#define BP_MVIE_FR_GAP_TOL 1e-12
#define BP_MVIE_FR_T_MULT 20.0

struct BpSerialRed {
  BP_HD int first() const { return 0; }
  BP_HD int stride() const { return 1; }
  BP_HD double sum(double v) const { return v; }
  BP_HD double min(double v) const { return v; }
  BP_HD bool all(bool p) const { return p; }
};

#ifdef __CUDACC__
struct BpWarpRed {
  __device__ __forceinline__ int first() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int stride() const { return 32; }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  __device__ __forceinline__ double min(double v) const {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
  }
  __device__ __forceinline__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
};
#endif

// rows: a(i,k), b(i) accessors; p: fixed centre; R: row-major 3x3 whose COLUMNS are the axes; a_lb: lower
// bound of the first semi-axis.  Out: x[3] semi-axes (the reference's `eigs`).  Every participant of RED
// returns the same status and x.
template <class ROWS, class RED>
BP_HD int bp_mvie_fixed_r(const ROWS& rows, int m, const double* p, const double* R, double a_lb, const RED& red,
                          double* xout, int* iters_out) {
  const int i0 = red.first(), di = red.stride();
  // strictly feasible start (see the header): x0 between a_lb and its largest feasible value, the two
  // other semi-axes small enough for every row
  double x0max = BP_INF, rin = BP_INF;
  bool inside = true;
  for (int i = i0; i < m; i += di) {
    const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
    const double nrm = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const double s = rows.b(i) - (a0 * p[0] + a1 * p[1] + a2 * p[2]);
    if (!(s > 0.0)) inside = false;
    if (nrm > 0.0) {
      const double g0 = fabs(R[0] * a0 + R[3] * a1 + R[6] * a2);
      if (g0 > 0.0) x0max = fmin(x0max, s / g0);
      rin = fmin(rin, s / nrm);
    }
  }
  inside = red.all(inside);
  x0max = red.min(x0max);
  rin = red.min(rin);
  if (!inside || !(rin < BP_INF)) return BP_MVIE_NO_INTERIOR;
  const double lb = a_lb > 0.0 ? a_lb : 0.0;
  if (!(x0max > lb)) return BP_MVIE_NO_INTERIOR;          // the reference's SOCP is infeasible here
  double y[3];
  {
    double x0 = fmin(0.5 * (lb + x0max), lb + 0.5 * rin);
    if (!(x0 > lb)) x0 = 0.5 * (lb + x0max);
    y[0] = x0 * x0;
    double rem = BP_INF;                                   // largest common y1 = y2
    for (int i = i0; i < m; i += di) {
      const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
      const double s = rows.b(i) - (a0 * p[0] + a1 * p[1] + a2 * p[2]);
      const double g0 = R[0] * a0 + R[3] * a1 + R[6] * a2;
      const double g1 = R[1] * a0 + R[4] * a1 + R[7] * a2;
      const double g2 = R[2] * a0 + R[5] * a1 + R[8] * a2;
      const double c12 = g1 * g1 + g2 * g2;
      if (c12 > 0.0) rem = fmin(rem, (s * s - g0 * g0 * y[0]) / c12);
    }
    rem = red.min(rem);
    if (!(rem > 0.0)) return BP_MVIE_NO_INTERIOR;
    if (!(rem < BP_INF)) rem = 1.0;
    y[1] = y[2] = 0.5 * rem;
  }
  const double lb2 = lb * lb;
  const bool has_lb = lb > 0.0;
  const double nu = (double)m + 5.0;
  const double t_final = nu / BP_MVIE_FR_GAP_TOL;
  double t = 1.0;
  int iters = 0, status = BP_OK;
  for (int outer = 0; outer < 64; ++outer) {
    const bool last = t >= t_final;
    const double inner_tol = last ? 1e-13 : 1e-2;
    bool centred = false;
    double lam2_prev = BP_INF;
    for (int inner = 0; inner < 60; ++inner) {
      ++iters;
      // gradient / Hessian of  -sum log(e_i - c_i . y)
      double g0 = 0, g1 = 0, g2 = 0, h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
      for (int i = i0; i < m; i += di) {
        const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
        const double s = rows.b(i) - (a0 * p[0] + a1 * p[1] + a2 * p[2]);
        const double r0 = R[0] * a0 + R[3] * a1 + R[6] * a2;
        const double r1 = R[1] * a0 + R[4] * a1 + R[7] * a2;
        const double r2 = R[2] * a0 + R[5] * a1 + R[8] * a2;
        const double c0 = r0 * r0, c1 = r1 * r1, c2 = r2 * r2;
        const double res = s * s - (c0 * y[0] + c1 * y[1] + c2 * y[2]);
        const double ir = 1.0 / res;
        const double w0 = c0 * ir, w1 = c1 * ir, w2 = c2 * ir;
        g0 += w0; g1 += w1; g2 += w2;
        h00 += w0 * w0; h01 += w0 * w1; h02 += w0 * w2; h11 += w1 * w1; h12 += w1 * w2; h22 += w2 * w2;
      }
      double g[3] = {red.sum(g0), red.sum(g1), red.sum(g2)};
      double H[6] = {red.sum(h00), red.sum(h01), red.sum(h11), red.sum(h02), red.sum(h12), red.sum(h22)};
      const double iy0 = 1.0 / y[0], iy1 = 1.0 / y[1], iy2 = 1.0 / y[2];
      g[0] -= t * iy0; g[1] -= 2.0 * t * iy1; g[2] -= t * iy2;
      H[0] += t * iy0 * iy0; H[2] += 2.0 * t * iy1 * iy1; H[5] += t * iy2 * iy2;
      double ilb = 0.0;
      if (has_lb) {
        ilb = 1.0 / (y[0] - lb2);
        g[0] -= ilb;
        H[0] += ilb * ilb;
      }
      double dy[3];
      if (!bp_ldl_solve<3>(H, g, dy)) { status = (t > 1e8) ? BP_OK : BP_MVIE_NOT_CONVERGED; goto done; }
      const double lam2 = -(g[0] * dy[0] + g[1] * dy[1] + g[2] * dy[2]);
      if (!(lam2 > 0.0)) { centred = true; break; }
      double alpha = 1.0;
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        bool ok = (y[0] + alpha * dy[0] > lb2) && (y[1] + alpha * dy[1] > 0.0) && (y[2] + alpha * dy[2] > 0.0) &&
                  (y[0] + alpha * dy[0] > 0.0);
        double logsum = 0.0;
        for (int i = i0; i < m; i += di) {
          const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
          const double s = rows.b(i) - (a0 * p[0] + a1 * p[1] + a2 * p[2]);
          const double r0 = R[0] * a0 + R[3] * a1 + R[6] * a2;
          const double r1 = R[1] * a0 + R[4] * a1 + R[7] * a2;
          const double r2 = R[2] * a0 + R[5] * a1 + R[8] * a2;
          const double c0 = r0 * r0, c1 = r1 * r1, c2 = r2 * r2;
          const double res = s * s - (c0 * y[0] + c1 * y[1] + c2 * y[2]);
          const double rel = -alpha * (c0 * dy[0] + c1 * dy[1] + c2 * dy[2]) / res;   // res_new / res - 1
          if (!(rel > -1.0)) ok = false;
          else logsum += log1p(rel);
        }
        ok = red.all(ok);
        if (ok) {
          if (lam2 < 0.01) accepted = true;
          else {
            logsum = red.sum(logsum);
            double dF = -t * (log1p(alpha * dy[0] * iy0) + 2.0 * log1p(alpha * dy[1] * iy1) + log1p(alpha * dy[2] * iy2)) -
                        logsum;
            if (has_lb) dF -= log1p(alpha * dy[0] * ilb);
            if (dF <= -0.25 * alpha * lam2) accepted = true;
          }
          if (accepted) {
            y[0] += alpha * dy[0]; y[1] += alpha * dy[1]; y[2] += alpha * dy[2];
            break;
          }
        }
        alpha *= 0.5;
      }
      if (!accepted) { centred = lam2 < 1e-2; break; }
      if (lam2 < inner_tol) { centred = true; break; }
      if (lam2 < 1e-3 && lam2 > 0.1 * lam2_prev) { centred = true; break; }
      lam2_prev = lam2;
    }
    if (last) {
      if (!centred) status = BP_MVIE_NOT_CONVERGED;
      break;
    }
    t *= BP_MVIE_FR_T_MULT;
    if (t > t_final) t = t_final;
  }
done:
  xout[0] = sqrt(y[0]); xout[1] = sqrt(y[1]); xout[2] = sqrt(y[2]);
  if (iters_out) *iters_out = iters;
  return status;
}

// E = R diag(y) R^T,  Q = R diag(1/y) R^T  with y the SQUARED semi-axes (:586-587; also the initial
// guess of find_set_around_line, :259-261, whose y is (a_lb, 1e-4, 1e-4))
BP_HD void bp_shape_from_axes_sq(const double* R, const double* y, double* E, double* Q, double* detQ) {
  const double iy[3] = {1.0 / y[0], 1.0 / y[1], 1.0 / y[2]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double e = 0.0, q = 0.0;
      for (int k = 0; k < 3; ++k) {
        e += R[3 * i + k] * y[k] * R[3 * j + k];
        q += R[3 * i + k] * iy[k] * R[3 * j + k];
      }
      if (E) E[3 * i + j] = e;
      Q[3 * i + j] = q;
    }
  if (detQ) *detQ = bp_det3(Q);
}
BP_HD void bp_shape_from_axes(const double* R, const double* x, double* E, double* Q, double* detQ) {
  const double y[3] = {x[0] * x[0], x[1] * x[1], x[2] * x[2]};
  bp_shape_from_axes_sq(R, y, E, Q, detQ);
}

// frame of find_set_around_line (:245-258): columns dp_ref, b1, b2
BP_HD void bp_line_frame(const double* dp1, double* R, double* l_seg) {
  const double l = sqrt(dp1[0] * dp1[0] + dp1[1] * dp1[1] + dp1[2] * dp1[2]);
  const double d[3] = {dp1[0] / l, dp1[1] / l, dp1[2] / l};
  double bd[3] = {0.0, 0.0, 1.0};
  if (!(fabs(d[2]) < 0.99)) { bd[1] = 1.0; bd[2] = 0.0; }
  const double w = d[0] * bd[0] + d[1] * bd[1] + d[2] * bd[2];      // gram_schmidt (util_functions.py:108-116)
  double b1[3] = {bd[0] - w * d[0], bd[1] - w * d[1], bd[2] - w * d[2]};
  const double n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
  b1[0] /= n1; b1[1] /= n1; b1[2] /= n1;
  double b2[3] = {d[1] * b1[2] - d[2] * b1[1], d[2] * b1[0] - d[0] * b1[2], d[0] * b1[1] - d[1] * b1[0]};
  const double n2 = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
  b2[0] /= n2; b2[1] /= n2; b2[2] /= n2;
  for (int k = 0; k < 3; ++k) { R[3 * k] = d[k]; R[3 * k + 1] = b1[k]; R[3 * k + 2] = b2[k]; }
  *l_seg = l;
}
