// bp_mvie.cuh -- K4: maximum-volume inscribed ellipsoid of {x : A x <= b}.
//
// Replaces the CVXPY/Clarabel SOCP solves of the reference:
//   mvie_socp            ConvexSetFinder.py:512-537 (cones :699-720)   NV = 9
//   mvie_socp_fixed_mid  ConvexSetFinder.py:539-562 (cones :722-743)   NV = 6
// The reference maximises t3 under t1^2<=L00 L11, t2^2<=L11 L22, t3^2<=t1 t2,
// i.e. (L00 L11^2 L22)^(1/4): the middle Cholesky pivot is weighted twice
// (SURVEY quirk Q1).  The monotone-equivalent smooth problem solved here is
//   min  -(log L00 + 2 log L11 + log L22)
//   s.t. || L^T a_i || <= b_i - a_i . d            (d fixed when NV == 6)
// by a log-barrier path-following Newton method in the NV primal variables
// x = [L00,L10,L11,L20,L21,L22,(d0,d1,d2)] (tril order of :534).
// Thread-serial: one thread solves one set; rows are read through ROWS
// (shared memory on the GPU, a plain array on the host harness).
#pragma once
#include "bp_math.cuh"

#ifndef BP_MVIE_T_MULT
#define BP_MVIE_T_MULT 20.0
#endif
#ifndef BP_MVIE_INNER_TOL
#define BP_MVIE_INNER_TOL 1e-2
#endif
#ifndef BP_MVIE_T_MULT_LATE
#define BP_MVIE_T_MULT_LATE 20.0
#endif
#ifndef BP_MVIE_T_LATE_FROM
#define BP_MVIE_T_LATE_FROM 1e30
#endif
#ifndef BP_MVIE_PRED_FROM
#define BP_MVIE_PRED_FROM 300.0
#endif
#ifndef BP_MVIE_T0
#define BP_MVIE_T0 1.0
#endif
#ifndef BP_MVIE_WS_BETA
#define BP_MVIE_WS_BETA 0.9
#endif
// Newton decrement^2 below which the full step is taken without the Armijo test: for a self-concordant function
// the full step at lambda < 1 is feasible and changes F_t by at most -lambda^2 + (-lambda - log(1 - lambda)),
// which is negative up to lambda = 0.68 (lambda^2 = 0.46): the (log-heavy) test can only confirm the step there.
// (0.25 -> 0.45: same Newton iteration counts on C2, fewer Armijo evaluations, solve 0.092 -> 0.089 ms.)
#ifndef BP_MVIE_FULLSTEP_LAM2
#define BP_MVIE_FULLSTEP_LAM2 0.45
#endif
#define BP_MVIE_OUTER_MAX 48
#define BP_MVIE_INNER_MAX 40
#ifndef BP_MVIE_GAP_TOL
#define BP_MVIE_GAP_TOL 1e-11
#endif

// LDL^T solve of an NV x NV SPD system, H stored as packed lower triangle
// (index r*(r+1)/2 + c).  Returns false when a pivot is not positive.
// 1/x.  On the device: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16,
// not correctly rounded) -- it only steers Newton directions, and it takes the
// IEEE-division slow path off the critical path of every pivot.
BP_HD double bp_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}

// 1/x for the pivots of the LDL^T factorisation only: one Newton step (relative error ~1e-12).  The pivots scale
// the Newton DIRECTION; the point the iteration converges to is where the (accurately evaluated) gradient vanishes,
// whatever the accuracy of the direction -- and every pivot's reciprocal sits on the solver's dependent chain.
BP_HD double bp_rcp_pivot(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}

// log(1 + x) for the Armijo tests of the warp solvers (x > -1): about 40 instructions instead of the ~125 of the
// library's log1p, which sat on the dependent chain of every tested step (11 % of the MVIE's stall samples).
// u = 1 + x = 2^e m with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh z, z = (m - 1) / (m + 1) (|z| <= 0.172: ten
// terms), plus the rounding of u repaired to first order: (x - (u - 1)) / u.  <= 4 ulp (checked against log1p on
// 350 k arguments from -0.999 to 1000).  The tests it feeds compare a decrease with a quarter of the predicted one.
BP_HD double bp_log1p(double x) {
#ifdef __CUDA_ARCH__
  const double u = 1.0 + x;
  if (!(u > 1e-300) || !(u < 1e300)) return log1p(x);
  const double c = x - (u - 1.0);
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int e = ((hi >> 20) & 0x7ff) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, lo);
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double z = (m - 1.0) * bp_rcp(m + 1.0);
  const double z2 = z * z;
  double p = 1.0 / 21.0;
  p = fma(p, z2, 1.0 / 19.0); p = fma(p, z2, 1.0 / 17.0); p = fma(p, z2, 1.0 / 15.0); p = fma(p, z2, 1.0 / 13.0);
  p = fma(p, z2, 1.0 / 11.0); p = fma(p, z2, 1.0 / 9.0); p = fma(p, z2, 1.0 / 7.0); p = fma(p, z2, 1.0 / 5.0);
  p = fma(p, z2, 1.0 / 3.0);
  const double lm = 2.0 * z * fma(p, z2, 1.0);
  const double ed = (double)e;
  return fma(ed, 6.93147180369123816490e-01, lm + fma(ed, 1.90821492927058770002e-10, c * bp_rcp(u)));
#else
  return log1p(x);
#endif
}

#ifdef BP_MVIE_COUNT
static int bp_mvie_count_armijo = 0;
#endif
template <int NV>
BP_HD bool bp_ldl_solve(double* H, const double* g, double* dx, double* lam2_out = nullptr) {
  // right-looking (outer-product) LDL^T: after column j is scaled the trailing
  // updates are mutually independent, so the dependent chain per column is
  // reciprocal -> scale -> one update of the next pivot.
  double dinv[NV];
  bool spd = true;                        // (one test at the end: no data-dependent branch per pivot on the chain)
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double dj = H[j * (j + 1) / 2 + j];
    spd = spd && (dj > 0.0);
    dinv[j] = bp_rcp_pivot(dj);
    double col[NV], col2[NV];             // unscaled column j below the diagonal: L_ij d_j, and its squares
#pragma unroll
    for (int i = j + 1; i < NV; ++i) {
      col[i] = H[i * (i + 1) / 2 + j];
      col2[i] = col[i] * col[i];          // (independent of the reciprocal: the next pivot waits for one FMA only)
      H[i * (i + 1) / 2 + j] = col[i] * dinv[j];
    }
#pragma unroll
    for (int i = j + 1; i < NV; ++i) {
#pragma unroll
      for (int k = j + 1; k < i; ++k) H[i * (i + 1) / 2 + k] -= H[i * (i + 1) / 2 + j] * col[k];
      H[i * (i + 1) / 2 + i] -= col2[i] * dinv[j];
    }
  }
  if (!spd) return false;
  // forward: L y = -g
  double y[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double v = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= H[i * (i + 1) / 2 + k] * y[k];
    y[i] = v;
  }
  // Newton decrement^2 = g^T H^-1 g = sum y_i^2 / d_i: available before the back substitution, so that the
  // caller's tests on it are off the dependent chain
  if (lam2_out) {
    double l2 = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) l2 += y[i] * y[i] * dinv[i];
    *lam2_out = l2;
  }
  // backward: L^T dx = D^-1 y.  The terms are taken from the OLDEST dx to the newest, so that only the last
  // multiply-add of a row waits for the entry computed just before (dependent chain: one FMA per row).
#pragma unroll
  for (int i = NV - 1; i >= 0; --i) {
    double v = y[i] * dinv[i];
#pragma unroll
    for (int k = NV - 1; k > i; --k) v -= H[k * (k + 1) / 2 + i] * dx[k];
    dx[i] = v;
  }
  return true;
}

// ROWS: a(i,k), b(i) accessors.  c0: fixed centre (NV==6) or interior hint (NV==9).
// Out: L[6] (tril packed), d[3] centre.  Returns a BP_* status.
template <int NV, class ROWS>
BP_HD int bp_mvie_solve(const ROWS& rows, int m, const double* c0, double* Lout, double* dout, int* iters_out,
                        const double* L0 = nullptr, double t0 = 1.0, const double* x_start = nullptr,
                        double t_stop = 0.0) {
  // x_start (NV entries, strictly feasible) with t0: continue the path from that point (bp_mvie_pd.cuh);
  // t_stop > 0: return after the centring of the first stage with t >= t_stop.
  constexpr int NH = NV * (NV + 1) / 2;
  double x[NV];
  // strictly feasible start: ball of half the inradius around c0
  double r = BP_INF;
  for (int i = 0; i < m; ++i) {
    double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
    double s = rows.b(i) - (a0 * c0[0] + a1 * c0[1] + a2 * c0[2]);
    double nrm = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    if (nrm > 0.0) {
      double q = s / nrm;
      r = q < r ? q : r;
    } else if (!(s > 0.0)) {
      r = -1.0;
    }
  }
  if (!(r > 0.0) || !(r < BP_INF)) return BP_MVIE_NO_INTERIOR;
  r *= 0.5;
  x[0] = r; x[1] = 0.0; x[2] = r; x[3] = 0.0; x[4] = 0.0; x[5] = r;
  if (NV == 9) { x[6] = c0[0]; x[7] = c0[1]; x[8] = c0[2]; }
  double cen[3] = {c0[0], c0[1], c0[2]};
  double t = BP_MVIE_T0;
  if (L0) {
    // warm start: a previous shape L0 around the same centre, scaled to BP_MVIE_WS_BETA of the
    // largest factor that keeps it inside, entered at barrier parameter t0
    double beta = BP_INF;
    for (int i = 0; i < m; ++i) {
      double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
      double s = rows.b(i) - (a0 * c0[0] + a1 * c0[1] + a2 * c0[2]);
      double u0 = L0[0] * a0 + L0[1] * a1 + L0[3] * a2;
      double u1 = L0[2] * a1 + L0[4] * a2;
      double u2 = L0[5] * a2;
      double un = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
      if (un > 0.0) { double q = s / un; beta = q < beta ? q : beta; }
    }
    if (beta < BP_INF && beta > 0.0 && L0[0] > 0.0 && L0[2] > 0.0 && L0[5] > 0.0) {
      beta *= BP_MVIE_WS_BETA;
#pragma unroll
      for (int k = 0; k < 6; ++k) x[k] = beta * L0[k];
      t = t0;
    }
  }

  if (x_start) {
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = x_start[k];
    t = t0;
  }

  const double nu = 2.0 * m + 4.0;
  const double t_final = nu / BP_MVIE_GAP_TOL;
  if (t > t_final) t = t_final;
  int iters = 0;
  int status = BP_OK;
  double xc_prev[NV];           // previous centre x(t_prev), for the secant predictor
  double t_prev = 0.0;
  for (int outer = 0; outer < BP_MVIE_OUTER_MAX; ++outer) {
    const bool last = (t >= t_final);
    const double inner_tol = last ? 1e-13 : BP_MVIE_INNER_TOL;
    double lam2_prev = BP_INF;
    bool centred = false;
    for (int inner = 0; inner < BP_MVIE_INNER_MAX; ++inner) {
      ++iters;
      double g[NV], H[NH];
#pragma unroll
      for (int k = 0; k < NV; ++k) g[k] = 0.0;
#pragma unroll
      for (int k = 0; k < NH; ++k) H[k] = 0.0;
      if (NV == 9) { cen[0] = x[6]; cen[1] = x[7]; cen[2] = x[8]; }
      double w00 = 0, w01 = 0, w02 = 0, w11 = 0, w12 = 0, w22 = 0;   // sum (2/psi) a a^T
      for (int i = 0; i < m; ++i) {
        double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
        double s = rows.b(i) - (a0 * cen[0] + a1 * cen[1] + a2 * cen[2]);
        double u0 = x[0] * a0 + x[1] * a1 + x[3] * a2;
        double u1 = x[2] * a1 + x[4] * a2;
        double u2 = x[5] * a2;
        double psi = s * s - (u0 * u0 + u1 * u1 + u2 * u2);
        double ip = 1.0 / psi;
        double tw = 2.0 * ip;
        // rr = 2 v / psi with v = s grad(s) - J^T u  (half gradient of psi)
        double rr[NV];
        rr[0] = -u0 * a0 * tw; rr[1] = -u0 * a1 * tw; rr[2] = -u1 * a1 * tw;
        rr[3] = -u0 * a2 * tw; rr[4] = -u1 * a2 * tw; rr[5] = -u2 * a2 * tw;
        if (NV == 9) { rr[6] = -s * a0 * tw; rr[7] = -s * a1 * tw; rr[8] = -s * a2 * tw; }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          g[j] -= rr[j];
#pragma unroll
          for (int k = 0; k <= j; ++k) H[j * (j + 1) / 2 + k] += rr[j] * rr[k];
        }
        w00 += tw * a0 * a0; w01 += tw * a0 * a1; w02 += tw * a0 * a2;
        w11 += tw * a1 * a1; w12 += tw * a1 * a2; w22 += tw * a2 * a2;
      }
      // + (2/psi) J^T J : groups {L00,L10,L20}=idx{0,1,3}, {L11,L21}=idx{2,4}, {L22}=idx{5}
      H[0] += w00;              // (0,0)
      H[1] += w01;              // (1,0)
      H[2] += w11;              // (1,1)
      H[6] += w02;              // (3,0)
      H[7] += w12;              // (3,1)
      H[9] += w22;              // (3,3)
      H[5] += w11;              // (2,2)
      H[12] += w12;             // (4,2)
      H[14] += w22;             // (4,4)
      H[20] += w22;             // (5,5)
      if (NV == 9) {            // - (2/psi) grad(s) grad(s)^T on the centre block
        H[27] -= w00;           // (6,6)
        H[34] -= w01;           // (7,6)
        H[35] -= w11;           // (7,7)
        H[42] -= w02;           // (8,6)
        H[43] -= w12;           // (8,7)
        H[44] -= w22;           // (8,8)
      }
      // objective  -t (log L00 + 2 log L11 + log L22)
      {
        double i0 = 1.0 / x[0], i2 = 1.0 / x[2], i5 = 1.0 / x[5];
        g[0] -= t * i0; g[2] -= 2.0 * t * i2; g[5] -= t * i5;
        H[0] += t * i0 * i0; H[5] += 2.0 * t * i2 * i2; H[20] += t * i5 * i5;
      }
      double dx[NV];
      double lam2 = 0.0;                     // Newton decrement^2, from the factorisation (sum y_i^2 / d_i)
      if (!bp_ldl_solve<NV>(H, g, dx, &lam2)) {
        status = (t > 1e8) ? BP_OK : BP_MVIE_NOT_CONVERGED;
        goto done;
      }
      if (!(lam2 > 0.0)) { centred = true; break; }
      // backtracking: strict feasibility (+ Armijo on F_t while lambda >= 0.1; below
      // that the full Newton step of a self-concordant function is safe).
      // psi(x + alpha dx) = psi + alpha B1 + alpha^2 A2 is evaluated without
      // cancellation so that the Armijo test stays meaningful at t ~ 1e13.
      double dcn[3] = {0.0, 0.0, 0.0};
      if (NV == 9) { dcn[0] = dx[6]; dcn[1] = dx[7]; dcn[2] = dx[8]; }
      double alpha = 1.0;
#ifdef BP_MVIE_DAMPED
      if (lam2 >= BP_MVIE_FULLSTEP_LAM2) alpha = 1.0 / (1.0 + sqrt(lam2));
#endif
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        bool ok = (x[0] + alpha * dx[0] > 0.0) && (x[2] + alpha * dx[2] > 0.0) && (x[5] + alpha * dx[5] > 0.0);
        double logsum = 0.0, prod = 1.0;
        for (int i = 0; i < m && ok; ++i) {
          double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
          double s = rows.b(i) - (a0 * cen[0] + a1 * cen[1] + a2 * cen[2]);
          double ds = -(a0 * dcn[0] + a1 * dcn[1] + a2 * dcn[2]);
          double u0 = x[0] * a0 + x[1] * a1 + x[3] * a2;
          double u1 = x[2] * a1 + x[4] * a2;
          double u2 = x[5] * a2;
          double e0 = dx[0] * a0 + dx[1] * a1 + dx[3] * a2;
          double e1 = dx[2] * a1 + dx[4] * a2;
          double e2 = dx[5] * a2;
          double psi = s * s - (u0 * u0 + u1 * u1 + u2 * u2);
          double B1 = 2.0 * (s * ds - (u0 * e0 + u1 * e1 + u2 * e2));
          double A2 = ds * ds - (e0 * e0 + e1 * e1 + e2 * e2);
          double rel = alpha * (B1 + alpha * A2) / psi;      // psi_new / psi - 1
          if (!(s + alpha * ds > 0.0) || !(rel > -1.0)) { ok = false; break; }
          prod *= 1.0 + rel;
          if ((i & 7) == 7) { logsum += log(prod); prod = 1.0; }
        }
        if (ok) {
          if (lam2 < BP_MVIE_FULLSTEP_LAM2) { accepted = true; }
          else {
#ifdef BP_MVIE_COUNT
            ++bp_mvie_count_armijo;
#endif
            logsum += log(prod);
            double dF = -t * (log1p(alpha * dx[0] / x[0]) + 2.0 * log1p(alpha * dx[2] / x[2]) +
                              log1p(alpha * dx[5] / x[5])) - logsum;
            if (dF <= -0.25 * alpha * lam2) accepted = true;
          }
          if (accepted) {
#pragma unroll
            for (int k = 0; k < NV; ++k) x[k] += alpha * dx[k];
            break;
          }
        }
        alpha *= 0.5;
      }
#ifdef BP_TRACE
      printf("outer %d t %.1e inner %d lam2 %.3e alpha %.3g acc %d\n", outer, t, inner, lam2, alpha, (int)accepted);
#endif
      if (!accepted) { centred = lam2 < 1e-2; break; }   // no representable progress
      if (lam2 < inner_tol) { centred = true; break; }
      // quadratic convergence has hit the rounding floor
      if (lam2 < 1e-3 && lam2 > 0.1 * lam2_prev) { centred = true; break; }
      lam2_prev = lam2;
    }
    if (last) {
      if (!centred) status = BP_MVIE_NOT_CONVERGED;
      break;
    }
    if (t_stop > 0.0 && t >= t_stop) break;
    // next barrier parameter; near the optimum the central path is linear in
    // tau = 1/t, so start the next centering from the secant extrapolation of
    // the last two centres (kept only if strictly feasible).
    double t_next = t * (t >= BP_MVIE_T_LATE_FROM ? BP_MVIE_T_MULT_LATE : BP_MVIE_T_MULT);
    if (t_next > t_final) t_next = t_final;
    if (t_prev > 0.0 && t >= BP_MVIE_PRED_FROM) {
      double w = ((t - t_next) * t_prev) / ((t_prev - t) * t_next);
      double xp[NV];
      bool ok = false;
      for (int tr = 0; tr < 4 && !ok; ++tr) {
#pragma unroll
        for (int k = 0; k < NV; ++k) xp[k] = x[k] + w * (x[k] - xc_prev[k]);
        ok = (xp[0] > 0.0) && (xp[2] > 0.0) && (xp[5] > 0.0);
        double cp[3] = {cen[0], cen[1], cen[2]};
        if (NV == 9) { cp[0] = xp[6]; cp[1] = xp[7]; cp[2] = xp[8]; }
        for (int i = 0; i < m && ok; ++i) {
          double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
          double s = rows.b(i) - (a0 * cp[0] + a1 * cp[1] + a2 * cp[2]);
          double u0 = xp[0] * a0 + xp[1] * a1 + xp[3] * a2;
          double u1 = xp[2] * a1 + xp[4] * a2;
          double u2 = xp[5] * a2;
          ok = (s > 0.0) && (s * s - (u0 * u0 + u1 * u1 + u2 * u2) > 0.0);
        }
        w *= 0.5;
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) { xc_prev[k] = x[k]; if (ok) x[k] = xp[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < NV; ++k) xc_prev[k] = x[k];
    }
    t_prev = t;
    t = t_next;
  }
done:
#pragma unroll
  for (int k = 0; k < 6; ++k) Lout[k] = x[k];
  if (NV == 9) { dout[0] = x[6]; dout[1] = x[7]; dout[2] = x[8]; }
  else { dout[0] = c0[0]; dout[1] = c0[1]; dout[2] = c0[2]; }
  if (iters_out) *iters_out = iters;
  return status;
}
