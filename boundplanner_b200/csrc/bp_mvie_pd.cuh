// bp_mvie_pd.cuh -- SPECIFICATION (thread-serial, host-tested) of the next MVIE solver: barrier centre at t = 1,
// then a primal-dual predictor-corrector, then the last barrier stages.  Not yet used by the kernels (the warp
// version is next round's work); it exists so that the algorithm, its constants and its
// agreement with the oracle are pinned by tests/test_host_harness.py before the kernel is written.
//
// Same problem as bp_mvie.cuh (mvie_socp / mvie_socp_fixed_mid, ConvexSetFinder.py:512-562, quirk Q1):
//     min  -(log L00 + 2 log L11 + log L22)   s.t.  c_i(x) = (b_i - a_i.d ; L^T a_i)  in  Q^4  (second-order cone)
// written WITHOUT a nonlinear objective: the three pivots x_k get dual variables y_k with the FIXED complementarity
// target x_k y_k = w_k (w = 1, 2, 1: a weighted centre), the cones get z_i with c_i o z_i = mu e, mu -> 0, and
// stationarity is linear,   sum_i G_i^T z_i + sum_k e_k y_k = 0.   Only bilinear products remain, so Mehrotra's
// predictor-corrector with Nesterov-Todd scaling applies.  Per cone:  scaling point w = (c~ + J z~) / (2 gamma),
// eta = (det c / det z)^(1/4), W = eta (2 v v^T - J), W^-2 = eta^-2 (2 (Jw)(Jw)^T - J), lambda = W z = W^-1 c.
// Reduced system (NV x NV, same shape as the barrier Hessian):
//     [ sum_i G_i^T W_i^-2 G_i + diag(y_k / x_k) ] dx = r_d + sum_i G_i^T t_i + sum_k e_k comp_k / x_k
// affine step: t_i = -z_i, comp = w - x y  =>  right-hand side = e_k w_k / x_k;   corrector: t_i = -z_i + kappa_i,
// kappa_i = W_i^-1 ( lambda_i \ (sigma mu e - (W_i^-1 dc_i^a) o (W_i dz_i^a)) ),  comp = w - x y - dx^a dy^a.
// The primal-dual x converges like sqrt(gap) while the barrier's converges like the gap, so the solve ends with
// the barrier path from t = 2 m / gap (bp_mvie_solve with x_start), which first has to re-centre that point.
// Measured on the 256 C2 sets (tests/test_host_harness.py): 5 + 9 + 12 = 26 Newton iterations against 35, the
// same L to 1e-15, no fall-back needed; earlier / later hand-over points and extra centring steps were tried and
// are not better (24-55 iterations).  tools/research/mvie_primal_dual_prototype.py has the NumPy prototype.
#pragma once
#include "bp_mvie.cuh"

#ifndef BP_PD_GAP_TOL
#define BP_PD_GAP_TOL 1e-9        // hand-over to the barrier stages
#endif
#ifndef BP_PD_CENTER_STEPS
#define BP_PD_CENTER_STEPS 0      // pure centring steps (sigma = 1) after the gap target is met
#endif
#ifndef BP_PD_TC_SCALE
#define BP_PD_TC_SCALE 0.0        // > 0: hand over at t_c = scale / sqrt(gap) instead of 2 m / gap
#endif
#define BP_PD_MAX_ITERS 30
#define BP_PD_STEP_FRAC 0.99

struct BpQ4 { double v[4]; };

BP_HD double bp_q4_det(const double* a) {
  const double n = sqrt(a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
  return (a[0] - n) * (a[0] + n);
}
BP_HD double bp_q4_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }
// u o v
BP_HD void bp_q4_prod(const double* u, const double* v, double* o) {
  o[0] = bp_q4_dot(u, v);
  o[1] = u[0] * v[1] + v[0] * u[1];
  o[2] = u[0] * v[2] + v[0] * u[2];
  o[3] = u[0] * v[3] + v[0] * u[3];
}
// solve l o y = d
BP_HD void bp_q4_div(const double* l, const double* d, double* y) {
  const double y0 = (l[0] * d[0] - (l[1] * d[1] + l[2] * d[2] + l[3] * d[3])) / bp_q4_det(l);
  y[0] = y0;
  y[1] = (d[1] - y0 * l[1]) / l[0];
  y[2] = (d[2] - y0 * l[2]) / l[0];
  y[3] = (d[3] - y0 * l[3]) / l[0];
}
// largest alpha with c + alpha dc inside the cone (BP_INF: none)
BP_HD double bp_q4_max_step(const double* c, const double* dc) {
  const double a = dc[0] * dc[0] - (dc[1] * dc[1] + dc[2] * dc[2] + dc[3] * dc[3]);
  const double b = 2.0 * (c[0] * dc[0] - (c[1] * dc[1] + c[2] * dc[2] + c[3] * dc[3]));
  const double cc = bp_q4_det(c);
  double am = BP_INF;
  if (dc[0] < 0.0) am = -c[0] / dc[0];
  if (a == 0.0) {
    if (b < 0.0) { const double r = -cc / b; if (r < am) am = r; }
    return am;
  }
  const double disc = b * b - 4.0 * a * cc;
  if (disc >= 0.0) {
    const double sq = sqrt(disc);
    const double q = -0.5 * (b + (b >= 0.0 ? sq : -sq));
    const double r1 = q / a, r2 = (q != 0.0) ? cc / q : BP_INF;
    if (r1 > 0.0 && r1 < am) am = r1;
    if (r2 > 0.0 && r2 < am) am = r2;
  }
  return am;
}

// Nesterov-Todd scaling of one cone
struct BpNt {
  double v[4];       // W = eta (2 v v^T - J)
  double wq[4];      // J * scaling point: the rank-one direction of W^-2
  double eta, lam[4];
};
BP_HD void bp_nt_init(const double* c, const double* z, BpNt* s) {
  const double dc = bp_q4_det(c), dz = bp_q4_det(z);
  const double ic = 1.0 / sqrt(dc), iz = 1.0 / sqrt(dz);
  double ct[4], zt[4];
  for (int k = 0; k < 4; ++k) { ct[k] = c[k] * ic; zt[k] = z[k] * iz; }
  const double gam = sqrt(0.5 * (1.0 + bp_q4_dot(ct, zt)));
  double wb[4] = {(ct[0] + zt[0]) / (2.0 * gam), (ct[1] - zt[1]) / (2.0 * gam), (ct[2] - zt[2]) / (2.0 * gam),
                  (ct[3] - zt[3]) / (2.0 * gam)};
  s->eta = sqrt(sqrt(dc / dz));
  s->wq[0] = wb[0]; s->wq[1] = -wb[1]; s->wq[2] = -wb[2]; s->wq[3] = -wb[3];
  const double nv = 1.0 / sqrt(2.0 * (wb[0] + 1.0));
  s->v[0] = (wb[0] + 1.0) * nv; s->v[1] = wb[1] * nv; s->v[2] = wb[2] * nv; s->v[3] = wb[3] * nv;
  // lambda = W z
  const double vz = bp_q4_dot(s->v, z);
  s->lam[0] = s->eta * (2.0 * s->v[0] * vz - z[0]);
  for (int k = 1; k < 4; ++k) s->lam[k] = s->eta * (2.0 * s->v[k] * vz + z[k]);
}
BP_HD void bp_nt_W(const BpNt& s, const double* u, double* o) {          // W u
  const double vu = bp_q4_dot(s.v, u);
  o[0] = s.eta * (2.0 * s.v[0] * vu - u[0]);
  for (int k = 1; k < 4; ++k) o[k] = s.eta * (2.0 * s.v[k] * vu + u[k]);
}
BP_HD void bp_nt_Winv(const BpNt& s, const double* u, double* o) {       // W^-1 u = (2 Jv (Jv)^T - J) u / eta
  const double ju = s.v[0] * u[0] - (s.v[1] * u[1] + s.v[2] * u[2] + s.v[3] * u[3]);
  const double ie = 1.0 / s.eta;
  o[0] = ie * (2.0 * s.v[0] * ju - u[0]);
  for (int k = 1; k < 4; ++k) o[k] = ie * (-2.0 * s.v[k] * ju + u[k]);
}
BP_HD void bp_nt_Winv2(const BpNt& s, const double* u, double* o) {      // W^-2 u = (2 q q^T - J) u / eta^2
  const double qu = bp_q4_dot(s.wq, u);
  const double ie2 = 1.0 / (s.eta * s.eta);
  o[0] = ie2 * (2.0 * s.wq[0] * qu - u[0]);
  for (int k = 1; k < 4; ++k) o[k] = ie2 * (2.0 * s.wq[k] * qu + u[k]);
}

// factor once, solve for several right-hand sides (packed lower triangle, as bp_ldl_solve)
template <int NV>
BP_HD bool bp_ldl_factor(double* H, double* dinv) {
  for (int j = 0; j < NV; ++j) {
    const double dj = H[j * (j + 1) / 2 + j];
    if (!(dj > 0.0)) return false;
    dinv[j] = 1.0 / dj;
    double col[NV];
    for (int i = j + 1; i < NV; ++i) { col[i] = H[i * (i + 1) / 2 + j]; H[i * (i + 1) / 2 + j] = col[i] * dinv[j]; }
    for (int i = j + 1; i < NV; ++i)
      for (int k = j + 1; k <= i; ++k) H[i * (i + 1) / 2 + k] -= H[i * (i + 1) / 2 + j] * col[k];
  }
  return true;
}
template <int NV>
BP_HD void bp_ldl_apply(const double* H, const double* dinv, const double* rhs, double* dx) {
  double y[NV];
  for (int i = 0; i < NV; ++i) {
    double v = rhs[i];
    for (int k = 0; k < i; ++k) v -= H[i * (i + 1) / 2 + k] * y[k];
    y[i] = v;
  }
  for (int i = NV - 1; i >= 0; --i) {
    double v = y[i] * dinv[i];
    for (int k = i + 1; k < NV; ++k) v -= H[k * (k + 1) / 2 + i] * dx[k];
    dx[i] = v;
  }
}

// cone of row i at x, and G_i dx
template <int NV, class ROWS>
BP_HD void bp_pd_cone(const ROWS& rows, int i, const double* x, const double* c0, double* c) {
  const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
  const double* cen = NV == 9 ? x + 6 : c0;
  c[0] = rows.b(i) - (a0 * cen[0] + a1 * cen[1] + a2 * cen[2]);
  c[1] = x[0] * a0 + x[1] * a1 + x[3] * a2;
  c[2] = x[2] * a1 + x[4] * a2;
  c[3] = x[5] * a2;
}
template <int NV, class ROWS>
BP_HD void bp_pd_dcone(const ROWS& rows, int i, const double* dx, double* dc) {
  const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
  dc[0] = NV == 9 ? -(a0 * dx[6] + a1 * dx[7] + a2 * dx[8]) : 0.0;
  dc[1] = dx[0] * a0 + dx[1] * a1 + dx[3] * a2;
  dc[2] = dx[2] * a1 + dx[4] * a2;
  dc[3] = dx[5] * a2;
}
// out += G_i^T p
template <int NV, class ROWS>
BP_HD void bp_pd_Gt(const ROWS& rows, int i, const double* p, double* out) {
  const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
  out[0] += p[1] * a0; out[1] += p[1] * a1; out[2] += p[2] * a1; out[3] += p[1] * a2; out[4] += p[2] * a2;
  out[5] += p[3] * a2;
  if (NV == 9) { out[6] -= p[0] * a0; out[7] -= p[0] * a1; out[8] -= p[0] * a2; }
}

// Primal-dual phase: x (strictly feasible, centred at barrier parameter t0) -> x with gap <= BP_PD_GAP_TOL.
// z [m][4] scratch for the cone duals.  Returns BP_OK, or BP_MVIE_NOT_CONVERGED when the iteration stalls
// (the caller then continues with the barrier path from the last good point).
template <int NV, class ROWS>
BP_HD int bp_mvie_pd_phase(const ROWS& rows, int m, const double* c0, double* x, double t0, BpQ4* z, int* iters_out,
                           double* gap_out) {
  constexpr int NH = NV * (NV + 1) / 2;
  const int D[3] = {0, 2, 5};
  const double w[3] = {1.0, 2.0, 1.0};
  int mr = 0;                                   // rows with a non-zero normal (padded rows carry no cone)
  double y[3];
  for (int k = 0; k < 3; ++k) y[k] = w[k] / x[D[k]];
  for (int i = 0; i < m; ++i) {
    double c[4];
    bp_pd_cone<NV>(rows, i, x, c0, c);
    const bool real = rows.a(i, 0) != 0.0 || rows.a(i, 1) != 0.0 || rows.a(i, 2) != 0.0;
    mr += real ? 1 : 0;
    const double sc = real ? 2.0 / (t0 * bp_q4_det(c)) : 0.0;        // z = (2 / (t psi)) J c: the barrier's dual
    z[i].v[0] = sc * c[0]; z[i].v[1] = -sc * c[1]; z[i].v[2] = -sc * c[2]; z[i].v[3] = -sc * c[3];
  }
  if (mr == 0) return BP_MVIE_NO_INTERIOR;
  int iters = 0, status = BP_MVIE_NOT_CONVERGED;
  double gap = BP_INF;
  int centring_left = -1;                       // >= 0: the gap target is met, that many centring steps remain
  for (int it = 0; it < BP_PD_MAX_ITERS; ++it) {
    if (centring_left == 0) { status = BP_OK; break; }
    ++iters;
    double H[NH], rd[NV], mu = 0.0;
    for (int k = 0; k < NH; ++k) H[k] = 0.0;
    for (int k = 0; k < NV; ++k) rd[k] = 0.0;
    for (int i = 0; i < m; ++i) {
      if (z[i].v[0] == 0.0) continue;
      double c[4];
      bp_pd_cone<NV>(rows, i, x, c0, c);
      BpNt nt;
      bp_nt_init(c, z[i].v, &nt);
      mu += bp_q4_dot(c, z[i].v);
      bp_pd_Gt<NV>(rows, i, z[i].v, rd);
      // G^T W^-2 G = eta^-2 (2 (G^T q)(G^T q)^T - G^T J G)
      double gq[NV];
      for (int k = 0; k < NV; ++k) gq[k] = 0.0;
      bp_pd_Gt<NV>(rows, i, nt.wq, gq);
      const double om = 1.0 / (nt.eta * nt.eta);
      for (int r = 0; r < NV; ++r)
        for (int q = 0; q <= r; ++q) H[r * (r + 1) / 2 + q] += 2.0 * om * gq[r] * gq[q];
      const double a0 = rows.a(i, 0), a1 = rows.a(i, 1), a2 = rows.a(i, 2);
      // -G^T J G: + a a^T on the groups {L00,L10,L20} {L11,L21} {L22}, - a a^T on the centre block
      H[0] += om * a0 * a0; H[1] += om * a0 * a1; H[2] += om * a1 * a1; H[6] += om * a0 * a2; H[7] += om * a1 * a2;
      H[9] += om * a2 * a2; H[5] += om * a1 * a1; H[12] += om * a1 * a2; H[14] += om * a2 * a2; H[20] += om * a2 * a2;
      if (NV == 9) {
        H[27] -= om * a0 * a0; H[34] -= om * a0 * a1; H[35] -= om * a1 * a1; H[42] -= om * a0 * a2;
        H[43] -= om * a1 * a2; H[44] -= om * a2 * a2;
      }
    }
    mu /= mr;
    for (int k = 0; k < 3; ++k) { rd[D[k]] += y[k]; H[D[k] * (D[k] + 1) / 2 + D[k]] += y[k] / x[D[k]]; }
    double dinv[NV];
    if (!bp_ldl_factor<NV>(H, dinv)) break;
    // ---- affine direction: H dx = e_k w_k / x_k
    double rhs[NV], dxa[NV], dya[3];
    for (int k = 0; k < NV; ++k) rhs[k] = 0.0;
    for (int k = 0; k < 3; ++k) rhs[D[k]] = w[k] / x[D[k]];
    bp_ldl_apply<NV>(H, dinv, rhs, dxa);
    double aa = BP_INF;
    for (int k = 0; k < 3; ++k) {
      dya[k] = (w[k] - x[D[k]] * y[k]) / x[D[k]] - (y[k] / x[D[k]]) * dxa[D[k]];
      if (dxa[D[k]] < 0.0) { const double r = -x[D[k]] / dxa[D[k]]; if (r < aa) aa = r; }
      if (dya[k] < 0.0) { const double r = -y[k] / dya[k]; if (r < aa) aa = r; }
    }
    for (int i = 0; i < m; ++i) {
      if (z[i].v[0] == 0.0) continue;
      double c[4], dc[4], t4[4], dz[4];
      bp_pd_cone<NV>(rows, i, x, c0, c);
      bp_pd_dcone<NV>(rows, i, dxa, dc);
      BpNt nt;
      bp_nt_init(c, z[i].v, &nt);
      bp_nt_Winv2(nt, dc, t4);
      for (int k = 0; k < 4; ++k) dz[k] = -z[i].v[k] - t4[k];
      double r = bp_q4_max_step(c, dc);
      if (r < aa) aa = r;
      r = bp_q4_max_step(z[i].v, dz);
      if (r < aa) aa = r;
    }
    if (aa > 1.0) aa = 1.0;
    const double sigma = centring_left > 0 ? 1.0 : (1.0 - aa) * (1.0 - aa) * (1.0 - aa);
    // ---- corrector right-hand side
    for (int k = 0; k < NV; ++k) rhs[k] = 0.0;
    for (int i = 0; i < m; ++i) {
      if (z[i].v[0] == 0.0) continue;
      double c[4], dc[4], t4[4], dz[4], p[4], q[4], pq[4], rr[4], kap[4];
      bp_pd_cone<NV>(rows, i, x, c0, c);
      bp_pd_dcone<NV>(rows, i, dxa, dc);
      BpNt nt;
      bp_nt_init(c, z[i].v, &nt);
      bp_nt_Winv2(nt, dc, t4);
      for (int k = 0; k < 4; ++k) dz[k] = -z[i].v[k] - t4[k];
      bp_nt_Winv(nt, dc, p);
      bp_nt_W(nt, dz, q);
      bp_q4_prod(p, q, pq);
      rr[0] = sigma * mu - pq[0]; rr[1] = -pq[1]; rr[2] = -pq[2]; rr[3] = -pq[3];
      bp_q4_div(nt.lam, rr, t4);
      bp_nt_Winv(nt, t4, kap);
      bp_pd_Gt<NV>(rows, i, kap, rhs);
    }
    double comp[3];
    for (int k = 0; k < 3; ++k) {
      comp[k] = w[k] - x[D[k]] * y[k] - dxa[D[k]] * dya[k];
      rhs[D[k]] += (w[k] - dxa[D[k]] * dya[k]) / x[D[k]];           // r_d - G^T z cancels:  e (y + comp / x)
    }
    double dx[NV], dy[3];
    bp_ldl_apply<NV>(H, dinv, rhs, dx);
    double al = BP_INF;
    for (int k = 0; k < 3; ++k) {
      dy[k] = comp[k] / x[D[k]] - (y[k] / x[D[k]]) * dx[D[k]];
      if (dx[D[k]] < 0.0) { const double r = -x[D[k]] / dx[D[k]]; if (r < al) al = r; }
      if (dy[k] < 0.0) { const double r = -y[k] / dy[k]; if (r < al) al = r; }
    }
    // second pass: dz of the combined direction needs kappa again (recomputed: the serial form keeps no per-row state)
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < m; ++i) {
        if (z[i].v[0] == 0.0) continue;
        double c[4], dca[4], dc[4], t4[4], dza[4], p[4], q[4], pq[4], rr[4], kap[4], dz[4];
        bp_pd_cone<NV>(rows, i, x, c0, c);
        bp_pd_dcone<NV>(rows, i, dxa, dca);
        bp_pd_dcone<NV>(rows, i, dx, dc);
        BpNt nt;
        bp_nt_init(c, z[i].v, &nt);
        bp_nt_Winv2(nt, dca, t4);
        for (int k = 0; k < 4; ++k) dza[k] = -z[i].v[k] - t4[k];
        bp_nt_Winv(nt, dca, p);
        bp_nt_W(nt, dza, q);
        bp_q4_prod(p, q, pq);
        rr[0] = sigma * mu - pq[0]; rr[1] = -pq[1]; rr[2] = -pq[2]; rr[3] = -pq[3];
        bp_q4_div(nt.lam, rr, t4);
        bp_nt_Winv(nt, t4, kap);
        bp_nt_Winv2(nt, dc, t4);
        for (int k = 0; k < 4; ++k) dz[k] = -z[i].v[k] + kap[k] - t4[k];
        if (pass == 0) {
          double r = bp_q4_max_step(c, dc);
          if (r < al) al = r;
          r = bp_q4_max_step(z[i].v, dz);
          if (r < al) al = r;
        } else {
          for (int k = 0; k < 4; ++k) z[i].v[k] += al * dz[k];
        }
      }
      if (pass == 0) {
        al *= BP_PD_STEP_FRAC;
        if (al > 1.0) al = 1.0;
        if (!(al > 1e-8)) goto stalled;
      }
    }
    for (int k = 0; k < NV; ++k) x[k] += al * dx[k];
    for (int k = 0; k < 3; ++k) y[k] += al * dy[k];
    // gap and residuals at the new point
    {
      double g = 0.0, res[NV], yn = 0.0;
      for (int k = 0; k < NV; ++k) res[k] = 0.0;
      bool inside = x[0] > 0.0 && x[2] > 0.0 && x[5] > 0.0;
      for (int i = 0; i < m; ++i) {
        if (z[i].v[0] == 0.0) continue;
        double c[4];
        bp_pd_cone<NV>(rows, i, x, c0, c);
        if (!(c[0] > 0.0) || !(bp_q4_det(c) > 0.0) || !(bp_q4_det(z[i].v) > 0.0)) inside = false;
        g += bp_q4_dot(c, z[i].v);
        bp_pd_Gt<NV>(rows, i, z[i].v, res);
      }
      if (!inside || !(g == g)) goto stalled;
      double rn = 0.0, cw = 0.0;
      for (int k = 0; k < 3; ++k) {
        res[D[k]] += y[k];
        yn += y[k] * y[k];
        const double e = fabs(x[D[k]] * y[k] - w[k]);
        cw = e > cw ? e : cw;
      }
      for (int k = 0; k < NV; ++k) rn += res[k] * res[k];
      gap = g;
      if (centring_left > 0) --centring_left;
      else if (g < BP_PD_GAP_TOL && sqrt(rn) < 1e-9 * (sqrt(yn) > 1.0 ? sqrt(yn) : 1.0) && cw < 1e-9)
        centring_left = BP_PD_CENTER_STEPS;
    }
  }
stalled:
  if (iters_out) *iters_out = iters;
  if (gap_out) *gap_out = gap;
  return status;
}

// Hybrid solve with the interface of bp_mvie_solve.
template <int NV, class ROWS>
BP_HD int bp_mvie_pd_solve(const ROWS& rows, int m, const double* c0, double* Lout, double* dout, int* iters_out,
                           BpQ4* zscratch, int* phase_iters = nullptr) {
  double x[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int it_a = 0, it_b = 0, it_c = 0;
  // phase A: barrier centre at t = 1 (robust far from the path)
  int st = bp_mvie_solve<NV>(rows, m, c0, x, x + 6, &it_a, nullptr, 1.0, nullptr, 1.0);
  if (st != BP_OK) { if (iters_out) *iters_out = it_a; return st; }
  // phase B: primal-dual from the centred point
  double xb[9], gap = BP_INF;
  for (int k = 0; k < 9; ++k) xb[k] = x[k];
  const int stb = bp_mvie_pd_phase<NV>(rows, m, c0, xb, 1.0, zscratch, &it_b, &gap);
  double t_c = 1.0;
  if (stb == BP_OK) {
    for (int k = 0; k < 9; ++k) x[k] = xb[k];
    t_c = 2.0 * m / gap;
    if (BP_PD_TC_SCALE > 0.0 && BP_PD_TC_SCALE / sqrt(gap) < t_c) t_c = BP_PD_TC_SCALE / sqrt(gap);
#ifdef BP_PD_TC_MULT
    t_c *= BP_PD_TC_MULT;
#endif
  }                                              // else: stalled -> the barrier path from the t = 1 centre
#ifdef BP_PD_SKIP_PHASE_C
  if (stb == BP_OK) {
    for (int k = 0; k < 6; ++k) Lout[k] = x[k];
    for (int k = 0; k < 3; ++k) dout[k] = NV == 9 ? x[6 + k] : c0[k];
    if (iters_out) *iters_out = it_a + it_b;
    if (phase_iters) { phase_iters[0] = it_a; phase_iters[1] = it_b; phase_iters[2] = 0; }
    return BP_OK;
  }
#endif
  // phase C: the barrier path from t_c to the final gap (the primal-dual x converges like sqrt(gap))
  st = bp_mvie_solve<NV>(rows, m, c0, Lout, dout, &it_c, nullptr, stb == BP_OK ? t_c : BP_MVIE_T_MULT, x);
  if (iters_out) *iters_out = it_a + it_b + it_c;
  if (phase_iters) { phase_iters[0] = it_a; phase_iters[1] = it_b; phase_iters[2] = it_c; }
  return st;
}
