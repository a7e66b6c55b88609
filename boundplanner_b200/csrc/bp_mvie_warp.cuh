// bp_mvie_warp.cuh -- K4, warp-cooperative: ONE WARP solves one MVIE.
//
// Same algorithm, same constants and the same summation order over rows as the
// thread-serial specification in bp_mvie.cuh (which the host harness tests
// against the oracle); only the work distribution differs:
//   * lanes own rows (row = lane, lane + 32): slack / cone residual / barrier
//     gradient pieces are evaluated one row per lane;
//   * the Hessian, gradient and second-moment sums are "dots of two feature
//     columns over the rows": every lane owns up to two of the NH + NV + 6
//     outputs and walks the rows in shared memory in ascending order;
//   * the NV x NV LDL^T solve is done redundantly by every lane in registers,
//     so all control flow (Newton / line-search / path-following decisions) is
//     warp-uniform -- no divergence, no votes needed except feasibility.
// The thread-per-set version ran at 2.1 active threads per warp (ncu,
// profiles/r01_baseline_mvie_pair_summary.txt); this one keeps 32 lanes busy
// on the row work and removes the serialisation.
#pragma once
#include "bp_mvie.cuh"

// feature columns: rr[NV], tw*a[3], a[3], one; the row stride is kept odd so that
// the per-lane row writes spread over the shared-memory banks
#ifdef BPGEO_PROFILE
__device__ long long g_prof_mvie[8 * 65536];   // [cta][rows, dots, ldl, linesearch, predictor, newton iters, armijo evals, backtracks]
// (accumulated in registers, written once per solve: a global read-modify-write per lap would cost ~600 cycles)
#define BP_MPROF_INIT() long long mprof_t_ = 0; long long mprof_a_[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define BP_MPROF_MARK() mprof_t_ = clock64()
#define BP_MPROF_LAP(slot) { const long long now_ = clock64(); mprof_a_[slot] += now_ - mprof_t_; mprof_t_ = now_; }
#define BP_MPROF_COUNT(slot) { mprof_a_[slot] += 1; }
#define BP_MPROF_FLUSH() { if ((threadIdx.x & 31) == 0 && blockIdx.x < 65536) { for (int q_ = 0; q_ < 8; ++q_) atomicAdd((unsigned long long*)&g_prof_mvie[8 * blockIdx.x + q_], (unsigned long long)mprof_a_[q_]); } }
#else
#define BP_MPROF_INIT()
#define BP_MPROF_MARK()
#define BP_MPROF_LAP(slot)
#define BP_MPROF_COUNT(slot)
#define BP_MPROF_FLUSH()
#endif

#define BP_MVIE_W(NV) ((((NV) + 7) & 1) ? ((NV) + 7) : ((NV) + 8))
#define BP_MVIE_SCRATCH_DOUBLES (BP_MAX_ROWS * 17 + 64)

template <int NV>
__device__ __forceinline__ void bp_mvie_decode_output(int e, int* cx, int* cy) {
  constexpr int NH = NV * (NV + 1) / 2;
  if (e < NH) {
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= e) ++j;
    *cx = j;
    *cy = e - j * (j + 1) / 2;
  } else if (e < NH + NV) {
    *cx = e - NH;
    *cy = NV + 6;                                 // "one" column: plain sum
  } else if (e < NH + NV + 6) {
    const int w = e - NH - NV;                    // (p,q) in 00,01,02,11,12,22
    const int p = w < 3 ? 0 : (w < 5 ? 1 : 2);
    const int q = w < 3 ? w : (w < 5 ? w - 2 : 2);
    *cx = NV + p;                                 // tw * a_p
    *cy = NV + 3 + q;                             // a_q
  } else {
    *cx = -1;
    *cy = -1;
  }
}

__device__ __forceinline__ double bp_warp_min(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ double bp_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ double bp_warp_prod(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// A, b: rows of this set; scratch: BP_MVIE_SCRATCH_DOUBLES doubles of shared memory private to the warp.
// RPL = rows per lane: 1 while m <= 32 (every set the reference can build: its buffers hold 20 rows), else 2.
struct BpMvieOut {
  double L[6], d[3];
  int status, iters;
};

#define BP_MVIE_ABORTED 99      // internal: a speculative solve was told to stop (never leaves the fused kernel)
// ABORTABLE: abort_flag (shared memory) is polled once per Newton iteration; non-zero -> return BP_MVIE_ABORTED.
// A template flag, not a run-time test: the poll costs the plain solver 1.7 % when it is compiled in.
template <int NV, int RPL, bool ABORTABLE = false>
__device__ __forceinline__ BpMvieOut bp_mvie_warp_impl(const double* __restrict__ A, const double* __restrict__ b,
                                                       int m, double c00, double c01, double c02, double* scratch,
                                                       const volatile int* abort_flag) {
  const double c0[3] = {c00, c01, c02};
  BpMvieOut res;
  BP_MPROF_INIT();
  constexpr int NH = NV * (NV + 1) / 2;
  constexpr int W = BP_MVIE_W(NV);
  constexpr int NOUT = NH + NV + 6;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double* F = scratch;                          // [m][W]
  double* OUT = scratch + BP_MAX_ROWS * 17;     // [64]

  // rows owned by this lane
  double ra[RPL][3], rb[RPL];
  bool rv[RPL];
#pragma unroll
  for (int q = 0; q < RPL; ++q) {
    const int i = lane + 32 * q;
    rv[q] = i < m;
    ra[q][0] = ra[q][1] = ra[q][2] = 0.0;
    rb[q] = 1.0;
    if (rv[q]) {
      ra[q][0] = A[3 * i]; ra[q][1] = A[3 * i + 1]; ra[q][2] = A[3 * i + 2];
      rb[q] = b[i];
      F[i * W + NV + 3] = ra[q][0]; F[i * W + NV + 4] = ra[q][1]; F[i * W + NV + 5] = ra[q][2];
      F[i * W + NV + 6] = 1.0;
    }
  }
  // rows m .. ceil4(m)-1 stay all-zero so that the column-dot loops can run in unrolled groups of four
  const int m4 = (m + 3) & ~3;
  for (int e = m * W + lane; e < m4 * W; e += 32) F[e] = 0.0;
  int cx[2], cy[2];
  bp_mvie_decode_output<NV>(lane, &cx[0], &cy[0]);
  bp_mvie_decode_output<NV>(lane + 32, &cx[1], &cy[1]);

  // strictly feasible start: ball of half the inradius around c0
  double r = BP_INF;
#pragma unroll
  for (int q = 0; q < RPL; ++q) {
    if (rv[q]) {
      double s = rb[q] - (ra[q][0] * c0[0] + ra[q][1] * c0[1] + ra[q][2] * c0[2]);
      double nrm = sqrt(ra[q][0] * ra[q][0] + ra[q][1] * ra[q][1] + ra[q][2] * ra[q][2]);
      if (nrm > 0.0) r = fmin(r, s / nrm);
      else if (!(s > 0.0)) r = -1.0;
    }
  }
  r = bp_warp_min(r);
  if (!(r > 0.0) || !(r < BP_INF)) {
    res.status = BP_MVIE_NO_INTERIOR;
    res.iters = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) res.L[k] = 0.0;
    res.d[0] = c0[0]; res.d[1] = c0[1]; res.d[2] = c0[2];
    return res;
  }
  r *= 0.5;
  double x[NV];
  x[0] = r; x[1] = 0.0; x[2] = r; x[3] = 0.0; x[4] = 0.0; x[5] = r;
  if (NV == 9) { x[6] = c0[0]; x[7] = c0[1]; x[8] = c0[2]; }
  double cen[3] = {c0[0], c0[1], c0[2]};

  const double nu = 2.0 * m + 4.0;
  const double t_final = nu / BP_MVIE_GAP_TOL;
  double t = 1.0;
  int iters = 0;
  int status = BP_OK;
  double xc_prev[NV];
  double t_prev = 0.0;
  for (int outer = 0; outer < BP_MVIE_OUTER_MAX; ++outer) {
    const bool last = (t >= t_final);
    const double inner_tol = last ? 1e-13 : BP_MVIE_INNER_TOL;
    double lam2_prev = BP_INF;
    bool centred = false;
    for (int inner = 0; inner < BP_MVIE_INNER_MAX; ++inner) {
      if (ABORTABLE && *abort_flag) { status = BP_MVIE_ABORTED; goto done; }
      ++iters;
      BP_MPROF_COUNT(5);
      BP_MPROF_MARK();
      if (NV == 9) { cen[0] = x[6]; cen[1] = x[7]; cen[2] = x[8]; }
      // ---- per-row barrier pieces -> feature rows
      double rs[RPL], ru[RPL][3], rpsi[RPL], rtw[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const double a0 = ra[q][0], a1 = ra[q][1], a2 = ra[q][2];
        const double s = rb[q] - (a0 * cen[0] + a1 * cen[1] + a2 * cen[2]);
        const double u0 = x[0] * a0 + x[1] * a1 + x[3] * a2;
        const double u1 = x[2] * a1 + x[4] * a2;
        const double u2 = x[5] * a2;
        const double psi = s * s - (u0 * u0 + u1 * u1 + u2 * u2);
        rs[q] = s; ru[q][0] = u0; ru[q][1] = u1; ru[q][2] = u2; rpsi[q] = psi;
        rtw[q] = 0.0;
        if (rv[q]) {
          const double tw = 2.0 * bp_rcp(psi);
          rtw[q] = tw;
          double* f = F + (lane + 32 * q) * W;
          f[0] = -u0 * a0 * tw; f[1] = -u0 * a1 * tw; f[2] = -u1 * a1 * tw;
          f[3] = -u0 * a2 * tw; f[4] = -u1 * a2 * tw; f[5] = -u2 * a2 * tw;
          if (NV == 9) { f[6] = -s * a0 * tw; f[7] = -s * a1 * tw; f[8] = -s * a2 * tw; }
          f[NV] = tw * a0; f[NV + 1] = tw * a1; f[NV + 2] = tw * a2;
        }
      }
      __syncwarp();
      BP_MPROF_LAP(0);
      // ---- column dots over the rows (ascending row order, like the serial code).
      // Control flow is kept warp-uniform: with more than 32 outputs every lane
      // walks two column pairs (lanes without a second output walk a dummy).
      {
        const int ax0 = cx[0], ay0 = cy[0];
        double acc0 = 0.0, acc0b = 0.0;
        if (NOUT > 40) {
          const int ax1 = cx[1] >= 0 ? cx[1] : 0, ay1 = cx[1] >= 0 ? cy[1] : 0;
          double acc1 = 0.0, acc1b = 0.0;
          for (int i = 0; i < m4; i += 4) {
            const double* f = F + i * W;
            const double p0 = f[ax0], q0 = f[ay0], p1 = f[W + ax0], q1 = f[W + ay0];
            const double p2 = f[2 * W + ax0], q2 = f[2 * W + ay0], p3 = f[3 * W + ax0], q3 = f[3 * W + ay0];
            const double r0 = f[ax1], s0 = f[ay1], r1 = f[W + ax1], s1 = f[W + ay1];
            const double r2 = f[2 * W + ax1], s2 = f[2 * W + ay1], r3 = f[3 * W + ax1], s3 = f[3 * W + ay1];
            acc0 += p0 * q0; acc0b += p1 * q1; acc0 += p2 * q2; acc0b += p3 * q3;
            acc1 += r0 * s0; acc1b += r1 * s1; acc1 += r2 * s2; acc1b += r3 * s3;
          }
          if (cx[1] >= 0) OUT[lane + 32] = acc1 + acc1b;
        } else {
          for (int i = 0; i < m4; i += 4) {
            const double* f = F + i * W;
            const double p0 = f[ax0], q0 = f[ay0], p1 = f[W + ax0], q1 = f[W + ay0];
            const double p2 = f[2 * W + ax0], q2 = f[2 * W + ay0], p3 = f[3 * W + ax0], q3 = f[3 * W + ay0];
            acc0 += p0 * q0; acc0b += p1 * q1; acc0 += p2 * q2; acc0b += p3 * q3;
          }
        }
        OUT[lane] = acc0 + acc0b;
      }
      __syncwarp();
      double g[NV], H[NH];
      double i0, i2, i5;                 // 1 / (L00, L11, L22)
#pragma unroll
      for (int k = 0; k < NH; ++k) H[k] = OUT[k];
#pragma unroll
      for (int k = 0; k < NV; ++k) g[k] = -OUT[NH + k];
      {
        i0 = bp_rcp(x[0]); i2 = bp_rcp(x[2]); i5 = bp_rcp(x[5]);
        const double w00 = OUT[NH + NV], w01 = OUT[NH + NV + 1], w02 = OUT[NH + NV + 2];
        const double w11 = OUT[NH + NV + 3], w12 = OUT[NH + NV + 4];
        // NV == 6 has 33 outputs for 32 lanes: the last one, W22 = sum tw a2^2, follows from the gradient sum of
        // rr5 = -x5 tw a2^2 (g[5] = -sum rr5 at this point)
        const double w22 = (NV == 6) ? g[5] * i5 : OUT[NH + NV + 5];
        H[0] += w00; H[1] += w01; H[2] += w11; H[6] += w02; H[7] += w12; H[9] += w22;
        H[5] += w11; H[12] += w12; H[14] += w22; H[20] += w22;
        if (NV == 9) { H[27] -= w00; H[34] -= w01; H[35] -= w11; H[42] -= w02; H[43] -= w12; H[44] -= w22; }
        g[0] -= t * i0; g[2] -= 2.0 * t * i2; g[5] -= t * i5;
        H[0] += t * i0 * i0; H[5] += 2.0 * t * i2 * i2; H[20] += t * i5 * i5;
      }
      __syncwarp();                      // OUT is rewritten next iteration
      BP_MPROF_LAP(1);
      double dx[NV];
      double lam2 = 0.0;
      if (!bp_ldl_solve<NV>(H, g, dx, &lam2)) {
        status = (t > 1e8) ? BP_OK : BP_MVIE_NOT_CONVERGED;
        goto done;
      }
      if (!(lam2 > 0.0)) { centred = true; break; }
      BP_MPROF_LAP(2);
      // ---- line search: psi(x + alpha dx) = psi + alpha B1 + alpha^2 A2 per row
      double dcn[3] = {0.0, 0.0, 0.0};
      if (NV == 9) { dcn[0] = dx[6]; dcn[1] = dx[7]; dcn[2] = dx[8]; }
      double rds[RPL], rB1[RPL], rA2[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const double a0 = ra[q][0], a1 = ra[q][1], a2 = ra[q][2];
        const double ds = -(a0 * dcn[0] + a1 * dcn[1] + a2 * dcn[2]);
        const double e0 = dx[0] * a0 + dx[1] * a1 + dx[3] * a2;
        const double e1 = dx[2] * a1 + dx[4] * a2;
        const double e2 = dx[5] * a2;
        rds[q] = ds;
        rB1[q] = 2.0 * (rs[q] * ds - (ru[q][0] * e0 + ru[q][1] * e1 + ru[q][2] * e2));
        rA2[q] = ds * ds - (e0 * e0 + e1 * e1 + e2 * e2);
      }
      double alpha = 1.0;
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        BP_MPROF_COUNT(7);
        bool ok = (x[0] + alpha * dx[0] > 0.0) && (x[2] + alpha * dx[2] > 0.0) && (x[5] + alpha * dx[5] > 0.0);
        double prod = 1.0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          if (rv[q]) {
            const double rel = alpha * (rB1[q] + alpha * rA2[q]) * (0.5 * rtw[q]);     // / psi
            if (!(rs[q] + alpha * rds[q] > 0.0) || !(rel > -1.0)) ok = false;
            prod *= 1.0 + rel;
          }
        }
        ok = __all_sync(full, ok);
        if (ok) {
          if (lam2 < BP_MVIE_FULLSTEP_LAM2) accepted = true;
          else {
            BP_MPROF_COUNT(6);
            // the four logarithms of dF are evaluated side by side: lane 0..2 the objective terms, lane 3 the
            // barrier term (log of the product over the rows)
            const double ptot = bp_warp_prod(prod);
            const double arg = lane == 0 ? alpha * dx[0] * i0
                             : lane == 1 ? alpha * dx[2] * i2
                             : lane == 2 ? alpha * dx[5] * i5
                             : ptot - 1.0;
            const double lg = bp_log1p(arg);
            const double l0 = __shfl_sync(full, lg, 0), l1 = __shfl_sync(full, lg, 1);
            const double l2 = __shfl_sync(full, lg, 2), l3 = __shfl_sync(full, lg, 3);
            const double dF = -t * (l0 + 2.0 * l1 + l2) - l3;
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
      BP_MPROF_LAP(3);
      if (!accepted) { centred = lam2 < 1e-2; break; }
      if (lam2 < inner_tol) { centred = true; break; }
      if (lam2 < 1e-3 && lam2 > 0.1 * lam2_prev) { centred = true; break; }
      lam2_prev = lam2;
    }
    if (last) {
      if (!centred) status = BP_MVIE_NOT_CONVERGED;
      break;
    }
    // ---- next barrier parameter + secant predictor in tau = 1/t
    double t_next = t * BP_MVIE_T_MULT;
    if (t_next > t_final) t_next = t_final;
    if (t_prev > 0.0 && t >= BP_MVIE_PRED_FROM) {
      // (1/t_next - 1/t) / (1/t - 1/t_prev), with one division
      double w = ((t - t_next) * t_prev) / ((t_prev - t) * t_next);
      double xp[NV];
      bool ok = false;
      for (int tr = 0; tr < 4 && !ok; ++tr) {
#pragma unroll
        for (int k = 0; k < NV; ++k) xp[k] = x[k] + w * (x[k] - xc_prev[k]);
        ok = (xp[0] > 0.0) && (xp[2] > 0.0) && (xp[5] > 0.0);
        double cp[3] = {cen[0], cen[1], cen[2]};
        if (NV == 9) { cp[0] = xp[6]; cp[1] = xp[7]; cp[2] = xp[8]; }
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          if (rv[q]) {
            const double a0 = ra[q][0], a1 = ra[q][1], a2 = ra[q][2];
            const double s = rb[q] - (a0 * cp[0] + a1 * cp[1] + a2 * cp[2]);
            const double u0 = xp[0] * a0 + xp[1] * a1 + xp[3] * a2;
            const double u1 = xp[2] * a1 + xp[4] * a2;
            const double u2 = xp[5] * a2;
            if (!(s > 0.0) || !(s * s - (u0 * u0 + u1 * u1 + u2 * u2) > 0.0)) ok = false;
          }
        }
        ok = __all_sync(full, ok);
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
  BP_MPROF_FLUSH();
#pragma unroll
  for (int k = 0; k < 6; ++k) res.L[k] = x[k];
  if (NV == 9) { res.d[0] = x[6]; res.d[1] = x[7]; res.d[2] = x[8]; }
  else { res.d[0] = c0[0]; res.d[1] = c0[1]; res.d[2] = c0[2]; }
  res.iters = iters;
  res.status = status;
  return res;
}

// One copy of the solver per kernel and NV (the Newton loop is ~30 KB of SASS: inlining it at every call site
// of the fused kernel tripled that): all call sites share this function.
template <int NV>
__device__ __noinline__ BpMvieOut bp_mvie_warp_fn(const double* A, const double* b, int m, double c00, double c01,
                                                  double c02, double* scratch) {
  if (m <= 32) return bp_mvie_warp_impl<NV, 1>(A, b, m, c00, c01, c02, scratch, nullptr);
  return bp_mvie_warp_impl<NV, 2>(A, b, m, c00, c01, c02, scratch, nullptr);
}
// the copy that a speculative solve runs (fused kernel, large scenes in waves): stops when *abort_flag != 0
template <int NV>
__device__ __noinline__ BpMvieOut bp_mvie_warp_abortable_fn(const double* A, const double* b, int m, double c00,
                                                            double c01, double c02, double* scratch,
                                                            const volatile int* abort_flag) {
  if (m <= 32) return bp_mvie_warp_impl<NV, 1, true>(A, b, m, c00, c01, c02, scratch, abort_flag);
  return bp_mvie_warp_impl<NV, 2, true>(A, b, m, c00, c01, c02, scratch, abort_flag);
}

template <int NV>
__device__ __forceinline__ int bp_mvie_warp(const double* A, const double* b, int m, const double* c0, double* scratch,
                                            double* Lout, double* dout, int* iters_out,
                                            const volatile int* abort_flag = nullptr) {
  const BpMvieOut o = abort_flag ? bp_mvie_warp_abortable_fn<NV>(A, b, m, c0[0], c0[1], c0[2], scratch, abort_flag)
                                 : bp_mvie_warp_fn<NV>(A, b, m, c0[0], c0[1], c0[2], scratch);
#pragma unroll
  for (int k = 0; k < 6; ++k) Lout[k] = o.L[k];
  dout[0] = o.d[0]; dout[1] = o.d[1]; dout[2] = o.d[2];
  if (iters_out) *iters_out = o.iters;
  return o.status;
}
