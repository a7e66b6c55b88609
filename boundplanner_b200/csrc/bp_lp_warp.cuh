// bp_lp_warp.cuh -- K6, warp-cooperative: ONE WARP decides one set pair.
//
// Same algorithm and constants as the thread-serial specification in bp_lp.cuh
// (host-tested against scipy/HiGHS); only the work distribution differs: lanes
// own rows of the stacked system (row = lane, lane+32, lane+64), the 4x4
// Newton system is assembled by column dots through shared memory and solved
// redundantly by every lane, so all control flow is warp-uniform.
// After the bounding-box filter only a few thousand pairs survive on the C2
// workload; one thread per pair left ~125 warps each walking 40 rows serially
// (2 ms, latency-bound); one warp per pair finishes in tens of microseconds.
#pragma once
#include "bp_lp.cuh"
#include "bp_mvie_warp.cuh"   // bp_warp_min / bp_warp_prod

#define BP_LP_SCRATCH_DOUBLES (96 * 5 + 16)

// ROWFN(i, a[3], c): row i of the system a.x <= c, i in [0, m), m <= 32 * BP_LP_SLOTS <= 96.
// BP_LP_SLOTS = rows per lane: the solver is instantiated for 1, 2 and 3 (a pair of the reference's sets has at
// most 40 rows, typically 26: one row per lane, a third of the row work of the 96-row form; same arithmetic per
// row, same reduction trees, so the results are bit-identical).
template <int BP_LP_SLOTS, class ROWFN>
__device__ __forceinline__ int bp_lp_feasible_warp_impl(const ROWFN& rowfn, int m, double* scratch, int* iters_out,
                                                        double* xout, const double* x0, double t0_scale,
                                                        const double* region) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double* F = scratch;              // [m][5]: v0 v1 v2 v3 (stride 5: conflict-free row writes)
  double* OUT = scratch + 96 * 5;   // [16]
  double ra[BP_LP_SLOTS][3], rc[BP_LP_SLOTS];
  bool rv[BP_LP_SLOTS];
  int mm = 0;
  bool empty = false;
#pragma unroll
  for (int q = 0; q < BP_LP_SLOTS; ++q) {
    const int i = lane + 32 * q;
    ra[q][0] = ra[q][1] = ra[q][2] = 0.0;
    rc[q] = 1.0;
    rv[q] = false;
    if (i < m) {
      rowfn(i, ra[q], rc[q]);
      rv[q] = (ra[q][0] != 0.0 || ra[q][1] != 0.0 || ra[q][2] != 0.0);
      if (!rv[q] && rc[q] < 0.0) empty = true;        // 0 <= c violated
    }
  }
  empty = __any_sync(full, empty);
  {
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < BP_LP_SLOTS; ++q) cnt += rv[q] ? 1 : 0;
    mm = __reduce_add_sync(full, cnt);
  }
  if (iters_out) *iters_out = 0;
  if (empty) return 0;
  if (mm == 0) return 1;
  const int m8 = (m + 7) & ~7;                        // F rows m .. m8-1 stay zero; F[.][4] = 1 on the real rows
  for (int i = lane; i < m8; i += 32) {
    F[i * 5 + 4] = i < m ? 1.0 : 0.0;
    if (i >= m) { F[i * 5] = 0.0; F[i * 5 + 1] = 0.0; F[i * 5 + 2] = 0.0; F[i * 5 + 3] = 0.0; }
  }
  double x[4] = {0.0, 0.0, 0.0, 0.0};
  if (x0) { x[0] = x0[0]; x[1] = x0[1]; x[2] = x0[2]; }
  double t = 1.0;
  {
    double smax = -BP_INF;
#pragma unroll
    for (int q = 0; q < BP_LP_SLOTS; ++q)
      if (rv[q]) smax = fmax(smax, ra[q][0] * x[0] + ra[q][1] * x[1] + ra[q][2] * x[2] - rc[q]);
    smax = -bp_warp_min(-smax);
    if (xout) { xout[0] = x[0]; xout[1] = x[1]; xout[2] = x[2]; }
    if (smax <= 0.0) return 1;
    x[3] = smax + 1.0;
    // entry barrier parameter (see bp_lp.cuh): resolves margins of the size of the initial violation
    if (t0_scale > 0.0) { t = t0_scale * mm / smax; if (!(t > 1.0)) t = 1.0; }
  }
  // output owned by this lane: e = lane & 15 -> H (10) or g (4); lanes >= 16 sum the odd rows
  const int e = lane & 15, half = lane >> 4;
  int cx = 0, cy = -1;                                  // cy = -1: plain sum of column cx
  if (e < 10) {
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= e) ++j;
    cx = j; cy = e - j * (j + 1) / 2;
  } else if (e < 14) {
    cx = e - 10;
  }
  int iters = 0, result = 0;
  for (int outer = 0; outer < BP_LP_OUTER_MAX; ++outer) {
    for (int inner = 0; inner < BP_LP_INNER_MAX; ++inner) {
      ++iters;
      double rslack[BP_LP_SLOTS], rinv[BP_LP_SLOTS];
      bool above = true;                                // every slack >= s: the iterate x is a common point
#pragma unroll
      for (int q = 0; q < BP_LP_SLOTS; ++q) {
        const double slack = rc[q] - (ra[q][0] * x[0] + ra[q][1] * x[1] + ra[q][2] * x[2]) + x[3];
        rslack[q] = slack;
        rinv[q] = 0.0;
        if (lane + 32 * q < m) {
          double* f = F + (lane + 32 * q) * 5;
          if (rv[q]) {
            const double r = bp_rcp(slack);             // steers the Newton direction only
            rinv[q] = r;
            above = above && (x[3] - slack <= 0.0);
            f[0] = ra[q][0] * r; f[1] = ra[q][1] * r; f[2] = ra[q][2] * r; f[3] = -r;
          } else {
            f[0] = 0.0; f[1] = 0.0; f[2] = 0.0; f[3] = 0.0;
          }
        }
      }
      // (a vote, not a five-stage shuffle reduction of the smallest slack: the warp is in-order, and everything
      // below waits for what is issued here)
      const bool common = __all_sync(full, above);
      __syncwarp();
      {
        // lanes 0-15 sum the even rows, lanes 16-31 the odd rows, four rows of a parity per trip (rows m .. m8-1 of
        // F are zero); the 14 outputs are H (10) and g (4), g as the dot with the all-ones "column" F[.][4]
        double acc = 0.0, accb = 0.0;
        const int cyy = cy >= 0 ? cy : 4;
        for (int i = half; i < m8; i += 8) {
          const double* f = F + i * 5;
          const double p0 = f[cx], q0 = f[cyy], p1 = f[10 + cx], q1 = f[10 + cyy];
          const double p2 = f[20 + cx], q2 = f[20 + cyy], p3 = f[30 + cx], q3 = f[30 + cyy];
          acc += p0 * q0; accb += p1 * q1; acc += p2 * q2; accb += p3 * q3;
        }
        acc += accb;
        acc += __shfl_xor_sync(full, acc, 16);
        if (lane < 14) OUT[lane] = acc;
      }
      __syncwarp();
      double H[10], g[4];
#pragma unroll
      for (int k = 0; k < 10; ++k) H[k] = OUT[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] = OUT[10 + k];
      __syncwarp();
      const double rn = -g[3];                          // sum r_i
      g[3] += t;
      // exits
      if (common) { result = 1; goto done; }            // s - min slack <= 0: max_i (a_i.x - c_i) <= 0, exactly
      {
        // weak-duality lower bound  lb = s - mm / rn - |sum lam_i a_i| * diam  with the barrier multipliers
        // lam_i = r_i / rn (r_i = 1 / slack_i, rn = sum r_i): multiplied through by rn > 0 and squared,
        //   lb > 0  <=>  A := s rn - mm > 0  and  A^2 > |sum r_i a_i|^2 diam^2
        // -- no division and no square root on the solver's dependent chain.  The dual residual is paid for over
        // the distance to the farthest point that could still be feasible: BP_LP_DIAMETER, or -- when the caller
        // knows a box that holds every feasible point (the overlap of the two sets' bounding boxes) -- the distance
        // from the iterate to the farthest corner of that box: if the sets intersected, a point x' of the box
        // would have max_i(a_i.x' - c_i) <= 0, but every x' of the box has max_i(.) >= sum lam_i (a_i.x' - c_i)
        // >= lb.  Typically 0.3 m instead of 100 m: "disjoint" is proven without centring the iterate to 1e-4 of
        // the margin first.  (1 + 1e-9) and the relative 1e-12 below keep the test on the safe side of rounding.
        double diam2 = BP_LP_DIAMETER * BP_LP_DIAMETER;
        if (region) {
          const double e0 = fmax(fabs(x[0] - region[0]), fabs(x[0] - region[3]));
          const double e1 = fmax(fabs(x[1] - region[1]), fabs(x[1] - region[4]));
          const double e2 = fmax(fabs(x[2] - region[2]), fabs(x[2] - region[5]));
          const double dd2 = (e0 * e0 + e1 * e1 + e2 * e2) * (1.0 + 1e-9) + 1e-24;
          if (dd2 < diam2) diam2 = dd2;
        }
        const double Aq = x[3] * rn - mm;
        const double g2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
        if (Aq > 0.0 && Aq * Aq * (1.0 - 1e-12) > g2 * diam2) { result = 0; goto done; }
      }
      double dx[4];
      if (!bp_ldl_solve<4>(H, g, dx)) { result = 0; goto done; }
      const double lam2 = -(g[0] * dx[0] + g[1] * dx[1] + g[2] * dx[2] + g[3] * dx[3]);
      if (!(lam2 > 0.0)) break;
      double rdsl[BP_LP_SLOTS];
#pragma unroll
      for (int q = 0; q < BP_LP_SLOTS; ++q)
        rdsl[q] = -(ra[q][0] * dx[0] + ra[q][1] * dx[1] + ra[q][2] * dx[2]) + dx[3];
      // full Newton step without the Armijo test while lambda^2 < 0.45: t s - sum log(slack_i) is self-concordant,
      // the full step then is feasible and cannot increase it (see BP_MVIE_FULLSTEP_LAM2)
      const bool want_armijo = lam2 >= 0.45;
      double alpha = 1.0;
      bool accepted = false;
      for (int bt = 0; bt < 60; ++bt) {
        bool ok = true;
        double prod = 1.0;
#pragma unroll
        for (int q = 0; q < BP_LP_SLOTS; ++q) {
          if (rv[q]) {
            if (!(rslack[q] + alpha * rdsl[q] > 0.0)) ok = false;
            if (want_armijo) prod *= 1.0 + alpha * rdsl[q] * rinv[q];
          }
        }
        ok = __all_sync(full, ok);
        if (ok) {
          if (!want_armijo) accepted = true;
          else {
            const double dF = t * alpha * dx[3] - bp_log1p(bp_warp_prod(prod) - 1.0);
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
  result = 0;       // |s*| below the resolvable gap: not strictly feasible
done:
  if (iters_out) *iters_out = iters;
  if (xout) { xout[0] = x[0]; xout[1] = x[1]; xout[2] = x[2]; }   // last iterate: strictly inside when result == 1
  return result;
}

template <class ROWFN>
__device__ int bp_lp_feasible_warp(const ROWFN& rowfn, int m, double* scratch, int* iters_out, double* xout = nullptr,
                                   const double* x0 = nullptr, double t0_scale = 0.0, const double* region = nullptr) {
  if (m <= 32) return bp_lp_feasible_warp_impl<1>(rowfn, m, scratch, iters_out, xout, x0, t0_scale, region);
  if (m <= 64) return bp_lp_feasible_warp_impl<2>(rowfn, m, scratch, iters_out, xout, x0, t0_scale, region);
  return bp_lp_feasible_warp_impl<3>(rowfn, m, scratch, iters_out, xout, x0, t0_scale, region);
}

// rows of set 1 then set 2, every offset shrunk by tol (BoundPlanner.set_intersection, :774-787)
struct BpPairRows {
  const double *A1, *b1, *A2, *b2;
  int m1;
  double tol;
  __device__ __forceinline__ void operator()(int i, double* a, double& c) const {
    const double* A = i < m1 ? A1 + 3 * i : A2 + 3 * (i - m1);
    a[0] = A[0]; a[1] = A[1]; a[2] = A[2];
    c = (i < m1 ? b1[i] : b2[i - m1]) - tol;
  }
};

__device__ __forceinline__ int bp_pair_feasible_warp(const double* __restrict__ A1, const double* __restrict__ b1,
                                                     int m1, const double* __restrict__ A2,
                                                     const double* __restrict__ b2, int m2, double tol, double* scratch,
                                                     int* iters_out, double* xout = nullptr,
                                                     const double* x0 = nullptr, double t0_scale = 0.0,
                                                     const double* region = nullptr) {
  // region (optional): lo[3] | hi[3] of a box that contains every point of both sets (finite entries only)
  BpPairRows rows{A1, b1, A2, b2, m1, tol};
  return bp_lp_feasible_warp(rows, m1 + m2, scratch, iters_out, xout, x0, t0_scale, region);
}
