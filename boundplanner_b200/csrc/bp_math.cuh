// bp_math.cuh -- small fp64 geometry primitives shared by every kernel.
//
// All functions are thread-serial and __host__ __device__ so that the same
// code is unit-tested on the CPU (tests/host_harness.cpp) and run on sm_100a.
// Reference sites are cited per function (paths relative to the reference
// repository root).
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define BP_HD __host__ __device__ __forceinline__
#else
#define BP_HD inline
#endif

#define BP_INF (1.0e300 * 1.0e300)

// Status codes written per work item (see include/bpgeo.h)
enum {
  BP_OK = 0,
  BP_ELLIPSE_VIOLATION = 1,   // ConvexSetFinder.py:433-438 RuntimeError
  BP_ROW_OVERFLOW = 2,        // more rows than BP_MAX_ROWS
  BP_MVIE_NO_INTERIOR = 3,    // centre / hint not strictly inside the polytope
  BP_MVIE_NOT_CONVERGED = 4,
  BP_ROW_CAP = 5,             // more rows than the caller's row_cap (reference MVIE buffers: 20 rows, quirk Q5)
  BP_NOT_A_POLYTOPE = 6,      // bp_polytope_vertices: unbounded or empty (util_functions.py:76 raises ValueError)
};

// ---------------------------------------------------------------------------
// 3x3 helpers (row-major double[9]); symmetric matrices stored full.
// ---------------------------------------------------------------------------
BP_HD void bp_mat3_vec(const double* M, const double* v, double* o) {
  o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}

// O = A * B^T
BP_HD void bp_mat3_mul_bt(const double* A, const double* B, double* O) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      O[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}

// O = A^T * A
BP_HD void bp_mat3_ata(const double* A, double* O) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      O[3 * i + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
}

BP_HD double bp_det3(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
         A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// general inverse through the adjugate; returns det
BP_HD double bp_inv3(const double* A, double* O) {
  double c00 = A[4] * A[8] - A[5] * A[7];
  double c01 = A[5] * A[6] - A[3] * A[8];
  double c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  double id = 1.0 / det;
  O[0] = c00 * id;
  O[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  O[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  O[3] = c01 * id;
  O[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  O[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  O[6] = c02 * id;
  O[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  O[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return det;
}

// smallest eigenvalue of a symmetric 3x3 (== min singular value for the PSD
// q_inv tested at ConvexSetFinder.py:232).  Closed form (Smith 1961).
BP_HD double bp_sym3_min_eig(const double* A) {
  double p1 = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
  double q = (A[0] + A[4] + A[8]) / 3.0;
  double d0 = A[0] - q, d1 = A[4] - q, d2 = A[8] - q;
  double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
  if (p2 <= 0.0) return q;
  double p = sqrt(p2 / 6.0);
  double ip = 1.0 / p;
  double B[9] = {d0 * ip, A[1] * ip, A[2] * ip, A[1] * ip, d1 * ip, A[5] * ip, A[2] * ip, A[5] * ip, d2 * ip};
  double r = 0.5 * bp_det3(B);
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  double phi = acos(r) / 3.0;
  return q + 2.0 * p * cos(phi + 2.0943951023931954923);
}

// L (tril packed [L00,L10,L11,L20,L21,L22]) -> E = L L^T, Q = E^-1 = L^-T L^-1,
// det(Q).  (ConvexSetFinder.py:533-535 builds q_new = L L^T; :227-229 inverts
// it through an SVD -- for an SPD matrix that is the plain inverse.)
BP_HD void bp_shape_from_L(const double* L, double* E, double* Q, double* detQ) {
  double l00 = L[0], l10 = L[1], l11 = L[2], l20 = L[3], l21 = L[4], l22 = L[5];
  E[0] = l00 * l00;
  E[1] = E[3] = l00 * l10;
  E[2] = E[6] = l00 * l20;
  E[4] = l10 * l10 + l11 * l11;
  E[5] = E[7] = l10 * l20 + l11 * l21;
  E[8] = l20 * l20 + l21 * l21 + l22 * l22;
  // W = L^-1 (lower)
  double w00 = 1.0 / l00, w11 = 1.0 / l11, w22 = 1.0 / l22;
  double w10 = -l10 * w00 * w11;
  double w21 = -l21 * w11 * w22;
  double w20 = -(l20 * w00 + l21 * w10) * w22;
  // Q = W^T W
  Q[0] = w00 * w00 + w10 * w10 + w20 * w20;
  Q[1] = Q[3] = w10 * w11 + w20 * w21;
  Q[2] = Q[6] = w20 * w22;
  Q[4] = w11 * w11 + w21 * w21;
  Q[5] = Q[7] = w21 * w22;
  Q[8] = w22 * w22;
  double dl = w00 * w11 * w22;
  *detQ = dl * dl;
}

// ---------------------------------------------------------------------------
// K1: closest point of an axis-aligned box to p in the metric M (SPD).
// Replaces the OSQP solve of ConvexSetFinder.py:465-489 (problem :10-49):
//   min ||x||^2  s.t. (A E) x <= b - A p0,   y = E x + p0
// which for A = [I; -I] (BoundPlanner.py:126-129) is, in y-space,
//   min (y-p)^T M (y-p)  s.t.  lb <= y <= ub,   M = E^-T E^-1.
// Exact: the minimiser is the unconstrained minimiser over the face / edge /
// vertex it lies on; enumerate the 27 bound patterns, keep the feasible ones,
// take the least objective (first wins on ties -- the point is unique).
// ---------------------------------------------------------------------------
struct BpMetric {
  double m00, m01, m02, m11, m12, m22;      // M (symmetric)
  // one free axis k: z_k = -(M_kf . z_f) * inv_kk
  double i00, i11, i22;
  // two free axes (i,j), fixed f: z_i = z_f * c_f[0], z_j = z_f * c_f[1]
  double c0[2];                             // f = 0, free (1,2)
  double c1[2];                             // f = 1, free (0,2)
  double c2[2];                             // f = 2, free (0,1)
};

BP_HD void bp_metric_init(const double* M, BpMetric* mt) {
  mt->m00 = M[0]; mt->m01 = M[1]; mt->m02 = M[2];
  mt->m11 = M[4]; mt->m12 = M[5]; mt->m22 = M[8];
  mt->i00 = 1.0 / M[0]; mt->i11 = 1.0 / M[4]; mt->i22 = 1.0 / M[8];
  double d;
  d = 1.0 / (mt->m11 * mt->m22 - mt->m12 * mt->m12);
  mt->c0[0] = -(mt->m22 * mt->m01 - mt->m12 * mt->m02) * d;
  mt->c0[1] = -(mt->m11 * mt->m02 - mt->m12 * mt->m01) * d;
  d = 1.0 / (mt->m00 * mt->m22 - mt->m02 * mt->m02);
  mt->c1[0] = -(mt->m22 * mt->m01 - mt->m02 * mt->m12) * d;
  mt->c1[1] = -(mt->m00 * mt->m12 - mt->m02 * mt->m01) * d;
  d = 1.0 / (mt->m00 * mt->m11 - mt->m01 * mt->m01);
  mt->c2[0] = -(mt->m11 * mt->m02 - mt->m01 * mt->m12) * d;
  mt->c2[1] = -(mt->m00 * mt->m12 - mt->m01 * mt->m02) * d;
}

BP_HD double bp_quad(const BpMetric& mt, double z0, double z1, double z2) {
  return z0 * (mt.m00 * z0 + 2.0 * (mt.m01 * z1 + mt.m02 * z2)) +
         z1 * (mt.m11 * z1 + 2.0 * mt.m12 * z2) + mt.m22 * z2 * z2;
}

// lo = lb - p, hi = ub - p.  Returns z = y - p (3) and a bitmask of which axes
// sit on a bound (bit k: on lb, bit 3+k: on ub).
BP_HD int bp_box_qp(const BpMetric& mt, const double* lo, const double* hi, double* z) {
  double best = BP_INF;
  int bestmask = 0;
  z[0] = z[1] = z[2] = 0.0;
  // interior: z = 0
  if (lo[0] <= 0.0 && 0.0 <= hi[0] && lo[1] <= 0.0 && 0.0 <= hi[1] && lo[2] <= 0.0 && 0.0 <= hi[2]) {
    return 0;   // p inside the box: objective 0 is the global minimum
  }
#define BP_TRY(Z0, Z1, Z2, MASK)                                  \
  {                                                               \
    double q_ = bp_quad(mt, (Z0), (Z1), (Z2));                    \
    if (q_ < best) { best = q_; z[0] = (Z0); z[1] = (Z1); z[2] = (Z2); bestmask = (MASK); } \
  }
  // faces: one axis fixed, two free
  for (int s = 0; s < 2; ++s) {
    double f;
    f = s ? hi[0] : lo[0];
    { double a = f * mt.c0[0], b = f * mt.c0[1];
      if (lo[1] <= a && a <= hi[1] && lo[2] <= b && b <= hi[2]) BP_TRY(f, a, b, s ? 8 : 1) }
    f = s ? hi[1] : lo[1];
    { double a = f * mt.c1[0], b = f * mt.c1[1];
      if (lo[0] <= a && a <= hi[0] && lo[2] <= b && b <= hi[2]) BP_TRY(a, f, b, s ? 16 : 2) }
    f = s ? hi[2] : lo[2];
    { double a = f * mt.c2[0], b = f * mt.c2[1];
      if (lo[0] <= a && a <= hi[0] && lo[1] <= b && b <= hi[1]) BP_TRY(a, b, f, s ? 32 : 4) }
  }
  // edges: two axes fixed, one free
  for (int s = 0; s < 4; ++s) {
    int sa = s & 1, sb = s >> 1;
    { double f1 = sa ? hi[1] : lo[1], f2 = sb ? hi[2] : lo[2];           // free axis 0
      double a = -(mt.m01 * f1 + mt.m02 * f2) * mt.i00;
      if (lo[0] <= a && a <= hi[0]) BP_TRY(a, f1, f2, (sa ? 16 : 2) | (sb ? 32 : 4)) }
    { double f0 = sa ? hi[0] : lo[0], f2 = sb ? hi[2] : lo[2];           // free axis 1
      double a = -(mt.m01 * f0 + mt.m12 * f2) * mt.i11;
      if (lo[1] <= a && a <= hi[1]) BP_TRY(f0, a, f2, (sa ? 8 : 1) | (sb ? 32 : 4)) }
    { double f0 = sa ? hi[0] : lo[0], f1 = sb ? hi[1] : lo[1];           // free axis 2
      double a = -(mt.m02 * f0 + mt.m12 * f1) * mt.i22;
      if (lo[2] <= a && a <= hi[2]) BP_TRY(f0, f1, a, (sa ? 8 : 1) | (sb ? 16 : 2)) }
  }
  // vertices
  for (int s = 0; s < 8; ++s) {
    double f0 = (s & 1) ? hi[0] : lo[0], f1 = (s & 2) ? hi[1] : lo[1], f2 = (s & 4) ? hi[2] : lo[2];
    BP_TRY(f0, f1, f2, ((s & 1) ? 8 : 1) | ((s & 2) ? 16 : 2) | ((s & 4) ? 32 : 4))
  }
#undef BP_TRY
  return bestmask;
}

// y from z and the bound mask: coordinates on a bound are EXACTLY the bound.
BP_HD void bp_box_point(const double* p, const double* lb, const double* ub, const double* z, int mask,
                        double* y) {
  for (int k = 0; k < 3; ++k) {
    double v = p[k] + z[k];
    if (mask & (1 << k)) v = lb[k];
    if (mask & (8 << k)) v = ub[k];
    y[k] = v;
  }
}

// min over the 8 box vertices of (a.v - b): rounding is monotone, so the
// minimum of the rounded sums is the rounded sum of the per-axis minima
// (vertex test of ConvexSetFinder.py:449-451 / :353-355).
BP_HD double bp_box_min_halfspace(const double* a, double b, const double* lb, const double* ub) {
  double t0 = fmin(a[0] * lb[0], a[0] * ub[0]);
  double t1 = fmin(a[1] * lb[1], a[1] * ub[1]);
  double t2 = fmin(a[2] * lb[2], a[2] * ub[2]);
  return ((t0 + t1) + t2) - b;
}

// ---------------------------------------------------------------------------
// K2: closest points between the segment p0 + phi d (0<=phi<=1) and a box.
// Replaces the qpOASES solve of ConvexSetFinder.py:491-510 (problem :52-99).
// g(phi) = dist^2(p(phi), box) is convex piecewise quadratic with breakpoints
// where p(phi) crosses a slab plane; minimise each piece in closed form.
// Among equal minimisers the smallest phi is returned.
// Returns phi; x = clamp(p(phi), lb, ub).
// ---------------------------------------------------------------------------
BP_HD double bp_seg_box(const double* p0, const double* d, const double* lb, const double* ub, double* x,
                        double* dist2_out) {
  double br[8];
  int nb = 0;
  br[nb++] = 0.0;
  for (int k = 0; k < 3; ++k) {
    if (d[k] != 0.0) {
      double t1 = (lb[k] - p0[k]) / d[k];
      double t2 = (ub[k] - p0[k]) / d[k];
      if (t1 > 0.0 && t1 < 1.0) br[nb++] = t1;
      if (t2 > 0.0 && t2 < 1.0) br[nb++] = t2;
    }
  }
  br[nb++] = 1.0;
  // insertion sort (<= 8 entries)
  for (int i = 1; i < nb; ++i) {
    double v = br[i];
    int j = i - 1;
    while (j >= 0 && br[j] > v) { br[j + 1] = br[j]; --j; }
    br[j + 1] = v;
  }
  double best = BP_INF, bphi = 0.0;
  for (int i = 0; i + 1 < nb; ++i) {
    double a = br[i], b = br[i + 1];
    double mid = 0.5 * (a + b);
    // on (a,b) each axis is in one regime: below lb, inside, above ub
    double qa = 0.0, qb = 0.0;          // g(phi) = qa phi^2 + 2 qb phi + c  (c irrelevant for argmin)
    for (int k = 0; k < 3; ++k) {
      double pm = p0[k] + mid * d[k];
      double e;
      if (pm < lb[k]) e = p0[k] - lb[k];
      else if (pm > ub[k]) e = p0[k] - ub[k];
      else continue;
      qa += d[k] * d[k];
      qb += d[k] * e;
    }
    double phi = a;                      // flat piece: smallest phi
    if (qa > 0.0) {
      phi = -qb / qa;
      phi = phi < a ? a : (phi > b ? b : phi);
    }
    double g = 0.0;
    for (int k = 0; k < 3; ++k) {
      double pk = p0[k] + phi * d[k];
      double e = pk < lb[k] ? lb[k] - pk : (pk > ub[k] ? pk - ub[k] : 0.0);
      g += e * e;
    }
    if (i == 0 || g < best * (1.0 - 1e-12) - 1e-24) {
      best = g;
      bphi = phi;
    }
  }
  for (int k = 0; k < 3; ++k) {
    double pk = p0[k] + bphi * d[k];
    x[k] = pk < lb[k] ? lb[k] : (pk > ub[k] ? ub[k] : pk);
  }
  *dist2_out = best;
  return bphi;
}

// ---------------------------------------------------------------------------
// K1p: closest point of a GENERAL convex polytope {y : a_r . y <= b_r, r < R} (R <= BP_OBS_ROWS, the 15-row
// obstacle sets of BoundPlanner.add_obstacle_reps / normalize_set_size, util_functions.py:119-133) to p in
// the metric M.  Replaces the same OSQP solve as K1 (ConvexSetFinder.py:465-489) for obstacles that are not
// boxes.  Exact: the minimiser is the M-projection of p onto the affine hull of its active rows (at most 3 in
// R^3); every subset of 1..3 rows is a candidate, the feasible candidate of least objective wins (first on
// ties).  Candidates are numbered 0 (no active row: p itself), then the single rows, the pairs (i < j) and
// the triples (i < j < k) in lexicographic order; bp_polytope_candidate evaluates ONE of them so that the
// enumeration can be split over the lanes of a warp (the host harness walks them serially).
// ---------------------------------------------------------------------------
#define BP_OBS_ROWS 15

struct BpPolyMetric {
  double M[6];    // m00 m01 m02 m11 m12 m22
  double W[6];    // M^-1, same packing
};

BP_HD void bp_poly_metric_init(const double* M9, BpPolyMetric* pm) {
  double Wi[9];
  bp_inv3(M9, Wi);
  pm->M[0] = M9[0]; pm->M[1] = M9[1]; pm->M[2] = M9[2]; pm->M[3] = M9[4]; pm->M[4] = M9[5]; pm->M[5] = M9[8];
  pm->W[0] = Wi[0]; pm->W[1] = 0.5 * (Wi[1] + Wi[3]); pm->W[2] = 0.5 * (Wi[2] + Wi[6]);
  pm->W[3] = Wi[4]; pm->W[4] = 0.5 * (Wi[5] + Wi[7]); pm->W[5] = Wi[8];
}

BP_HD void bp_sym_mul(const double* S, const double* v, double* o) {
  o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
  o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
  o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}

// ROWS: a(r,k), b(r).  Candidate with active rows (i), (i,j) or (i,j,k) (unused = -1).  Returns true and
// y / obj (squared M-distance) when the candidate exists (independent rows) and is feasible.
template <class ROWS>
BP_HD bool bp_polytope_candidate(const BpPolyMetric& pm, const ROWS& rows, int R, const double* p, int i, int j, int k,
                                 double* y, double* obj) {
  if (i < 0) {
    y[0] = p[0]; y[1] = p[1]; y[2] = p[2];
  } else if (j < 0) {
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    double w[3];
    bp_sym_mul(pm.W, a, w);
    const double den = a[0] * w[0] + a[1] * w[1] + a[2] * w[2];
    if (!(den > 0.0)) return false;                       // zero (padded) row
    const double lam = (a[0] * p[0] + a[1] * p[1] + a[2] * p[2] - rows.b(i)) / den;
    y[0] = p[0] - lam * w[0]; y[1] = p[1] - lam * w[1]; y[2] = p[2] - lam * w[2];
  } else if (k < 0) {
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    const double c[3] = {rows.a(j, 0), rows.a(j, 1), rows.a(j, 2)};
    double wa[3], wc[3];
    bp_sym_mul(pm.W, a, wa);
    bp_sym_mul(pm.W, c, wc);
    const double g11 = a[0] * wa[0] + a[1] * wa[1] + a[2] * wa[2];
    const double g12 = a[0] * wc[0] + a[1] * wc[1] + a[2] * wc[2];
    const double g22 = c[0] * wc[0] + c[1] * wc[1] + c[2] * wc[2];
    const double det = g11 * g22 - g12 * g12;
    if (!(det > 1e-14 * g11 * g22)) return false;         // parallel rows (or a padded one)
    const double r1 = a[0] * p[0] + a[1] * p[1] + a[2] * p[2] - rows.b(i);
    const double r2 = c[0] * p[0] + c[1] * p[1] + c[2] * p[2] - rows.b(j);
    const double l1 = (g22 * r1 - g12 * r2) / det, l2 = (g11 * r2 - g12 * r1) / det;
    y[0] = p[0] - (l1 * wa[0] + l2 * wc[0]);
    y[1] = p[1] - (l1 * wa[1] + l2 * wc[1]);
    y[2] = p[2] - (l1 * wa[2] + l2 * wc[2]);
  } else {
    // a vertex: solve the three planes directly (no Gram matrix: it would square the condition number)
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    const double c[3] = {rows.a(j, 0), rows.a(j, 1), rows.a(j, 2)};
    const double d[3] = {rows.a(k, 0), rows.a(k, 1), rows.a(k, 2)};
    const double n0 = a[1] * c[2] - a[2] * c[1], n1 = a[2] * c[0] - a[0] * c[2], n2 = a[0] * c[1] - a[1] * c[0];
    const double det = n0 * d[0] + n1 * d[1] + n2 * d[2];
    const double na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], nc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double nd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    if (!(det * det > 1e-24 * na * nc * nd)) return false;
    const double e0 = c[1] * d[2] - c[2] * d[1], e1 = c[2] * d[0] - c[0] * d[2], e2 = c[0] * d[1] - c[1] * d[0];
    const double f0 = d[1] * a[2] - d[2] * a[1], f1 = d[2] * a[0] - d[0] * a[2], f2 = d[0] * a[1] - d[1] * a[0];
    const double ab = rows.b(i), cb = rows.b(j), db = rows.b(k), id = 1.0 / det;
    y[0] = (ab * e0 + cb * f0 + db * n0) * id;
    y[1] = (ab * e1 + cb * f1 + db * n1) * id;
    y[2] = (ab * e2 + cb * f2 + db * n2) * id;
  }
  // feasibility on every row
  for (int r = 0; r < R; ++r) {
    const double q0 = rows.a(r, 0), q1 = rows.a(r, 1), q2 = rows.a(r, 2), br = rows.b(r);
    const double viol = q0 * y[0] + q1 * y[1] + q2 * y[2] - br;
    const double mag = fabs(q0 * y[0]) + fabs(q1 * y[1]) + fabs(q2 * y[2]) + (fabs(br) > 1.0 ? fabs(br) : 1.0);
    if (viol > 1e-10 * mag) return false;
  }
  const double z0 = y[0] - p[0], z1 = y[1] - p[1], z2 = y[2] - p[2];
  *obj = z0 * (pm.M[0] * z0 + 2.0 * (pm.M[1] * z1 + pm.M[2] * z2)) + z1 * (pm.M[3] * z1 + 2.0 * pm.M[4] * z2) +
         pm.M[5] * z2 * z2;
  return true;
}

// Serial walk over all candidates (host harness / one thread).  Returns false when no candidate is feasible
// (empty polytope).
template <class ROWS>
BP_HD bool bp_polytope_qp(const BpPolyMetric& pm, const ROWS& rows, int R, const double* p, double* y) {
  double best = BP_INF, yc[3], oc;
  bool found = false;
  if (bp_polytope_candidate(pm, rows, R, p, -1, -1, -1, yc, &oc)) { y[0] = yc[0]; y[1] = yc[1]; y[2] = yc[2]; return true; }
  for (int i = 0; i < R; ++i)
    if (bp_polytope_candidate(pm, rows, R, p, i, -1, -1, yc, &oc) && oc < best) {
      best = oc; found = true; y[0] = yc[0]; y[1] = yc[1]; y[2] = yc[2];
    }
  for (int i = 0; i < R; ++i)
    for (int j = i + 1; j < R; ++j)
      if (bp_polytope_candidate(pm, rows, R, p, i, j, -1, yc, &oc) && oc < best) {
        best = oc; found = true; y[0] = yc[0]; y[1] = yc[1]; y[2] = yc[2];
      }
  for (int i = 0; i < R; ++i)
    for (int j = i + 1; j < R; ++j)
      for (int k = j + 1; k < R; ++k)
        if (bp_polytope_candidate(pm, rows, R, p, i, j, k, yc, &oc) && oc < best) {
          best = oc; found = true; y[0] = yc[0]; y[1] = yc[1]; y[2] = yc[2];
        }
  return found;
}

// ---------------------------------------------------------------------------
// K2p: closest points between the segment p0 + phi d (0 <= phi <= 1) and a GENERAL convex polytope whose offsets
// are shrunk by `shrink` (b - 0.001, ConvexSetFinder.py:496).  Replaces the same qpOASES solve as K2 (:491-510,
// problem :52-99) for obstacles that are not boxes.  Exact: at the optimum x is the projection of p(phi) onto the
// affine hull of its active rows (<= 3) and phi is 0, 1 or the minimiser of the distance between the segment's
// line and that hull; every (row subset, phi status) is a candidate, the feasible one of least distance wins and,
// among equal ones, the smallest phi (like bp_seg_box).
// pm: 0 -> phi = 0, 1 -> phi = 1, 2 -> phi free (kept only when 0 < phi < 1).
// ---------------------------------------------------------------------------
template <class ROWS>
BP_HD bool bp_seg_polytope_candidate(const ROWS& rows, int R, double shrink, const double* p0, const double* d, int i,
                                     int j, int k, int pm, double* x, double* phi_out, double* obj) {
  double phi = (double)pm;
  if (i < 0) {
    if (pm == 2) return false;
    x[0] = p0[0] + phi * d[0]; x[1] = p0[1] + phi * d[1]; x[2] = p0[2] + phi * d[2];
  } else if (k >= 0) {
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    const double c[3] = {rows.a(j, 0), rows.a(j, 1), rows.a(j, 2)};
    const double e[3] = {rows.a(k, 0), rows.a(k, 1), rows.a(k, 2)};
    const double n0 = a[1] * c[2] - a[2] * c[1], n1 = a[2] * c[0] - a[0] * c[2], n2 = a[0] * c[1] - a[1] * c[0];
    const double det = n0 * e[0] + n1 * e[1] + n2 * e[2];
    const double na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], nc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double ne = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    if (!(det * det > 1e-24 * na * nc * ne)) return false;
    const double e0 = c[1] * e[2] - c[2] * e[1], e1 = c[2] * e[0] - c[0] * e[2], e2 = c[0] * e[1] - c[1] * e[0];
    const double f0 = e[1] * a[2] - e[2] * a[1], f1 = e[2] * a[0] - e[0] * a[2], f2 = e[0] * a[1] - e[1] * a[0];
    const double ab = rows.b(i) - shrink, cb = rows.b(j) - shrink, eb = rows.b(k) - shrink, id = 1.0 / det;
    x[0] = (ab * e0 + cb * f0 + eb * n0) * id;
    x[1] = (ab * e1 + cb * f1 + eb * n1) * id;
    x[2] = (ab * e2 + cb * f2 + eb * n2) * id;
    if (pm == 2) {
      const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (!(dd > 0.0)) return false;
      phi = (d[0] * (x[0] - p0[0]) + d[1] * (x[1] - p0[1]) + d[2] * (x[2] - p0[2])) / dd;
      if (!(phi > 0.0 && phi < 1.0)) return false;
    }
  } else if (j < 0) {
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    const double g = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    if (!(g > 0.0)) return false;
    const double u = a[0] * p0[0] + a[1] * p0[1] + a[2] * p0[2] - (rows.b(i) - shrink);
    const double w = a[0] * d[0] + a[1] * d[1] + a[2] * d[2];
    if (pm == 2) {
      const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (!(w * w > 1e-24 * g * dd)) return false;          // segment parallel to the plane
      phi = -u / w;                                          // the segment's line crosses the plane
      if (!(phi > 0.0 && phi < 1.0)) return false;
    }
    const double q[3] = {p0[0] + phi * d[0], p0[1] + phi * d[1], p0[2] + phi * d[2]};
    const double lam = (u + phi * w) / g;
    x[0] = q[0] - lam * a[0]; x[1] = q[1] - lam * a[1]; x[2] = q[2] - lam * a[2];
  } else {
    const double a[3] = {rows.a(i, 0), rows.a(i, 1), rows.a(i, 2)};
    const double c[3] = {rows.a(j, 0), rows.a(j, 1), rows.a(j, 2)};
    const double g11 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], g12 = a[0] * c[0] + a[1] * c[1] + a[2] * c[2];
    const double g22 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double det = g11 * g22 - g12 * g12;
    if (!(det > 1e-14 * g11 * g22)) return false;
    const double u1 = a[0] * p0[0] + a[1] * p0[1] + a[2] * p0[2] - (rows.b(i) - shrink);
    const double u2 = c[0] * p0[0] + c[1] * p0[1] + c[2] * p0[2] - (rows.b(j) - shrink);
    const double w1 = a[0] * d[0] + a[1] * d[1] + a[2] * d[2], w2 = c[0] * d[0] + c[1] * d[1] + c[2] * d[2];
    if (pm == 2) {
      // minimise (u + phi w)^T G^-1 (u + phi w)
      const double gw1 = (g22 * w1 - g12 * w2) / det, gw2 = (g11 * w2 - g12 * w1) / det;   // G^-1 w
      const double den = w1 * gw1 + w2 * gw2;
      const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      if (!(den > 1e-24 * dd)) return false;                 // segment parallel to the edge's normal plane
      phi = -(u1 * gw1 + u2 * gw2) / den;
      if (!(phi > 0.0 && phi < 1.0)) return false;
    }
    const double r1 = u1 + phi * w1, r2 = u2 + phi * w2;
    const double l1 = (g22 * r1 - g12 * r2) / det, l2 = (g11 * r2 - g12 * r1) / det;
    x[0] = p0[0] + phi * d[0] - (l1 * a[0] + l2 * c[0]);
    x[1] = p0[1] + phi * d[1] - (l1 * a[1] + l2 * c[1]);
    x[2] = p0[2] + phi * d[2] - (l1 * a[2] + l2 * c[2]);
  }
  for (int r = 0; r < R; ++r) {
    const double q0 = rows.a(r, 0), q1 = rows.a(r, 1), q2 = rows.a(r, 2), br = rows.b(r) - shrink;
    const double viol = q0 * x[0] + q1 * x[1] + q2 * x[2] - br;
    const double mag = fabs(q0 * x[0]) + fabs(q1 * x[1]) + fabs(q2 * x[2]) + (fabs(br) > 1.0 ? fabs(br) : 1.0);
    if (viol > 1e-10 * mag) return false;
  }
  const double z0 = p0[0] + phi * d[0] - x[0], z1 = p0[1] + phi * d[1] - x[1], z2 = p0[2] + phi * d[2] - x[2];
  *obj = z0 * z0 + z1 * z1 + z2 * z2;
  *phi_out = phi;
  return true;
}

// candidate (obj, phi) beats the incumbent: strictly closer, or as close (to rounding) with a smaller phi
BP_HD bool bp_seg_better(double obj, double phi, double best, double best_phi) {
  if (obj < best * (1.0 - 1e-12) - 1e-24) return true;
  return obj <= best * (1.0 + 1e-12) + 1e-24 && phi < best_phi;
}

// Serial walk (host harness / one thread).  Returns false for an empty polytope.
template <class ROWS>
BP_HD bool bp_seg_polytope_qp(const ROWS& rows, int R, double shrink, const double* p0, const double* d, double* x,
                              double* phi_out, double* dist2_out) {
  double best = BP_INF, bphi = BP_INF, xc[3], pc, oc;
  bool found = false;
#define BP_SEGP_TRY(I, J, K)                                                                         \
  for (int pm = 0; pm < 3; ++pm)                                                                     \
    if (bp_seg_polytope_candidate(rows, R, shrink, p0, d, (I), (J), (K), pm, xc, &pc, &oc) &&        \
        (!found || bp_seg_better(oc, pc, best, bphi))) {                                             \
      best = oc; bphi = pc; found = true; x[0] = xc[0]; x[1] = xc[1]; x[2] = xc[2];                  \
    }
  BP_SEGP_TRY(-1, -1, -1)
  for (int i = 0; i < R; ++i) BP_SEGP_TRY(i, -1, -1)
  for (int i = 0; i < R; ++i)
    for (int j = i + 1; j < R; ++j) BP_SEGP_TRY(i, j, -1)
  for (int i = 0; i < R; ++i)
    for (int j = i + 1; j < R; ++j)
      for (int k = j + 1; k < R; ++k) BP_SEGP_TRY(i, j, k)
#undef BP_SEGP_TRY
  *phi_out = bphi;
  *dist2_out = best;
  return found;
}
