"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's
``ConvexSetFinder`` (bound_planner/BoundPlanner/ConvexSetFinder.py:102-766).

Same public surface, same control flow, same thresholds and quirks; the
third-party solvers are replaced by exact ones:

* ``proj_solver`` (OSQP through casadi==3.6.7, :10-49, eps 1e-6) ->
  active-set enumeration of  min ||x||^2  s.t. (A E) x <= b - A p0   (exact).
* ``projl_solver`` (qpOASES, :52-99) -> KKT enumeration over the 3^4 bound
  patterns of (x, phi) for box obstacles (exact; singular Hessian, quirk Q9:
  among equal minimisers the smallest phi is returned).
* ``socp*_prob`` (CVXPY/Clarabel, :590-680) -> oracle/mvie.py.

PARITY UNPINNED against OSQP/qpOASES/Clarabel themselves (not installable
here; the reference has no tests or golden vectors) -- see oracle/__init__.py.
Vectorised over obstacles so that the C2-sized cases finish in seconds; the
per-obstacle loop structure of the reference (:468-486) is kept in
``compute_set_projs_loop`` for the timed CPU baseline.
"""
from __future__ import annotations

import itertools
import time

import numpy as np

from . import mvie as _mvie

_ACTIVE_SETS = {}


def _active_sets(nrows, kmax=3):
    key = (nrows, kmax)
    if key not in _ACTIVE_SETS:
        combos = []
        for k in range(1, kmax + 1):
            combos += [c for c in itertools.combinations(range(nrows), k)]
        _ACTIVE_SETS[key] = combos
    return _ACTIVE_SETS[key]


def min_norm_point_polytopes(G, h):
    """Solve  min ||x||^2  s.t.  G[j] x <= h[j]  for a batch of polytopes.

    G: (N, r, 3), h: (N, r).  Exact: the optimum is the min-norm point of the
    affine hull of its active rows (<= 3 of them in R^3); enumerate, keep the
    feasible candidates, take the one of least norm."""
    N, r, _ = G.shape
    best = np.full(N, np.inf)
    xbest = np.zeros((N, 3))
    scale = np.maximum(1.0, np.abs(h))
    # empty active set: x = 0
    feas = np.all(0.0 <= h + 1e-12 * scale, axis=1)
    best[feas] = 0.0
    for S in _active_sets(r):
        Gs = G[:, S, :]                      # (N,k,3)
        hs = h[:, S]                         # (N,k)
        nrm = np.prod(np.einsum("nkj,nkj->nk", Gs, Gs), axis=1)
        if len(S) == 3:
            # a vertex: solve G_S x = h_S directly -- the Gram matrix squares the condition number, which for a
            # strongly anisotropic metric (fixed-R ellipsoids, sigma ratios 1e3) drops below any safe threshold
            det = np.linalg.det(Gs)
            ok = det * det > 1e-24 * nrm
            if not np.any(ok):
                continue
            x = np.zeros((N, 3))
            x[ok] = np.linalg.solve(Gs[ok], hs[ok][..., None])[..., 0]
        else:
            gram = Gs @ Gs.transpose(0, 2, 1)    # (N,k,k)
            det = np.linalg.det(gram)
            ok = det > 1e-14 * nrm               # parallel rows (opposite box faces) give exactly 0
            if not np.any(ok):
                continue
            lam = np.zeros_like(hs)
            lam[ok] = np.linalg.solve(gram[ok], hs[ok][..., None])[..., 0]
            x = np.einsum("nkj,nk->nj", Gs, lam)
        viol = np.einsum("nrj,nj->nr", G, x) - h
        mag = np.einsum("nrj,nj->nr", np.abs(G), np.abs(x)) + scale
        feas = ok & np.all(viol <= 1e-10 * mag, axis=1)
        n2 = np.einsum("nj,nj->n", x, x)
        upd = feas & (n2 < best)
        best[upd] = n2[upd]
        xbest[upd] = x[upd]
    if not np.all(np.isfinite(best)):
        raise RuntimeError("closest-point QP infeasible (empty obstacle)")
    return xbest


def closest_points_segment_boxes(lb, ub, p0, p1):
    """min ||p0 + phi (p1-p0) - x||^2, lb <= x <= ub, 0 <= phi <= 1 for N boxes
    (ConvexSetFinder.py:52-99 with A=[I;-I]).  KKT enumeration over bound
    patterns.  Returns x (N,3), phi (N,)."""
    lb = np.asarray(lb, float)
    ub = np.asarray(ub, float)
    N = lb.shape[0]
    dvec = p1 - p0
    best = np.full(N, np.inf)
    cands = []
    for pat in itertools.product((0, 1, 2), repeat=3):      # 0 free, 1 at lb, 2 at ub
        fixed = np.array([s != 0 for s in pat])
        cfix = np.where(np.array(pat)[None, :] == 1, lb, ub)   # value if fixed
        for phis in (0, 1, 2):                                # 0 free, 1 phi=0, 2 phi=1
            if phis == 0:
                den = float(np.sum(dvec[fixed] ** 2))
                if den > 0.0:
                    num = np.sum((cfix[:, fixed] - p0[fixed]) * dvec[fixed], axis=1)
                    phi = num / den
                else:
                    phi = np.zeros(N)          # objective flat in phi: smallest phi
            else:
                phi = np.full(N, float(phis - 1))
            pphi = p0[None, :] + phi[:, None] * dvec[None, :]
            x = np.where(fixed[None, :], cfix, pphi)
            feas = (phi >= 0.0) & (phi <= 1.0) & np.all(x >= lb - 1e-15, axis=1) & np.all(x <= ub + 1e-15, axis=1)
            obj = np.sum((pphi - x) ** 2, axis=1)
            obj = np.where(feas, obj, np.inf)
            cands.append((obj, phi, x))
            best = np.minimum(best, obj)
    # among (numerically) equal minimisers take the smallest phi (quirk Q9)
    phi_out = np.full(N, np.inf)
    x_out = np.zeros((N, 3))
    thr = best * (1.0 + 1e-12) + 1e-24
    for obj, phi, x in cands:
        sel = (obj <= thr) & (phi < phi_out)
        phi_out[sel] = phi[sel]
        x_out[sel] = x[sel]
    return x_out, phi_out


def closest_points_segment_polytopes(A, b, p0, p1):
    """min ||p0 + phi (p1-p0) - x||^2, A[n] x <= b[n], 0 <= phi <= 1 for N polytopes (A: (N,r,3) with zero rows as
    padding, b already shrunk) -- the QP of ConvexSetFinder.py:52-99 for obstacles that are not boxes.
    Exact: x is the projection of p(phi) onto the affine hull of its active rows (<= 3), phi is 0, 1 or the
    minimiser of the line-to-hull distance; enumerate, keep the feasible candidates, take the least distance and
    among equal ones the smallest phi (quirk Q9).  Returns x (N,3), phi (N,)."""
    N, r, _ = A.shape
    d = p1 - p0
    dd = float(d @ d)
    best = np.full(N, np.inf)
    bphi = np.full(N, np.inf)
    xbest = np.zeros((N, 3))
    scale = np.maximum(1.0, np.abs(b))

    def consider(x, phi, ok):
        viol = np.einsum("nrj,nj->nr", A, x) - b
        mag = np.einsum("nrj,nj->nr", np.abs(A), np.abs(x)) + scale
        feas = ok & np.all(viol <= 1e-10 * mag, axis=1)
        q = p0[None] + phi[:, None] * d[None]
        obj = np.sum((q - x) ** 2, axis=1)
        better = feas & ((obj < best * (1 - 1e-12) - 1e-24) | ((obj <= best * (1 + 1e-12) + 1e-24) & (phi < bphi)))
        best[better] = obj[better]
        bphi[better] = phi[better]
        xbest[better] = x[better]

    ones = np.ones(N, bool)
    for fixed in (0.0, 1.0):                               # no active row: x = p(phi)
        consider(np.repeat((p0 + fixed * d)[None], N, axis=0), np.full(N, fixed), ones)
    for S in _active_sets(r):
        As = A[:, S, :]                                      # (N,k,3)
        bs = b[:, S]
        nrm = np.prod(np.einsum("nkj,nkj->nk", As, As), axis=1)
        if len(S) == 3:
            det = np.linalg.det(As)
            ok = det * det > 1e-24 * nrm
            if not np.any(ok):
                continue
            v = np.zeros((N, 3))
            v[ok] = np.linalg.solve(As[ok], bs[ok][..., None])[..., 0]
            for fixed in (0.0, 1.0):
                consider(v, np.full(N, fixed), ok)
            if dd > 0:
                phi = (v - p0[None]) @ d / dd
                consider(v, phi, ok & (phi > 0) & (phi < 1))
            continue
        gram = As @ As.transpose(0, 2, 1)
        det = np.linalg.det(gram)
        ok = det > 1e-14 * nrm
        if not np.any(ok):
            continue
        ginv = np.zeros_like(gram)
        ginv[ok] = np.linalg.inv(gram[ok])
        u = np.einsum("nkj,j->nk", As, p0) - bs               # A_S p0 - b_S
        w = np.einsum("nkj,j->nk", As, d)
        gw = np.einsum("nkl,nl->nk", ginv, w)
        den = np.einsum("nk,nk->n", w, gw)
        free_ok = ok & (den > 1e-24 * dd)
        phi_free = np.where(free_ok, -np.einsum("nk,nk->n", u, gw) / np.where(free_ok, den, 1.0), 0.0)
        for phi, okc in ((np.zeros(N), ok), (np.ones(N), ok), (phi_free, free_ok & (phi_free > 0) & (phi_free < 1))):
            rres = u + phi[:, None] * w
            lam = np.einsum("nkl,nl->nk", ginv, rres)
            x = p0[None] + phi[:, None] * d[None] - np.einsum("nkj,nk->nj", As, lam)
            consider(x, phi, okc)
    if not np.all(np.isfinite(best)):
        raise RuntimeError("segment QP infeasible (empty obstacle)")
    return xbest, bphi


class ConvexSetFinder:
    """Oracle twin of the reference class (same constructor / method names)."""

    def __init__(self, obs_sets, obs_points_sets, e_max, e_min, max_rows=20):
        self.rng = np.random.default_rng(0)                 # :105 (unused there too)
        self.ell_time = 0.0
        self.set_line_time = 0.0
        self.proj_time = 0.0
        self.obs_sets = list(obs_sets)
        self.obs_points_sets = list(obs_points_sets)
        self.e_max = e_max
        self.e_min = e_min
        self.max_iter = 5                                   # :133
        # reference MVIE buffers hold 20 rows and raise ValueError beyond (Q5)
        self.max_rows = max_rows
        self.verbose = False
        self.last_iters = 0

    # ---- helpers -------------------------------------------------------
    def _stack(self):
        A = np.stack([s[0] for s in self.obs_sets])          # (N,15,3)
        b = np.stack([s[1] for s in self.obs_sets])          # (N,15)
        # vertices (N,Vmax,3): boxes have 8 each; ragged polytope lists are padded by repeating the first vertex
        vs = [np.asarray(v, float).reshape(-1, 3) for v in self.obs_points_sets]
        vmax = max(v.shape[0] for v in vs)
        V = np.stack([np.vstack((v, np.repeat(v[:1], vmax - v.shape[0], axis=0))) for v in vs])
        return A, b, V

    @staticmethod
    def _nonzero_rows(A):
        """Rows to keep: all obstacles share the padded layout (real rows first, zero rows after); with ragged row
        counts the zero rows stay in (they are never active and always satisfied, b = 10)."""
        nz = np.linalg.norm(A, axis=2) > 0
        return int(nz.sum(axis=1).max())

    # ---- :377-421 ---------------------------------------------------------
    def init_halfspaces(self):
        a_set_init, b_set_init = [], []
        for i in range(3):
            a_set_init.append(np.eye(3)[i, :])
            b_set_init.append(float(self.e_max[i]))
            a_set_init.append(-np.eye(3)[i, :])
            b_set_init.append(-float(self.e_min[i]))
        return a_set_init, b_set_init

    def init_halfspaces_point(self, p, e_max=0.3):
        a_set_init, b_set_init = [], []
        for i in range(3):
            a_set_init.append(np.eye(3)[i, :])
            b_set_init.append(p[i] + e_max)
            a_set_init.append(-np.eye(3)[i, :])
            b_set_init.append(-p[i] + e_max)
        return a_set_init, b_set_init

    # ---- :465-489 ---------------------------------------------------------
    def compute_set_projs(self, obs_sets, p0, ellipse_mat):
        start = time.perf_counter()
        A = np.stack([s[0] for s in obs_sets])
        b = np.stack([s[1] for s in obs_sets])
        r = self._nonzero_rows(A)
        G = A[:, :r, :] @ ellipse_mat                        # rows of (A E)
        h = b[:, :r] - A[:, :r, :] @ p0
        x = min_norm_point_polytopes(G, h)
        obs_points = x @ ellipse_mat.T + p0                  # E x + p0 (:486)
        self.proj_time += time.perf_counter() - start
        return obs_points

    def compute_set_projs_loop(self, obs_sets, p0, ellipse_mat):
        """Same result, reference loop structure: one QP solve per obstacle (:468-486)."""
        obs_points = np.empty((len(obs_sets), 3))
        start = time.perf_counter()
        for i, (a_set, b_set) in enumerate(obs_sets):
            nz = np.linalg.norm(a_set, axis=1) > 0
            G = (a_set[nz] @ ellipse_mat)[None]
            h = (b_set[nz] - a_set[nz] @ p0)[None]
            x = min_norm_point_polytopes(G, h)[0]
            obs_points[i, :] = ellipse_mat @ x + p0
        self.proj_time += time.perf_counter() - start
        return obs_points

    # ---- :491-510 ---------------------------------------------------------
    def compute_set_projs_line(self, obs_sets, p0, p1):
        start = time.perf_counter()
        A = np.stack([s[0] for s in obs_sets])
        b = np.stack([s[1] for s in obs_sets])
        box = np.concatenate((np.eye(3), -np.eye(3)))
        if self._nonzero_rows(A) != 6 or not np.all(A[:, :6, :] == box[None]):
            # general polytopes: shrink every real row (the padded ones keep 0 x <= 10)
            r = self._nonzero_rows(A)
            real = np.linalg.norm(A[:, :r, :], axis=2) > 0
            x, phi = closest_points_segment_polytopes(A[:, :r, :], b[:, :r] - 0.001 * real, np.asarray(p0, float),
                                                      np.asarray(p1, float))
            self.proj_time += time.perf_counter() - start
            return x, phi
        ub = b[:, :3] - 0.001                                # b - 0.001 (:496)
        lb = -(b[:, 3:6] - 0.001)
        x, phi = closest_points_segment_boxes(lb, ub, np.asarray(p0, float), np.asarray(p1, float))
        self.proj_time += time.perf_counter() - start
        return x, phi

    # ---- :423-463 ---------------------------------------------------------
    def compute_polyhedron(self, q_inv, q_ellipse, p_seed, a_set_init, b_set_init):
        _, _, V = self._stack()
        a_set = list(a_set_init)
        b_set = list(b_set_init)
        obs_points = self.compute_set_projs(self.obs_sets, p_seed, q_inv)
        dists = np.linalg.norm(q_ellipse @ (obs_points - p_seed).T, axis=0)
        alive = np.ones(len(self.obs_sets), bool)
        self.last_picks = []
        while np.any(alive):
            idx = int(np.argmin(np.where(alive, dists, np.inf)))   # first index on ties (Q11)
            closest_point = obs_points[idx]
            if dists[idx] < 0.99:
                raise RuntimeError("Ellipse violates constraints")
            a_h = 2 * (q_ellipse @ q_ellipse.T) @ (closest_point - p_seed)
            b_h = a_h @ closest_point
            norm_a = np.linalg.norm(a_h)
            a_h = a_h / norm_a
            b_h = b_h / norm_a
            vmin = np.min(V @ a_h - b_h, axis=1)             # (N,)
            alive &= ~(vmin >= -1e-4)
            alive[idx] = False
            a_set.append(a_h)
            b_set.append(b_h)
            self.last_picks.append(idx)
        return a_set, b_set

    # ---- :512-588 ---------------------------------------------------------
    def _check_rows(self, a_set):
        if self.max_rows is not None and a_set.shape[0] > self.max_rows:
            # d2[:m] = b_set with m > 20 (:516)
            raise ValueError(
                f"could not broadcast input array from shape ({a_set.shape[0]},) into shape ({self.max_rows},)"
            )

    def mvie_socp(self, a_set, b_set, p_hint=None):
        self._check_rows(a_set)
        return _mvie.mvie_free(a_set, b_set, p_hint=p_hint)

    def mvie_socp_fixed_mid(self, a_set, b_set, p_mid):
        self._check_rows(a_set)
        return _mvie.mvie_fixed_mid(a_set, b_set, p_mid)

    def mvie_socp_fixed_r(self, a_set, b_set, p_mid, r_ellipse, a_lb):
        self._check_rows(a_set)
        return _mvie.mvie_fixed_r(a_set, b_set, p_mid, r_ellipse, a_lb)

    # ---- :190-240 ---------------------------------------------------------
    def find_set_around_point(self, p_seed, fixed_mid=False, optimize=True):
        p_seed = np.copy(p_seed)
        a = b = c = 1e-4
        q_inv = np.diag((a, b, c))
        q_ellipse = np.diag((1 / a, 1 / b, 1 / c))
        a_set_init, b_set_init = self.init_halfspaces()
        det_ellipse_old = 1
        det_ellipse = 100
        k = 0
        while np.abs(det_ellipse - det_ellipse_old) / det_ellipse_old > 0.01:
            k += 1
            if k > self.max_iter:
                break
            a_set, b_set = self.compute_polyhedron(q_inv, q_ellipse, p_seed, a_set_init, b_set_init)
            a_set_np = np.array(a_set)
            b_set_np = np.array(b_set)
            if not optimize:
                self.last_iters = k
                return a_set_np, b_set_np, q_ellipse, p_seed
            det_ellipse_old = np.copy(det_ellipse)
            start = time.perf_counter()
            if fixed_mid:
                q_inv, p_seed = self.mvie_socp_fixed_mid(a_set_np, b_set_np, p_seed)
            else:
                q_inv, p_seed = self.mvie_socp(a_set_np, b_set_np, p_hint=p_seed)
            self.ell_time += time.perf_counter() - start
            svd = np.linalg.svd(q_inv)
            q_ellipse = svd.Vh.T @ np.diag(1 / svd.S) @ svd.U.T
            det_ellipse = np.linalg.det(q_ellipse)
            if np.min(svd.S) < 1e-3:
                break
        self.last_iters = k
        if fixed_mid:
            q_inv, p_seed = self.mvie_socp(a_set_np, b_set_np, p_hint=p_seed)
            svd = np.linalg.svd(q_inv)
            q_ellipse = svd.Vh.T @ np.diag(1 / svd.S) @ svd.U.T
        return a_set_np, b_set_np, q_ellipse, p_seed

    # ---- :242-307 (not called by the reference planner on main, BoundPlanner.py:378-380) ----
    def find_set_around_line(self, p0, dp1, optimize=True):
        p0 = np.asarray(p0, float)
        dp1 = np.asarray(dp1, float)
        p1 = p0 + dp1
        r_ellipse, l_seg = line_frame(dp1)
        p_seed = (p0 + p1) / 2
        a_lb = l_seg**2 / 4
        b = c = 1e-4
        q_inv = r_ellipse @ np.diag((a_lb, b, c)) @ r_ellipse.T
        q_ellipse = r_ellipse @ np.diag((1 / a_lb, 1 / b, 1 / c)) @ r_ellipse.T
        a_set_init, b_set_init = self.init_halfspaces()
        det_ellipse_old = 1
        det_ellipse = 100
        k = 0
        a_set = b_set = None
        while np.abs(det_ellipse - det_ellipse_old) / det_ellipse_old > 0.01:
            k += 1
            if k > self.max_iter:
                break
            a_set, b_set = self.compute_polyhedron(q_inv, q_ellipse, p_seed, a_set_init, b_set_init)
            a_set_np = np.array(a_set)
            b_set_np = np.array(b_set)
            if not optimize:
                q_inv, p_seed = self.mvie_socp(a_set_np, b_set_np, p_hint=p_seed)
                svd = np.linalg.svd(q_inv)
                q_ellipse = svd.Vh.T @ np.diag(1 / svd.S) @ svd.U.T
                break
            det_ellipse_old = np.copy(det_ellipse)
            q_inv, q_ellipse, eigs = self.mvie_socp_fixed_r(a_set_np, b_set_np, p_seed, r_ellipse, a_lb)
            if np.min(eigs) < 1e-3:
                break
            det_ellipse = np.linalg.det(q_ellipse)
        self.last_iters = k
        return a_set, b_set, q_ellipse, p_seed

    # ---- :309-375 ---------------------------------------------------------
    def find_set_collision_avoidance(self, p0, p1, compute_ellipsoid=False, limit_space=False, e_max=0.3):
        collision = False
        if limit_space:
            a_set_init, b_set_init = self.init_halfspaces_point(p0, e_max)
        else:
            a_set_init, b_set_init = self.init_halfspaces()
        _, _, V = self._stack()
        a_set = list(a_set_init)
        b_set = list(b_set_init)
        obs_points, phi = self.compute_set_projs_line(self.obs_sets, p0, p1)
        p_closest = p0[None, :] + phi[:, None] * (p1 - p0)[None, :]
        dists = np.linalg.norm(obs_points - p_closest, axis=1)
        alive = np.ones(len(self.obs_sets), bool)
        self.last_picks = []
        while np.any(alive):
            idx = int(np.argmin(np.where(alive, dists, np.inf)))
            closest_point = obs_points[idx]
            a_h = closest_point - p_closest[idx]
            norm_a = np.linalg.norm(a_h)
            if norm_a < 1e-6:
                collision = True                             # "Line is touching an obstacle" (:337)
                a_h = closest_point - p0
                norm_a = np.linalg.norm(a_h)
                if norm_a < 1e-6:                            # "P0 is touching an obstacle" (:343)
                    a_h = p1 - p0
                    norm_a = np.linalg.norm(a_h)
            a_h = a_h / norm_a
            b_h = a_h @ closest_point - 0.001
            vmin = np.min(V @ a_h - b_h, axis=1)
            alive &= ~(vmin >= -1e-4)
            alive[idx] = False
            a_set.append(a_h)
            b_set.append(b_h)
            self.last_picks.append(idx)
        a_set_np = np.array(a_set)
        b_set_np = np.array(b_set)
        if compute_ellipsoid:
            q_inv, p_seed = self.mvie_socp(a_set_np, b_set_np, p_hint=np.asarray(p0, float))
            svd = np.linalg.svd(q_inv)
            q_ellipse = svd.Vh.T @ np.diag(1 / svd.S) @ svd.U.T
            return a_set_np, b_set_np, q_ellipse, p_seed, collision
        return a_set_np, b_set_np, collision


def line_frame(dp1):
    """Rotation used by find_set_around_line (ConvexSetFinder.py:245-258): columns dp_ref, b1, b2.
    gram_schmidt is util_functions.py:108-116."""
    dp1 = np.asarray(dp1, float)
    l_seg = np.linalg.norm(dp1)
    dp_ref = dp1 / l_seg
    if np.abs(dp_ref[2]) < 0.99:
        b1d = np.array([0, 0, 1.0])
    else:
        b1d = np.array([0, 1.0, 0])
    b1 = b1d - (dp_ref.T @ b1d) * dp_ref
    b1 /= np.linalg.norm(b1)
    b2 = np.cross(dp_ref, b1)
    b2 /= np.linalg.norm(b2)
    return np.vstack((dp_ref, b1, b2)).T, l_seg
