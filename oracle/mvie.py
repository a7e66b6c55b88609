"""Oracle (TEST INFRASTRUCTURE): maximum-volume inscribed ellipsoid SOCPs.

Restates the three SOCPs the reference hands to CVXPY/Clarabel
(bound_planner/BoundPlanner/ConvexSetFinder.py):

* ``mvie_socp``            :512-537  (problem factory :590-618, cones :699-720)
* ``mvie_socp_fixed_mid``  :539-562  (factory :620-648, cones :722-743)
* ``mvie_socp_fixed_r``    :564-588  (factory :650-680, cones :745-766)

The reference's decision vector is x = [shape vars, (centre), t1, t2, t3] and the
objective is ``max t3`` under three rotated-cone constraints
``t1^2 <= x_p x_q, t2^2 <= x_q x_r, t3^2 <= t1 t2`` (:699-766).  At the optimum
all three are tight, so ``t3 = (x_p x_q^2 x_r)^(1/4)``: the MIDDLE pivot x_q is
used by two cones and is therefore weighted twice.  Because
``t -> t^(1/4)`` is monotone the problem is equivalent to

    maximise  log x_p + 2 log x_q + log x_r
    s.t.      || a2_i^T x || <= c2_i . x + d2_i     (i < m, one cone per row)

which is what is solved here (primal log-barrier, damped Newton, followed by an
active-set KKT polish).  Clarabel is a third-party dependency absent from
/root/reference (cvxpy==1.6.3 pins it transitively, requirements.txt:6); parity
against it is UNPINNED -- this file is checked against analytic answers and
SciPy's SLSQP in tests/test_oracle_mvie.py.
"""
from __future__ import annotations

import numpy as np

# weights of (first, middle, last) pivot in the equivalent log objective (Q1)
_W = np.array([1.0, 2.0, 1.0])


def _row_operators_L(a_set: np.ndarray, n: int) -> np.ndarray:
    """J[i] (3 x n) with J[i] @ x = L^T a_i, L packed as tril order
    [L00, L10, L11, L20, L21, L22] (ConvexSetFinder.py:519-522, :534)."""
    m = a_set.shape[0]
    J = np.zeros((m, 3, n))
    J[:, 0, 0] = a_set[:, 0]
    J[:, 0, 1] = a_set[:, 1]
    J[:, 0, 3] = a_set[:, 2]
    J[:, 1, 2] = a_set[:, 1]
    J[:, 1, 4] = a_set[:, 2]
    J[:, 2, 5] = a_set[:, 2]
    return J


class MVIEError(RuntimeError):
    pass


def _solve_logdet_soc(J, c, d, diag_idx, x0, extra_lb=None, tol=1e-11, max_newton=400):
    """maximise sum_k W[k] log x[diag_idx[k]]  s.t. ||J_i x|| <= c_i.x + d_i,
    optionally x[0] >= extra_lb (fixed-R variant, ConvexSetFinder.py:667-669).

    Log-barrier path following; returns x accurate to ~tol * scale."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    m = J.shape[0]

    def parts(xv):
        u = J @ xv                      # (m,3)
        s = c @ xv + d                  # (m,)
        psi = s * s - np.einsum("ij,ij->i", u, u)
        return u, s, psi

    def feasible(xv):
        u, s, psi = parts(xv)
        ok = np.all(s > 0) and np.all(psi > 0) and np.all(xv[diag_idx] > 0)
        if extra_lb is not None:
            ok = ok and xv[0] - extra_lb > 0
        return ok

    def value(xv, t):
        u, s, psi = parts(xv)
        v = -t * np.dot(_W, np.log(xv[diag_idx])) - np.sum(np.log(psi))
        if extra_lb is not None:
            v -= np.log(xv[0] - extra_lb)
        return v

    if not feasible(x):
        raise MVIEError("MVIE start point is not strictly feasible")

    nu = 2.0 * m + 4.0 + (1.0 if extra_lb is not None else 0.0)
    t = 1.0
    it = 0
    while True:
        # centering
        for _ in range(60):
            it += 1
            if it > max_newton:
                break
            u, s, psi = parts(x)
            # half-gradient of psi: v_i = s_i c_i - J_i^T u_i
            v = s[:, None] * c - np.einsum("ikj,ik->ij", J, u)      # (m,n)
            g = -2.0 * (v / psi[:, None]).sum(axis=0)
            H = 4.0 * np.einsum("i,ij,ik->jk", 1.0 / psi**2, v, v)
            H += 2.0 * np.einsum("i,ilj,ilk->jk", 1.0 / psi, J, J)
            H -= 2.0 * np.einsum("i,ij,ik->jk", 1.0 / psi, c, c)
            xd = x[diag_idx]
            g[diag_idx] += -t * _W / xd
            H[diag_idx, diag_idx] += t * _W / xd**2
            if extra_lb is not None:
                r = x[0] - extra_lb
                g[0] += -1.0 / r
                H[0, 0] += 1.0 / r**2
            try:
                Lc = np.linalg.cholesky(H)
            except np.linalg.LinAlgError:
                Lc = np.linalg.cholesky(H + 1e-14 * np.trace(H) * np.eye(n))
            dx = -np.linalg.solve(Lc.T, np.linalg.solve(Lc, g))
            lam2 = float(-g @ dx)
            if lam2 < 1e-20:
                break
            step = 1.0
            f0 = value(x, t)
            while True:
                xn = x + step * dx
                if feasible(xn) and value(xn, t) <= f0 - 0.25 * step * lam2:
                    break
                step *= 0.5
                if step < 1e-12:
                    break
            if step < 1e-12:
                break
            x = xn
            if lam2 < 1e-16 * max(1.0, t):
                break
        if nu / t < tol or it > max_newton:
            break
        t *= 20.0
    return x


def _polish(J, c, d, diag_idx, x, act_tol=1e-6, iters=6):
    """Active-set KKT Newton polish: solve grad f0 + sum_i lam_i grad g_i = 0,
    g_i = 0 for the rows that are tight at the barrier solution.  Falls back to
    the barrier point when the polish does not contract."""
    n = x.size
    x = x.copy()

    def gfun(xv):
        u = J @ xv
        s = c @ xv + d
        return np.linalg.norm(u, axis=1) - s, u, s

    g, u, s = gfun(x)
    act = np.where(g > -act_tol * np.maximum(1.0, np.abs(s)))[0]
    if act.size == 0 or act.size > n:
        return x
    lam = None
    best = (np.inf, x.copy())
    for _ in range(iters):
        g, u, s = gfun(x)
        nu_ = np.linalg.norm(u[act], axis=1)
        if np.any(nu_ < 1e-14):
            return best[1]
        # gradients of active constraints
        G = np.einsum("ikj,ik->ij", J[act], u[act] / nu_[:, None]) - c[act]     # (k,n)
        gf = np.zeros(n)
        gf[diag_idx] = -_W / x[diag_idx]
        if lam is None:
            lam = np.linalg.lstsq(G.T, -gf, rcond=None)[0]
        # Lagrangian Hessian
        Hl = np.zeros((n, n))
        Hl[diag_idx, diag_idx] = _W / x[diag_idx] ** 2
        for k, i in enumerate(act):
            P = (np.eye(3) - np.outer(u[i], u[i]) / nu_[k] ** 2) / nu_[k]
            Hl += lam[k] * J[i].T @ P @ J[i]
        r1 = gf + G.T @ lam
        r2 = g[act]
        res = max(np.max(np.abs(r1)), np.max(np.abs(r2)))
        if res < best[0]:
            best = (res, x.copy())
        K = np.block([[Hl, G.T], [G, np.zeros((act.size, act.size))]])
        try:
            sol = np.linalg.solve(K, -np.concatenate((r1, r2)))
        except np.linalg.LinAlgError:
            return best[1]
        if not np.all(np.isfinite(sol)):
            return best[1]
        x = x + sol[:n]
        lam = lam + sol[n:]
    g, u, s = gfun(x)
    gf = np.zeros(n)
    if np.any(x[diag_idx] <= 0) or np.any(lam < -1e-9) or np.max(g) > 1e-12:
        return best[1]
    return x


def _start_radius(a_set, slack):
    nrm = np.linalg.norm(a_set, axis=1)
    ok = nrm > 1e-12
    if not np.all(slack[~ok] > 0):
        raise MVIEError("padded row infeasible")
    return 0.5 * np.min(slack[ok] / nrm[ok])


def mvie_free(a_set, b_set, p_hint=None, polish=True):
    """Free-centre MVIE, ConvexSetFinder.py:512-537.  Returns (L L^T, centre)."""
    a_set = np.asarray(a_set, float)
    b_set = np.asarray(b_set, float)
    keep = np.linalg.norm(a_set, axis=1) > 0
    A, b = a_set[keep], b_set[keep]
    m = A.shape[0]
    J = _row_operators_L(A, 9)
    c = np.zeros((m, 9))
    c[:, 6:9] = -A
    d = b.copy()
    if p_hint is None or np.min(b - A @ p_hint) <= 0:
        p_hint = chebyshev_centre(A, b)
    r = _start_radius(A, b - A @ p_hint)
    if r <= 0:
        raise MVIEError("polytope has empty interior")
    x0 = np.array([r, 0, r, 0, 0, r, *p_hint])
    di = np.array([0, 2, 5])
    x = _solve_logdet_soc(J, c, d, di, x0)
    if polish:
        x = _polish(J, c, d, di, x)
    Lm = np.zeros((3, 3))
    Lm[np.tril_indices(3)] = x[:6]
    return Lm @ Lm.T, x[6:9].copy()


def mvie_fixed_mid(a_set, b_set, p_mid, polish=True):
    """Fixed-centre MVIE, ConvexSetFinder.py:539-562.  Returns (L L^T, p_mid)."""
    a_set = np.asarray(a_set, float)
    b_set = np.asarray(b_set, float)
    keep = np.linalg.norm(a_set, axis=1) > 0
    A, b = a_set[keep], b_set[keep]
    m = A.shape[0]
    J = _row_operators_L(A, 6)
    c = np.zeros((m, 6))
    d = b - A @ p_mid
    r = _start_radius(A, d)
    if r <= 0:
        raise MVIEError("centre is not strictly inside the polytope")
    x0 = np.array([r, 0, r, 0, 0, r])
    di = np.array([0, 2, 5])
    x = _solve_logdet_soc(J, c, d, di, x0)
    if polish:
        x = _polish(J, c, d, di, x)
    Lm = np.zeros((3, 3))
    Lm[np.tril_indices(3)] = x[:6]
    return Lm @ Lm.T, p_mid


def mvie_fixed_r(a_set, b_set, p_mid, r_ellipse, a_lb, polish=True):
    """Fixed-rotation MVIE, ConvexSetFinder.py:564-588.
    Returns (R diag(s^2) R^T, R diag(1/s^2) R^T, s)."""
    a_set = np.asarray(a_set, float)
    b_set = np.asarray(b_set, float)
    keep = np.linalg.norm(a_set, axis=1) > 0
    A, b = a_set[keep], b_set[keep]
    m = A.shape[0]
    Ar = A @ r_ellipse
    J = np.zeros((m, 3, 3))
    for k in range(3):
        J[:, k, k] = Ar[:, k]
    c = np.zeros((m, 3))
    d = b - A @ p_mid
    if np.min(d) <= 0:
        raise MVIEError("centre is not strictly inside the polytope")
    with np.errstate(divide="ignore"):
        s0_max = np.min(d / np.abs(Ar[:, 0]))
    if s0_max <= a_lb:
        raise MVIEError("fixed-R MVIE infeasible: a_lb too large")
    s0 = 0.5 * (a_lb + min(s0_max, a_lb + 2 * _start_radius(A, d)))
    # strictly feasible start: shrink the two minor axes until inside
    e = _start_radius(A, d)
    while True:
        x0 = np.array([s0, e, e])
        u = np.einsum("ikj,j->ik", J, x0)
        if np.all(d * d - np.einsum("ij,ij->i", u, u) > 0):
            break
        e *= 0.5
        if e < 1e-300:
            raise MVIEError("fixed-R MVIE: no strictly feasible start")
    di = np.array([0, 1, 2])
    x = _solve_logdet_soc(J, c, d, di, x0, extra_lb=a_lb)
    s = x[:3]
    q_new = r_ellipse @ np.diag(s) ** 2 @ r_ellipse.T
    q_ell = r_ellipse @ np.diag(1.0 / s**2) @ r_ellipse.T
    return q_new, q_ell, s


def chebyshev_centre(A, b):
    """Deepest point of {Ax<=b} (only used to start the free-centre barrier
    when no interior hint is available)."""
    from scipy.optimize import linprog

    nrm = np.linalg.norm(A, axis=1)
    res = linprog(
        np.array([0, 0, 0, -1.0]),
        A_ub=np.hstack((A, nrm[:, None])),
        b_ub=b,
        bounds=[(None, None)] * 3 + [(0, None)],
    )
    if not res.success or res.x[3] <= 0:
        raise MVIEError("polytope has empty interior")
    return res.x[:3]
