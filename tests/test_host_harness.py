"""CPU: the thread-serial device primitives (csrc/bp_*.cuh), compiled for the host by
tests/host_harness.cpp, against the oracle.  This is the exact code the kernels run
(the warp-cooperative kernels follow the same algorithm, constants and row order)."""
import ctypes
import os

import numpy as np

from oracle import fk_iiwa14 as ofk
from oracle import mvie as omvie
from oracle.convex_set_finder import closest_points_segment_boxes, min_norm_point_polytopes
from oracle.set_graph import intersection_margin, set_intersection

P = ctypes.POINTER(ctypes.c_double)
BOX = np.vstack((np.eye(3), -np.eye(3)))


def dp(a):
    return a.ctypes.data_as(P)


def test_mvie_primitive(host_harness):
    rng = np.random.default_rng(3)
    worst, iters = 0.0, []
    for _ in range(25):
        k = rng.integers(3, 20)
        c = rng.uniform(-0.5, 0.5, 3)
        c[2] += 0.6
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A = np.ascontiguousarray(np.vstack((BOX, An)))
        b = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.005, 0.4, k)))
        if np.min(b - A @ c) <= 1e-3:
            continue
        for free in (0, 1):
            E, Q, cen, it = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
            st = host_harness.hh_mvie(dp(A), dp(b), A.shape[0], free, dp(c), dp(E), dp(Q), dp(cen), ctypes.byref(it))
            Eo, co = omvie.mvie_free(A, b, p_hint=c) if free else omvie.mvie_fixed_mid(A, b, c)
            assert st == 0
            worst = max(worst, np.abs(E - Eo).max() / np.abs(Eo).max(), np.abs(cen - co).max())
            assert np.abs(Q @ E - np.eye(3)).max() < 1e-9
            iters.append(it.value)
    assert worst < 1e-9
    assert np.mean(iters) < 45          # path-following with the secant predictor


def test_mvie_primitive_rejects_exterior_centre(host_harness):
    b = np.array([1, 1, 1, 1, 1, 1.0])
    E, Q, cen, it = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
    st = host_harness.hh_mvie(dp(BOX.copy()), dp(b), 6, 0, dp(np.array([2.0, 0, 0])), dp(E), dp(Q), dp(cen),
                              ctypes.byref(it))
    assert st == 3                      # BP_MVIE_NO_INTERIOR


def test_box_qp_primitive(host_harness):
    rng = np.random.default_rng(4)
    N = 300
    lb = np.ascontiguousarray(rng.uniform(-1, 1, (N, 3)))
    ub = np.ascontiguousarray(lb + rng.uniform(0.02, 0.3, (N, 3)))
    for trial in range(4):
        p = rng.uniform(-1, 1, 3)
        L = np.tril(rng.normal(size=(3, 3))) * 0.2
        L[np.diag_indices(3)] = rng.uniform(0.05, 0.5, 3)
        E = np.ascontiguousarray(L @ L.T if trial else np.diag([1e-4] * 3))
        y, dist = np.zeros((N, 3)), np.zeros(N)
        host_harness.hh_box_qp(dp(E), dp(p), dp(lb), dp(ub), N, dp(y), dp(dist))
        A = np.broadcast_to(BOX, (N, 6, 3))
        x = min_norm_point_polytopes(A @ E, np.hstack((ub, -lb)) - A @ p)
        yo = x @ E.T + p
        do = np.linalg.norm(np.linalg.inv(E) @ (yo - p).T, axis=0)
        assert np.abs(y - yo).max() < 1e-7
        assert np.abs(dist - do).max() < 1e-7 * do.max()


def test_seg_box_primitive(host_harness):
    rng = np.random.default_rng(5)
    N = 300
    lb = np.ascontiguousarray(rng.uniform(-1, 1, (N, 3)))
    ub = np.ascontiguousarray(lb + rng.uniform(0.02, 0.3, (N, 3)))
    for trial in range(5):
        p0 = rng.uniform(-1, 1, 3)
        p1 = p0 + (np.array([0.4, 0, 0]) if trial == 3 else rng.normal(size=3) * 0.3)
        x, phi = np.zeros((N, 3)), np.zeros(N)
        host_harness.hh_seg_box(dp(p0), dp(p1), dp(lb), dp(ub), N, dp(x), dp(phi))
        xo, phio = closest_points_segment_boxes(lb, ub, p0, p1)
        assert np.abs(x - xo).max() < 1e-12 and np.abs(phi - phio).max() < 1e-12


def test_pair_lp_primitive_vs_highs(host_harness):
    rng = np.random.default_rng(6)
    sets = []
    for _ in range(45):
        k = rng.integers(4, 14)
        c = rng.uniform(-0.6, 0.6, 3)
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        sets.append([np.ascontiguousarray(np.vstack((BOX, An))),
                     np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.05, 0.5, k)))])
    n_yes = 0
    for i in range(45):
        for j in range(i):
            xo, it = np.zeros(3), ctypes.c_int()
            r = host_harness.hh_pair_lp(dp(sets[i][0]), dp(sets[i][1]), sets[i][0].shape[0], dp(sets[j][0]),
                                        dp(sets[j][1]), sets[j][0].shape[0], ctypes.c_double(0.01), dp(xo),
                                        ctypes.byref(it))
            ok = bool(set_intersection(sets[i], sets[j], 0.01)[2])
            if bool(r) != ok:
                assert abs(intersection_margin(sets[i], sets[j], 0.01)) < 1e-6
            n_yes += r
            if r:       # the returned iterate is a point of the shrunk intersection
                assert np.max(sets[i][0] @ xo - sets[i][1]) <= -0.01 + 1e-12
                assert np.max(sets[j][0] @ xo - sets[j][1]) <= -0.01 + 1e-12
    assert 50 < n_yes < 900


def test_fk_primitive(host_harness):
    rng = np.random.default_rng(7)
    n = 100
    q = np.ascontiguousarray(rng.uniform(ofk.Q_LOWER, ofk.Q_UPPER, (n, 7)))
    pe, pc, T, J = np.zeros((n, 3)), np.zeros((n, 7, 3)), np.zeros((n, 4, 4)), np.zeros((n, 6, 7))
    host_harness.hh_fk(dp(q), n, dp(pe), dp(pc), dp(T), dp(J))
    for i in range(n):
        assert np.abs(pe[i] - ofk.fk_pos(q[i])).max() < 1e-12
        assert np.abs(pc[i] - ofk.fk_pos_col_all(q[i])).max() < 1e-12
        assert np.abs(T[i] - ofk.hom_transform_endeffector(q[i])).max() < 1e-12
        assert np.abs(J[i] - ofk.jacobian_fk(q[i])).max() < 1e-12


def test_fk_jacobian_time_variation_primitive(host_harness):
    """djacobian_fk (RobotModel.py:233-251) of the device code (host build) vs the oracle and vs the reference's
    jacobian.ca differentiated along dq (golden)."""
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_reference_blobs.npz"))
    q, dq = np.ascontiguousarray(ref["q"]), np.ascontiguousarray(ref["dq"])
    n = q.shape[0]
    T, J, dJ = np.zeros((n, 4, 4)), np.zeros((n, 6, 7)), np.zeros((n, 6, 7))
    host_harness.hh_fk_kin(dp(q), dp(dq), n, dp(T), dp(J), dp(dJ))
    assert np.abs(J - ref["jacobian"]).max() < 1e-12 and np.abs(T - ref["hom_trans"]).max() < 1e-12
    assert np.abs(dJ - ref["djacobian"]).max() < 1e-12
    for i in range(0, n, 7):
        assert np.abs(dJ[i] - ofk.djacobian_fk(q[i], dq[i])).max() < 1e-12
    # finite-difference cross-check of the oracle formula
    h = 1e-6
    fd = (ofk.jacobian_fk(q[3] + h * dq[3]) - ofk.jacobian_fk(q[3] - h * dq[3])) / (2 * h)
    assert np.abs(fd - dJ[3]).max() < 1e-8


def test_min_eig_primitive(host_harness):
    rng = np.random.default_rng(8)
    for _ in range(50):
        L = np.tril(rng.normal(size=(3, 3)))
        L[np.diag_indices(3)] = rng.uniform(1e-3, 1, 3)
        E = np.ascontiguousarray(L @ L.T)
        assert abs(host_harness.hh_min_eig(dp(E)) - np.linalg.svd(E)[1].min()) < 1e-13 * max(1.0, np.abs(E).max())


def _random_frame(rng):
    dp1 = rng.normal(size=3) * rng.uniform(0.05, 0.6)
    R, l = np.zeros((3, 3)), ctypes.c_double()
    return dp1, R, l


def test_line_frame_primitive(host_harness):
    """bp_line_frame against find_set_around_line's frame (ConvexSetFinder.py:245-258)."""
    from oracle.convex_set_finder import line_frame

    rng = np.random.default_rng(11)
    cases = [rng.normal(size=3) for _ in range(10)] + [np.array([0, 0, 0.3]), np.array([1e-3, 0, -0.5])]
    for dp1 in cases:
        R, l = np.zeros((3, 3)), ctypes.c_double()
        host_harness.hh_line_frame(dp(np.ascontiguousarray(dp1)), dp(R), ctypes.byref(l))
        Ro, lo = line_frame(dp1)
        assert abs(l.value - lo) < 1e-15 and np.abs(R - Ro).max() < 1e-14
        assert np.abs(R.T @ R - np.eye(3)).max() < 1e-12


def test_mvie_fixed_r_primitive(host_harness):
    """mvie_socp_fixed_r (ConvexSetFinder.py:564-588): device primitive vs the oracle and analytic boxes."""
    from oracle.convex_set_finder import line_frame

    rng = np.random.default_rng(12)
    host_harness.hh_mvie_fixed_r.restype = ctypes.c_int
    worst = 0.0
    n_done = 0
    for trial in range(30):
        k = rng.integers(2, 14)
        c = rng.uniform(-0.4, 0.4, 3)
        c[2] += 0.6
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A = np.ascontiguousarray(np.vstack((BOX, An)))
        b = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.02, 0.4, k)))
        R, _ = line_frame(rng.normal(size=3))
        R = np.ascontiguousarray(R)
        a_lb = 0.0 if trial % 3 == 0 else rng.uniform(0.0, 0.02)
        E, Q, eigs, it = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
        st = host_harness.hh_mvie_fixed_r(dp(A), dp(b), A.shape[0], dp(c), dp(R), ctypes.c_double(a_lb), dp(E), dp(Q),
                                          dp(eigs), ctypes.byref(it))
        try:
            Eo, Qo, so = omvie.mvie_fixed_r(A, b, c, R, a_lb)
        except omvie.MVIEError:
            assert st == 3
            continue
        assert st == 0
        assert eigs[0] >= a_lb - 1e-12
        worst = max(worst, np.abs(eigs - so).max() / so.max())
        assert np.abs(E - R @ np.diag(eigs**2) @ R.T).max() < 1e-14
        assert np.abs(Q @ E - np.eye(3)).max() < 1e-9
        n_done += 1
    assert n_done >= 20 and worst < 1e-7
    # analytic: axis-aligned box, R = I  ->  semi-axes = half widths; with a_lb above the optimum of x0 the bound is active
    b = np.array([0.3, 0.2, 0.5, 0.3, 0.2, 0.5])
    E, Q, eigs, it = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
    st = host_harness.hh_mvie_fixed_r(dp(BOX.copy()), dp(b), 6, dp(np.zeros(3)), dp(np.eye(3)), ctypes.c_double(0.0),
                                      dp(E), dp(Q), dp(eigs), ctypes.byref(it))
    assert st == 0 and np.abs(eigs - b[:3]).max() < 1e-9
    st = host_harness.hh_mvie_fixed_r(dp(BOX.copy()), dp(b), 6, dp(np.zeros(3)), dp(np.eye(3)), ctypes.c_double(0.31),
                                      dp(E), dp(Q), dp(eigs), ctypes.byref(it))
    assert st == 3                      # a_lb larger than the box allows: the reference's SOCP is infeasible


def test_polytope_qp_primitive(host_harness):
    """General-polytope closest points (bp_polytope_qp) vs the oracle's active-set enumeration, ragged row counts."""
    from boundplanner_b200 import scenes

    rng = np.random.default_rng(17)
    obs_sets, obs_points = scenes.random_polytope_scene(120, rng)
    R = 15
    rows4 = np.ascontiguousarray(np.stack([np.hstack((s[0], s[1][:, None])) for s in obs_sets]))
    assert len({int((np.linalg.norm(s[0], axis=1) > 0).sum()) for s in obs_sets}) > 3        # ragged
    worst = 0.0
    for trial in range(4):
        p = rng.uniform(-1, 1, 3)
        p[2] = abs(p[2])
        L = np.tril(rng.normal(size=(3, 3))) * 0.2 + np.diag(rng.uniform(0.05, 0.6, 3))
        E = L @ L.T if trial else 1e-4 * np.eye(3)
        E = np.ascontiguousarray(E)
        y, dist = np.zeros((120, 3)), np.zeros(120)
        host_harness.hh_polytope_qp.restype = ctypes.c_int
        n_empty = host_harness.hh_polytope_qp(dp(E), dp(np.ascontiguousarray(p)), dp(rows4), 120, R, dp(y), dp(dist))
        assert n_empty == 0
        A = rows4[:, :, :3]
        b = rows4[:, :, 3]
        x = min_norm_point_polytopes(A @ E, b - A @ p)
        yo = x @ E.T + p
        do = np.linalg.norm(np.linalg.solve(E, (yo - p).T), axis=0)
        worst = max(worst, np.abs(y - yo).max(), (np.abs(dist - do) / np.maximum(do, 1e-12)).max())
        # the points lie in their polytopes and (when p is outside) on their boundary
        viol = np.einsum("nrk,nk->nr", A, y) - b
        assert viol.max() < 1e-9
    assert worst < 1e-7          # (north_star tolerance 1e-6; the oracle's own Gram solves limit the agreement)


def test_segment_polytope_primitive(host_harness):
    """Segment-polytope closest points (bp_seg_polytope_qp): boxes given as polytopes reproduce the box oracle
    (incl. the smallest-phi rule, quirk Q9); random polytopes agree with an SLSQP solve of the reference's QP
    (ConvexSetFinder.py:52-99)."""
    from scipy.optimize import minimize

    from boundplanner_b200 import scenes

    host_harness.hh_seg_polytope.restype = ctypes.c_int
    rng = np.random.default_rng(23)
    # (1) boxes as polytopes
    N = 200
    lb = rng.uniform(-1, 1, (N, 3))
    ub = lb + rng.uniform(0.02, 0.3, (N, 3))
    rows4 = np.zeros((N, 15, 4))
    rows4[:, :, 3] = 10.0
    rows4[:, :3, :3] = np.eye(3)
    rows4[:, 3:6, :3] = -np.eye(3)
    rows4[:, :3, 3] = ub
    rows4[:, 3:6, 3] = -lb
    rows4 = np.ascontiguousarray(rows4)
    for trial in range(5):
        p0 = rng.uniform(-1, 1, 3)
        p1 = p0 + rng.normal(size=3) * rng.uniform(0.05, 0.8)
        if trial == 3:
            p1 = p0 + np.array([0.0, 0.0, 0.4])            # axis-aligned: non-unique minimisers are common
        if trial == 4:
            p0 = 0.5 * (lb[0] + ub[0]); p1 = p0 + np.array([0.5, 0.1, 0.0])    # starts inside box 0
        p0 = np.ascontiguousarray(p0); p1 = np.ascontiguousarray(p1)
        x, phi, d2 = np.zeros((N, 3)), np.zeros(N), np.zeros(N)
        assert host_harness.hh_seg_polytope(dp(p0), dp(p1), dp(rows4), N, 15, ctypes.c_double(0.001), dp(x), dp(phi),
                                            dp(d2)) == 0
        xo, phio = closest_points_segment_boxes(lb + 0.001, ub - 0.001, p0, p1)
        do2 = np.sum((p0 + phio[:, None] * (p1 - p0) - xo) ** 2, axis=1)
        assert np.abs(d2 - do2).max() < 1e-12
        assert np.abs(phi - phio).max() < 1e-9
        assert np.abs(x - xo).max() < 1e-9
    # (2) random polytopes vs SLSQP on the QP itself
    obs_sets, _ = scenes.random_polytope_scene(12, rng, 0.1, 0.5)
    rows4 = np.ascontiguousarray(np.stack([np.hstack((s[0], s[1][:, None])) for s in obs_sets]))
    p0 = np.ascontiguousarray(rng.uniform(-1, 1, 3)); p1 = np.ascontiguousarray(p0 + rng.normal(size=3) * 0.5)
    n = len(obs_sets)
    x, phi, d2 = np.zeros((n, 3)), np.zeros(n), np.zeros(n)
    assert host_harness.hh_seg_polytope(dp(p0), dp(p1), dp(rows4), n, 15, ctypes.c_double(0.001), dp(x), dp(phi),
                                        dp(d2)) == 0
    for j, (A, b) in enumerate(obs_sets):
        nz = np.linalg.norm(A, axis=1) > 0
        A, b = A[nz], b[nz] - 0.001
        assert np.max(A @ x[j] - b) < 1e-9 and -1e-12 <= phi[j] <= 1 + 1e-12
        fun = lambda z: np.sum((p0 + z[3] * (p1 - p0) - z[:3]) ** 2)
        cons = [{"type": "ineq", "fun": lambda z, A=A, b=b: b - A @ z[:3]}]
        best = np.inf
        for z3 in (0.0, 0.5, 1.0):
            res = minimize(fun, np.concatenate((x[j] + 1e-3, [z3])), constraints=cons, bounds=[(None, None)] * 3 + [(0, 1)],
                           method="SLSQP", options={"ftol": 1e-14, "maxiter": 300})
            if res.success:
                best = min(best, res.fun)
        assert d2[j] <= best + 1e-9 and abs(d2[j] - best) < 1e-7 * max(1.0, best)


def test_mvie_primal_dual_specification(host_harness):
    """bp_mvie_pd.cuh (barrier centre -> primal-dual -> final barrier stages, the specification of the next MVIE
    kernel) against the barrier solver and the oracle: same ellipsoid, fewer Newton iterations."""
    rng = np.random.default_rng(3)
    it_pd, it_bar, worst, worst_o = [], [], 0.0, 0.0
    for _ in range(25):
        k = rng.integers(3, 20)
        c = rng.uniform(-0.5, 0.5, 3)
        c[2] += 0.6
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A = np.ascontiguousarray(np.vstack((BOX, An)))
        b = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.005, 0.4, k)))
        if np.min(b - A @ c) <= 1e-3:
            continue
        # padded rows (A = 0, b = 10) carry no cone
        Ap = np.ascontiguousarray(np.vstack((A, np.zeros((2, 3)))))
        bp_ = np.concatenate((b, [10.0, 10.0]))
        for free in (0, 1):
            E, Q, cen, it = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
            ph = (ctypes.c_int * 3)()
            st = host_harness.hh_mvie_pd(dp(Ap), dp(bp_), Ap.shape[0], free, dp(c), dp(E), dp(Q), dp(cen), ctypes.byref(it),
                                         ph)
            assert st == 0
            E0, Q0, cen0, it0 = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), ctypes.c_int()
            assert host_harness.hh_mvie(dp(A), dp(b), A.shape[0], free, dp(c), dp(E0), dp(Q0), dp(cen0), ctypes.byref(it0)) == 0
            Eo, co = omvie.mvie_free(A, b, p_hint=c) if free else omvie.mvie_fixed_mid(A, b, c)
            worst = max(worst, np.abs(E - E0).max() / np.abs(E0).max(), np.abs(cen - cen0).max())
            worst_o = max(worst_o, np.abs(E - Eo).max() / np.abs(Eo).max(), np.abs(cen - co).max())
            assert ph[1] < 30                      # the primal-dual phase converged (no fall-back to the barrier path)
            it_pd.append(it.value)
            it_bar.append(it0.value)
    assert worst < 1e-10 and worst_o < 1e-9        # both end with the same barrier stages
    assert np.mean(it_pd) < 0.85 * np.mean(it_bar)
