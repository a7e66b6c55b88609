"""CPU: pin the MVIE oracle (oracle/mvie.py) -- analytic answers, an independent
SciPy solve of the reference's own SOCP statement, and the Q1 objective quirk."""
import numpy as np
from scipy.optimize import minimize

from oracle import mvie

BOX = np.vstack((np.eye(3), -np.eye(3)))


def _random_polytope(rng, k=9):
    c = rng.uniform(-0.3, 0.3, 3)
    c[2] += 0.5
    An = rng.normal(size=(k, 3))
    An /= np.linalg.norm(An, axis=1)[:, None]
    A = np.vstack((BOX, An))
    b = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ c + rng.uniform(0.05, 0.4, k)))
    return A, b, c


def test_box_known_answer():
    b = np.array([0.2, 1.0, 1.5, 0.4, 0.5, 0.3])
    Q, c = mvie.mvie_free(BOX, b, p_hint=np.zeros(3))
    h = np.array([0.3, 0.75, 0.9])
    assert np.abs(Q - np.diag(h**2)).max() < 1e-10
    assert np.abs(c - np.array([-0.1, 0.25, 0.6])).max() < 1e-10
    Qf, cf = mvie.mvie_fixed_mid(BOX, b, np.zeros(3))
    assert np.abs(Qf - np.diag([0.2, 0.5, 0.3]) ** 2).max() < 1e-10 and np.all(cf == 0)


def _slsqp_reference_statement(A, b, centre=None):
    """Solve the reference's SOCP as stated (max t3 with the three rotated cones,
    ConvexSetFinder.py:699-743) with SLSQP, variables [L(6), (d), t1, t2, t3]."""
    free = centre is None
    n = 12 if free else 9
    it = n - 3

    def unpack(x):
        L = np.zeros((3, 3))
        L[np.tril_indices(3)] = x[:6]
        d = x[6:9] if free else centre
        return L, d

    cons = []
    for i in range(A.shape[0]):
        cons.append({"type": "ineq", "fun": lambda x, i=i: (b[i] - A[i] @ unpack(x)[1]) ** 2
                     - np.sum((unpack(x)[0].T @ A[i]) ** 2)})
        cons.append({"type": "ineq", "fun": lambda x, i=i: b[i] - A[i] @ unpack(x)[1]})
    cons += [{"type": "ineq", "fun": lambda x: x[0] * x[2] - x[it] ** 2},
             {"type": "ineq", "fun": lambda x: x[2] * x[5] - x[it + 1] ** 2},
             {"type": "ineq", "fun": lambda x: x[it] * x[it + 1] - x[it + 2] ** 2},
             {"type": "ineq", "fun": lambda x: x[0]}, {"type": "ineq", "fun": lambda x: x[2]},
             {"type": "ineq", "fun": lambda x: x[5]}, {"type": "ineq", "fun": lambda x: x[it]},
             {"type": "ineq", "fun": lambda x: x[it + 1]}]
    x0 = np.zeros(n)
    x0[[0, 2, 5]] = 0.02
    if free:
        x0[6:9] = mvie.chebyshev_centre(A, b)
    x0[it:] = 0.01
    res = minimize(lambda x: -x[-1], x0, constraints=cons, method="SLSQP", options={"maxiter": 500, "ftol": 1e-14})
    L, d = unpack(res.x)
    return L @ L.T, d


def test_matches_independent_solver_on_reference_statement():
    rng = np.random.default_rng(7)
    for _ in range(4):
        A, b, c = _random_polytope(rng)
        Q, cen = mvie.mvie_free(A, b, p_hint=c)
        Qs, cs = _slsqp_reference_statement(A, b)
        assert np.abs(Q - Qs).max() <= 2e-5 * np.abs(Q).max()
        assert np.abs(cen - cs).max() <= 2e-5
        Qf, _ = mvie.mvie_fixed_mid(A, b, c)
        Qfs, _ = _slsqp_reference_statement(A, b, centre=c)
        assert np.abs(Qf - Qfs).max() <= 2e-5 * np.abs(Qf).max()


def test_objective_is_double_weighted_not_logdet():
    """Quirk Q1: the optimum maximises L00 * L11^2 * L22, which differs from the
    log-det (true max-volume) ellipsoid on a generic polytope."""
    rng = np.random.default_rng(11)
    A, b, c = _random_polytope(rng, 7)
    Q, cen = mvie.mvie_free(A, b, p_hint=c)
    L = np.linalg.cholesky(Q)
    w_obj = L[0, 0] * L[1, 1] ** 2 * L[2, 2]
    # perturb along feasible directions: shrink slightly and re-orient -> objective must not improve
    for _ in range(50):
        dL = np.tril(rng.normal(size=(3, 3))) * 1e-3
        L2 = (L + dL) * 0.999
        u = np.linalg.norm(A @ L2, axis=1)
        if np.all(u <= b - A @ cen):
            assert L2[0, 0] * L2[1, 1] ** 2 * L2[2, 2] <= w_obj * (1 + 1e-9)
    # and the KKT polish agrees with the pure barrier solution
    Q2, cen2 = mvie.mvie_free(A, b, p_hint=c, polish=False)
    assert np.abs(Q - Q2).max() < 1e-9 and np.abs(cen - cen2).max() < 1e-9


def test_padded_rows_and_inscribed():
    rng = np.random.default_rng(3)
    A, b, c = _random_polytope(rng)
    Ap = np.vstack((A, np.zeros((5, 3))))
    bp = np.concatenate((b, 10 * np.ones(5)))
    Q1, c1 = mvie.mvie_free(A, b, p_hint=c)
    Q2, c2 = mvie.mvie_free(Ap, bp, p_hint=c)
    assert np.abs(Q1 - Q2).max() < 1e-12 and np.abs(c1 - c2).max() < 1e-12
    # {L u + d : |u| <= 1} is inside the polytope and touches it
    L = np.linalg.cholesky(Q1)
    slack = b - A @ c1 - np.linalg.norm(A @ L, axis=1)
    assert slack.min() > -1e-9 and np.sum(slack < 1e-7) >= 3


def test_fixed_r_axis_aligned_box():
    b = np.array([0.2, 1.0, 1.5, 0.4, 0.5, 0.3])
    q_new, q_ell, s = mvie.mvie_fixed_r(BOX, b, np.zeros(3), np.eye(3), 0.05)
    assert np.abs(s - np.array([0.2, 0.5, 0.3])).max() < 1e-9
    assert np.abs(q_new @ q_ell - np.eye(3)).max() < 1e-9
