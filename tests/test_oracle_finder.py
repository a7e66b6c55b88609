"""CPU: pin the ConvexSetFinder oracle -- closest-point QPs against independent
SciPy solves and analytic cases, structural invariants of the greedy polyhedron
and of the IRIS loop, on the C1 example scene and a seeded random scene."""
import numpy as np
from scipy.optimize import minimize

from boundplanner_b200 import scenes
from oracle.convex_set_finder import closest_points_segment_boxes, min_norm_point_polytopes
from tests.util import oracle_finder

BOX = np.vstack((np.eye(3), -np.eye(3)))


def test_closest_point_initial_sphere_is_clip():
    """With E = eps*I the closest point of a box is clip(p, lb, ub)."""
    rng = np.random.default_rng(0)
    lb = rng.uniform(-1, 1, (50, 3))
    ub = lb + rng.uniform(0.05, 0.4, (50, 3))
    p = rng.uniform(-1, 1, 3)
    E = np.diag([1e-4] * 3)
    A = np.broadcast_to(BOX, (50, 6, 3))
    h = np.hstack((ub, -lb)) - A @ p
    x = min_norm_point_polytopes(A @ E, h)
    assert np.abs(x @ E.T + p - np.clip(p, lb, ub)).max() < 1e-12


def test_closest_point_metric_vs_scipy():
    rng = np.random.default_rng(1)
    for _ in range(6):
        lb = rng.uniform(-1, 1, 3)
        ub = lb + rng.uniform(0.05, 0.4, 3)
        p = rng.uniform(-1, 1, 3)
        L = np.tril(rng.normal(size=(3, 3))) * 0.1
        L[np.diag_indices(3)] = rng.uniform(0.05, 0.5, 3)
        E = L @ L.T
        M = np.linalg.inv(E) @ np.linalg.inv(E)
        x = min_norm_point_polytopes((BOX @ E)[None], (np.concatenate((ub, -lb)) - BOX @ p)[None])[0]
        y = E @ x + p
        res = minimize(lambda v: (v - p) @ M @ (v - p), np.clip(p, lb, ub), jac=lambda v: 2 * M @ (v - p),
                       bounds=list(zip(lb, ub)), method="L-BFGS-B", options={"ftol": 1e-15, "gtol": 1e-12})
        assert (y - p) @ M @ (y - p) <= res.fun * (1 + 1e-7) + 1e-12
        assert np.all(y >= lb - 1e-12) and np.all(y <= ub + 1e-12)


def test_segment_box_vs_scipy_and_ties():
    rng = np.random.default_rng(2)
    lb = rng.uniform(-1, 1, (40, 3))
    ub = lb + rng.uniform(0.05, 0.4, (40, 3))
    p0 = rng.uniform(-1, 1, 3)
    p1 = p0 + rng.normal(size=3) * 0.4
    x, phi = closest_points_segment_boxes(lb, ub, p0, p1)
    for j in range(40):
        f = lambda v: np.sum((p0 + v[3] * (p1 - p0) - v[:3]) ** 2)   # noqa: E731
        res = minimize(f, np.concatenate((np.clip(p0, lb[j], ub[j]), [0.5])),
                       bounds=list(zip(lb[j], ub[j])) + [(0, 1)], method="L-BFGS-B",
                       options={"ftol": 1e-15, "gtol": 1e-12})
        d = np.sum((p0 + phi[j] * (p1 - p0) - x[j]) ** 2)
        assert d <= res.fun + 1e-9
    # segment parallel to a face: minimiser not unique, smallest phi is returned (quirk Q9)
    x, phi = closest_points_segment_boxes(np.array([[0.2, -1, -1.0]]), np.array([[0.6, 1, 0.0]]),
                                          np.array([0.0, 0, 0.5]), np.array([1.0, 0, 0.5]))
    assert abs(phi[0] - 0.2) < 1e-12 and np.abs(x[0] - [0.2, 0, 0.0]).max() < 1e-12


def _check_set_invariants(f, A, b, seed, boxes, inflate):
    assert np.all(A @ seed - b < 0), "seed must be strictly inside its set"
    assert np.abs(np.linalg.norm(A, axis=1) - 1).max() < 1e-12
    # every obstacle is cut off: some row has all 8 inflated vertices on its outside (>= -1e-4)
    V = np.stack(f.obs_points_sets)
    vals = np.einsum("mk,nvk->nmv", A[6:], V) - b[6:][None, :, None]
    assert np.all((vals.min(axis=2) >= -1e-4).any(axis=1))


def test_polyhedron_invariants_example_scene():
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    p0 = np.array([0.3, 0.0, 0.7])
    A, b, Q, p = f.find_set_around_point(p0, fixed_mid=False, optimize=False)
    assert np.array_equal(A[:6], np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1.0]]))
    assert np.allclose(b[:6], [1.0, 0.14, 0.38, 1.0, 1.0, 0.0])
    assert np.array_equal(Q, np.diag([1e4] * 3)) and np.array_equal(p, p0)
    _check_set_invariants(f, A, b, p0, boxes, inflate)
    # first pick is the Euclidean-nearest obstacle, its row is the face normal through the clipped point
    lb, ub = boxes[:, :3] - inflate, boxes[:, 3:] + inflate
    d = np.linalg.norm(np.clip(p0, lb, ub) - p0, axis=1)
    assert f.last_picks[0] == int(np.argmin(d))


def test_iris_loop_example_and_random_scene():
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    p0 = np.array([0.3, 0.0, 0.7])
    for fixed_mid in (True, False):
        A, b, Q, p = f.find_set_around_point(p0, fixed_mid=fixed_mid)
        _check_set_invariants(f, A, b, p0 if fixed_mid else p, boxes, inflate)
        assert 1 <= f.last_iters <= 6
        # the returned ellipsoid (semi-axis matrix = q_inv = Q^-1, quirk Q2) is inscribed in the LAST polyhedron
        E = np.linalg.inv(Q)
        assert np.all(np.linalg.norm(A @ np.linalg.cholesky(E), axis=1) <= b - A @ p + 1e-8)
    rng = np.random.default_rng(5)
    boxes = scenes.random_box_scene(150, rng, 0.04, 0.2)
    seeds = scenes.free_points(6, boxes, 0.01, rng)
    f = oracle_finder(boxes, 0.01, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX)
    for s in seeds:
        A, b, Q, p = f.find_set_around_point(s, fixed_mid=True)
        _check_set_invariants(f, A, b, s, boxes, 0.01)


def test_line_set_contains_segment_and_flags_collision():
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    p0 = np.array([0.45, -0.5, 0.4])
    p1 = p0 + np.array([0.0, 0.0, 0.05])
    A, b, coll = f.find_set_collision_avoidance(p0, p1)
    assert not coll
    for t in np.linspace(0, 1, 5):
        assert np.all(A @ (p0 + t * (p1 - p0)) - b <= 1e-12)
    A2, b2, _ = f.find_set_collision_avoidance(p0, p1, limit_space=True, e_max=0.7)
    assert np.allclose(b2[:6], [p0[0] + 0.7, -p0[0] + 0.7, p0[1] + 0.7, -p0[1] + 0.7, p0[2] + 0.7, -p0[2] + 0.7])
    # a segment that pierces an obstacle sets the collision flag (ConvexSetFinder.py:336-338)
    _, _, coll = f.find_set_collision_avoidance(np.array([0.6, 0.0, 0.3]), np.array([0.6, 0.0, -0.05]))
    assert coll


def test_row_cap_raises_like_reference():
    import pytest

    rng = np.random.default_rng(9)
    boxes = scenes.random_box_scene(4000, rng, 0.01, 0.03)
    seeds = scenes.free_points(40, boxes, 0.0, rng)
    f = oracle_finder(boxes, 0.0, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX, max_rows=20)
    raised = 0
    for s in seeds:
        try:
            f.find_set_around_point(s, fixed_mid=True)
        except ValueError:
            raised += 1
        except RuntimeError:
            pass
    assert raised >= 1          # quirk Q5: > 20 rows overflows the reference's MVIE buffers
    with pytest.raises(ValueError):
        f.mvie_socp(np.vstack([BOX] * 4), np.ones(24))
