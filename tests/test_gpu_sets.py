"""GPU parity: set construction kernels (K1-K5) through the C ABI vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import RTOL, assert_rows_close, oracle_finder  # noqa: E402


@pytest.fixture(scope="module")
def geo():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from boundplanner_b200 import geometry

    return geometry


@pytest.fixture(scope="module")
def scene_small():
    from boundplanner_b200 import scenes

    rng = np.random.default_rng(11)
    boxes = scenes.random_box_scene(120, rng, 0.04, 0.2)
    inflate = 0.01
    seeds = scenes.free_points(24, boxes, inflate, rng)
    return boxes, inflate, seeds, scenes.WORKSPACE_MIN.copy(), scenes.WORKSPACE_MAX.copy()


def test_closest_points_match_oracle(geo, scene_small):
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    rng = np.random.default_rng(5)
    E = []
    for s in range(seeds.shape[0]):
        L = np.tril(rng.normal(size=(3, 3))) * 0.05
        L[np.diag_indices(3)] = rng.uniform(0.02, 0.4, 3)
        E.append(L @ L.T if s else np.diag([1e-4] * 3))
    E = np.array(E)
    y, dist = geo.closest_points(sc, seeds, E)
    y, dist = y.cpu().numpy(), dist.cpu().numpy()
    for s in range(seeds.shape[0]):
        yo = f.compute_set_projs(f.obs_sets, seeds[s], E[s])
        Q = np.linalg.inv(E[s])
        do = np.linalg.norm(Q @ (yo - seeds[s]).T, axis=0)
        assert np.abs(y[s] - yo).max() <= 1e-7
        assert np.abs(dist[s] - do).max() <= 1e-7 * max(1.0, do.max())


def test_closest_points_line_match_oracle(geo, scene_small):
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    rng = np.random.default_rng(6)
    p0 = seeds[:8]
    p1 = p0 + rng.normal(size=p0.shape) * 0.2
    p1[3] = p0[3] + np.array([0.3, 0.0, 0.0])          # axis-aligned segment (quirk Q9)
    x, phi = geo.closest_points_line(sc, p0, p1)
    x, phi = x.cpu().numpy(), phi.cpu().numpy()
    for s in range(p0.shape[0]):
        xo, phio = f.compute_set_projs_line(f.obs_sets, p0[s], p1[s])
        pc = p0[s] + phi[s][:, None] * (p1[s] - p0[s])
        pco = p0[s] + phio[:, None] * (p1[s] - p0[s])
        # the difference vector and the distance are unique even where (x, phi) is not
        assert np.abs((x[s] - pc) - (xo - pco)).max() <= 1e-12
        assert np.abs(x[s] - xo).max() <= 1e-9
        assert np.abs(phi[s] - phio).max() <= 1e-9


def test_first_pass_polyhedron_bit_exact_picks(geo, scene_small):
    """optimize=False: one compute_polyhedron pass with the initial sphere.  The
    picked-obstacle sequence is index work -> must match exactly."""
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=False)
    sets = out.to_sets()
    assert out.status.cpu().numpy().tolist() == [0] * seeds.shape[0]
    for s in range(seeds.shape[0]):
        Ao, bo, Qo, po = f.find_set_around_point(seeds[s], fixed_mid=True, optimize=False)
        assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"seed {s}")
        assert np.allclose(out.q_ellipse[s].cpu().numpy(), Qo)
        assert np.allclose(out.p_mid[s].cpu().numpy(), po)


def test_mvie_matches_oracle(geo):
    from oracle import mvie as omvie

    rng = np.random.default_rng(3)
    box = np.vstack((np.eye(3), -np.eye(3)))
    S, m_max = 40, 32
    A = np.zeros((S, m_max, 3))
    b = np.full((S, m_max), 10.0)
    m = np.zeros(S, np.int32)
    c = np.zeros((S, 3))
    for s in range(S):
        while True:
            k = rng.integers(3, 20)
            cs = rng.uniform(-0.5, 0.5, 3)
            An = rng.normal(size=(k, 3))
            An /= np.linalg.norm(An, axis=1)[:, None]
            As = np.vstack((box, An))
            bs = np.concatenate((np.array([1, 1, 1.2, 1, 1, 0.0]), An @ cs + rng.uniform(0.005, 0.4, k)))
            if np.min(bs - As @ cs) > 1e-3:
                break
        A[s, : 6 + k], b[s, : 6 + k], m[s], c[s] = As, bs, 6 + k, cs
    for free in (False, True):
        q_inv, q_ell, cen, status, its = geo.mvie(A, b, m, c, free)
        q_inv, q_ell, cen = q_inv.cpu().numpy(), q_ell.cpu().numpy(), cen.cpu().numpy()
        assert status.cpu().numpy().tolist() == [0] * S
        for s in range(S):
            if free:
                Eo, co = omvie.mvie_free(A[s, : m[s]], b[s, : m[s]], p_hint=c[s])
            else:
                Eo, co = omvie.mvie_fixed_mid(A[s, : m[s]], b[s, : m[s]], c[s])
            assert np.abs(q_inv[s] - Eo).max() <= RTOL * np.abs(Eo).max()
            assert np.abs(cen[s] - co).max() <= RTOL
            assert np.abs(q_ell[s] @ q_inv[s] - np.eye(3)).max() <= 1e-8


def test_mvie_box_known_answer(geo):
    """MVIE of an axis-aligned box is diag(half widths) at the box centre."""
    box = np.vstack((np.eye(3), -np.eye(3)))
    A = np.zeros((1, 8, 3))
    b = np.full((1, 8), 10.0)
    A[0, :6] = box
    b[0, :6] = [0.2, 1.0, 1.5, 0.4, 0.5, 0.3]
    m = np.array([6], np.int32)
    q_inv, _, cen, status, _ = geo.mvie(A, b, m, np.zeros((1, 3)), True)
    h = np.array([0.3, 0.75, 0.9])
    assert status.item() == 0
    assert np.abs(q_inv[0].cpu().numpy() - np.diag(h**2)).max() < 1e-9
    assert np.abs(cen[0].cpu().numpy() - np.array([-0.1, 0.25, 0.6])).max() < 1e-9


@pytest.mark.parametrize("fixed_mid", [True, False])
def test_iris_loop_matches_oracle(geo, scene_small, fixed_mid):
    """find_set_around_point end to end: same number of loop iterations, same
    picked rows, q_ellipse / p_mid within tolerance."""
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=fixed_mid, optimize=True)
    sets = out.to_sets()
    status = out.status.cpu().numpy()
    iters = out.iters.cpu().numpy()
    Q = out.q_ellipse.cpu().numpy()
    P = out.p_mid.cpu().numpy()
    for s in range(seeds.shape[0]):
        try:
            Ao, bo, Qo, po = f.find_set_around_point(seeds[s], fixed_mid=fixed_mid, optimize=True)
        except RuntimeError:
            assert status[s] == 1, f"seed {s}: oracle raised 'Ellipse violates constraints'"
            continue
        assert status[s] == 0, f"seed {s}: status {status[s]}"
        assert iters[s] == f.last_iters, f"seed {s}: loop iterations {iters[s]} vs {f.last_iters}"
        assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"seed {s}")
        assert np.abs(Q[s] - Qo).max() <= 1e-5 * np.abs(Qo).max(), f"seed {s}"
        assert np.abs(P[s] - po).max() <= RTOL, f"seed {s}"


def test_line_sets_match_oracle(geo, scene_small):
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    rng = np.random.default_rng(8)
    p0 = seeds[:10]
    p1 = p0 + rng.normal(size=p0.shape) * 0.05
    for limit_space in (False, True):
        out = geo.build_sets_line(sc, p0, p1, ws_min, ws_max, compute_ellipsoid=False, limit_space=limit_space,
                                  e_max=0.7)
        sets = out.to_sets()
        coll = out.collision.cpu().numpy()
        for s in range(p0.shape[0]):
            Ao, bo, co = f.find_set_collision_avoidance(p0[s], p1[s], False, limit_space, 0.7)
            assert bool(coll[s]) == bool(co)
            assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"segment {s}")


def test_line_sets_with_ellipsoid(geo, scene_small):
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    p0 = seeds[:6]
    p1 = p0 + np.array([0.0, 0.0, 0.03])
    out = geo.build_sets_line(sc, p0, p1, ws_min, ws_max, compute_ellipsoid=True)
    sets = out.to_sets()
    Q = out.q_ellipse.cpu().numpy()
    P = out.p_mid.cpu().numpy()
    for s in range(p0.shape[0]):
        if out.collision[s].item():
            continue
        Ao, bo, Qo, po, co = f.find_set_collision_avoidance(p0[s], p1[s], True)
        assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"segment {s}")
        assert np.abs(Q[s] - Qo).max() <= 1e-5 * np.abs(Qo).max()
        assert np.abs(P[s] - po).max() <= RTOL


def test_empty_inputs(geo, scene_small):
    boxes, inflate, seeds, ws_min, ws_max = scene_small
    sc = geo.Scene(boxes, inflate)
    out = geo.build_sets_point(sc, np.zeros((0, 3)), ws_min, ws_max)
    assert out.A.shape[0] == 0
    empty = geo.Scene(np.zeros((0, 6)), 0.0)
    out = geo.build_sets_point(empty, seeds[:3], ws_min, ws_max, fixed_mid=True)
    # no obstacles: the set is the workspace box and its MVIE the inscribed axis-aligned ellipsoid
    assert out.m.cpu().numpy().tolist() == [6, 6, 6]
    h = 0.5 * (ws_max - ws_min)
    Q = out.q_ellipse.cpu().numpy()
    assert np.abs(np.linalg.inv(Q[0]) - np.diag(h**2)).max() < 1e-8
