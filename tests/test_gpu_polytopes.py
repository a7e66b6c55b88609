"""GPU: general convex polytope obstacles (obs_sets with up to 15 rows + their vertices, as
ConvexSetFinder receives them) in the point-set path -- closest points, polyhedron pass, IRIS loop -- against the
oracle, and the segment path (closest points to a segment, find_set_collision_avoidance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import RTOL, assert_rows_close  # noqa: E402


@pytest.fixture(scope="module")
def poly_scene():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes
    from oracle.convex_set_finder import ConvexSetFinder as OracleFinder

    rng = np.random.default_rng(31)
    obs_sets, obs_points = scenes.random_polytope_scene(300, rng, 0.04, 0.16)
    ws_min, ws_max = scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX
    seeds = scenes.polytope_free_points(14, obs_sets, 0.03, rng)
    scene = geo.PolytopeScene(obs_sets, obs_points)
    ora = OracleFinder(obs_sets, obs_points, list(ws_max), list(ws_min))
    return geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max


def test_polytope_closest_points_match_oracle(poly_scene):
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    rng = np.random.default_rng(1)
    for trial in range(3):
        L = np.tril(rng.normal(size=(3, 3))) * 0.1 + np.diag(rng.uniform(0.05, 0.4, 3))
        E = L @ L.T if trial else 1e-4 * np.eye(3)
        p = seeds[trial]
        y, dist = geo.closest_points(scene, p[None], E[None])
        y, dist = y[0].cpu().numpy(), dist[0].cpu().numpy()
        yo = ora.compute_set_projs(obs_sets, p, E)
        do = np.linalg.norm(np.linalg.solve(E, (yo - p).T), axis=0)
        assert np.abs(y - yo).max() < 1e-7
        assert (np.abs(dist - do) / do).max() < 1e-7


def test_polytope_iris_loop_matches_oracle(poly_scene):
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    out = geo.build_sets_point(scene, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    one = geo.build_sets_point(scene, seeds, ws_min, ws_max, fixed_mid=True, optimize=False)
    status, m, iters = out.status.cpu().numpy(), out.m.cpu().numpy(), out.iters.cpu().numpy()
    A, b, Q, P = out.A.cpu().numpy(), out.b.cpu().numpy(), out.q_ellipse.cpu().numpy(), out.p_mid.cpu().numpy()
    A1, b1, m1 = one.A.cpu().numpy(), one.b.cpu().numpy(), one.m.cpu().numpy()
    n_ok = 0
    for s in range(len(seeds)):
        # first pass alone (compute_polyhedron around the initial sphere)
        a_o, b_o, _, _ = ora.find_set_around_point(seeds[s], fixed_mid=True, optimize=False)
        assert_rows_close(A1[s, : m1[s]], b1[s, : m1[s]], a_o, b_o, f"seed {s}, first pass")
        try:
            a_o, b_o, q_o, p_o = ora.find_set_around_point(seeds[s], fixed_mid=True, optimize=True)
        except (RuntimeError, ValueError):
            assert status[s] != 0
            continue
        assert status[s] == 0 and iters[s] == ora.last_iters, (s, status[s], iters[s], ora.last_iters)
        assert_rows_close(A[s, : m[s]], b[s, : m[s]], a_o, b_o, f"seed {s}")
        assert np.abs(Q[s] - q_o).max() <= 1e-5 * np.abs(q_o).max()
        assert np.abs(P[s] - p_o).max() <= 1e-5
        n_ok += 1
    assert n_ok >= 10


def test_boxes_as_polytopes_give_the_box_answer(poly_scene):
    """The same box obstacles through the polytope path (rows + 8 corners) and through the box path."""
    geo, *_ = poly_scene
    from boundplanner_b200 import scenes
    from oracle.obstacles import obstacle_reps

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(400, 16)
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    sp = geo.PolytopeScene(obs_sets, pts)
    sb = geo.Scene(boxes, inflate)
    op = geo.build_sets_point(sp, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    ob = geo.build_sets_point(sb, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    assert np.array_equal(op.m.cpu().numpy(), ob.m.cpu().numpy())
    assert np.array_equal(op.iters.cpu().numpy(), ob.iters.cpu().numpy())
    assert np.abs(op.A.cpu().numpy() - ob.A.cpu().numpy()).max() < 1e-9
    assert np.abs(op.b.cpu().numpy() - ob.b.cpu().numpy()).max() < 1e-9
    qp, qb = op.q_ellipse.cpu().numpy(), ob.q_ellipse.cpu().numpy()
    assert np.abs(qp - qb).max() <= 1e-7 * np.abs(qb).max()


def test_dropin_with_polytope_obstacles(poly_scene):
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    import boundplanner_b200 as bp

    gpu = bp.ConvexSetFinder(obs_sets, obs_points, list(ws_max), list(ws_min))
    A, b, Q, p = gpu.find_set_around_point(seeds[0], fixed_mid=True)
    Ao, bo, Qo, po = ora.find_set_around_point(seeds[0], fixed_mid=True)
    assert_rows_close(A, b, Ao, bo, "drop-in polytope set")
    assert np.abs(Q - Qo).max() <= 1e-5 * np.abs(Qo).max() and np.abs(p - po).max() <= RTOL
    a_init, b_init = gpu.init_halfspaces()
    q_inv = np.diag([1e-4] * 3)
    rows = gpu.compute_polyhedron(q_inv, np.linalg.inv(q_inv), seeds[1], a_init, b_init)
    rows_o = ora.compute_polyhedron(q_inv, np.linalg.inv(q_inv), seeds[1], a_init, b_init)
    assert_rows_close(np.array(rows[0]), np.array(rows[1]), np.array(rows_o[0]), np.array(rows_o[1]), "compute_polyhedron")
    # BoundMPC's call (BoundMPC.py:486-488) and the planner's (BoundPlanner.py:387-389) on polytope obstacles
    for k, kw in enumerate((dict(limit_space=True, e_max=0.7), dict(compute_ellipsoid=True))):
        p0, p1 = seeds[2 + k], seeds[2 + k] + np.array([0.03, -0.02, 0.04])
        got = gpu.find_set_collision_avoidance(p0, p1, **kw)
        want = ora.find_set_collision_avoidance(p0, p1, **kw)
        assert got[-1] == want[-1]
        assert_rows_close(got[0], got[1], want[0], want[1], "line set over polytopes")
        if "compute_ellipsoid" in kw:
            assert np.abs(got[2] - want[2]).max() <= 1e-5 * np.abs(want[2]).max()
            assert np.abs(got[3] - want[3]).max() <= 1e-5


def test_polytope_segment_closest_points_and_line_sets(poly_scene):
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    rng = np.random.default_rng(5)
    p0 = seeds.copy()
    d = rng.normal(size=p0.shape)
    d *= (rng.uniform(0.02, 0.25, (p0.shape[0], 1)) / np.linalg.norm(d, axis=1)[:, None])
    d[1] = [0.0, 0.0, 0.05]                                  # the planner's l_ee (axis-aligned)
    p1 = p0 + d
    x, phi = geo.closest_points_line(scene, p0[:4], p1[:4])
    x, phi = x.cpu().numpy(), phi.cpu().numpy()
    for s in range(4):
        xo, phio = ora.compute_set_projs_line(obs_sets, p0[s], p1[s])
        do = np.linalg.norm(p0[s] + phio[:, None] * d[s] - xo, axis=1)
        dg = np.linalg.norm(p0[s] + phi[s][:, None] * d[s] - x[s], axis=1)
        assert np.abs(dg - do).max() < 1e-9
        assert np.abs(phi[s] - phio).max() < 1e-7 and np.abs(x[s] - xo).max() < 1e-7
    out = geo.build_sets_line(scene, p0, p1, ws_min, ws_max, compute_ellipsoid=False)
    A, b, m = out.A.cpu().numpy(), out.b.cpu().numpy(), out.m.cpu().numpy()
    coll = out.collision.cpu().numpy()
    n_coll = 0
    for s in range(p0.shape[0]):
        a_o, b_o, c_o = ora.find_set_collision_avoidance(p0[s], p1[s])
        assert bool(coll[s]) == bool(c_o)
        n_coll += bool(c_o)
        if not c_o:                                           # touching segments: the fallback normals are not unique
            assert_rows_close(A[s, : m[s]], b[s, : m[s]], a_o, b_o, f"segment {s}")


def test_sample_filter_over_polytope_obstacles(poly_scene):
    """K11 on a polytope scene: the reference's rejection test max(A x - b) < 1e-3 over the obstacle's own rows."""
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    from oracle.planner_graph import first_free_sample, sample_flags

    rng = np.random.default_rng(77)
    cand = rng.uniform(ws_min, ws_max, (2, 400, 3))
    # candidates just outside a face (inside / outside the 1 mm margin) and at obstacle centroids
    A0, b0 = obs_sets[0]
    on_face = obs_points[0][np.abs(obs_points[0] @ A0[0] - b0[0]) < 1e-9]
    face_mid = on_face.mean(axis=0)
    cand[0, 0] = face_mid + 5e-4 * A0[0]
    cand[0, 1] = face_mid + 5e-3 * A0[0]
    for k in range(20):
        cand[1, k] = obs_points[3 * k].mean(axis=0)
    first, flags = geo.sample_filter(scene, cand, want_flags=True)
    first, flags = first.cpu().numpy(), flags.cpu().numpy()
    n_coll = 0
    for q in range(2):
        ref = [sample_flags(obs_sets, [], c)[0] for c in cand[q]]
        assert [bool(f & 1) for f in flags[q]] == ref
        assert first[q] == first_free_sample(obs_sets, [], cand[q])
        n_coll += sum(ref)
    assert n_coll > 10 and bool(flags[0, 0] & 1) and not bool(flags[0, 1] & 1)


def test_polytope_scene_update_in_the_reference_assignment_order(poly_scene):
    """add_obstacle_reps(update=True) assigns set_finder.obs_sets FIRST and obs_points_sets second
    (BoundPlanner.py:150-152).  A polytope update -- same obstacle count with moved obstacles, and a changed
    count -- must use the NEW vertex lists (vertex test of ConvexSetFinder.py:449-451)."""
    import boundplanner_b200 as bp
    from boundplanner_b200 import scenes
    from oracle.convex_set_finder import ConvexSetFinder as OracleFinder

    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    gpu = bp.ConvexSetFinder(obs_sets, obs_points, list(ws_max), list(ws_min), strict_rows=False)
    rng = np.random.default_rng(21)
    for n_new in (len(obs_sets), len(obs_sets) // 2):
        sets2, pts2 = scenes.random_polytope_scene(n_new, rng)
        gpu.obs_sets = sets2                       # reference order: rows first ...
        gpu.obs_points_sets = pts2                 # ... vertices second
        ora2 = OracleFinder(sets2, pts2, list(ws_max), list(ws_min), max_rows=None)
        pts = scenes.polytope_free_points(3, sets2, 0.02, rng)
        for p in pts:
            A, b, Q, c = gpu.find_set_around_point(p, fixed_mid=True)
            Ao, bo, Qo, co = ora2.find_set_around_point(p, fixed_mid=True)
            assert_rows_close(A, b, Ao, bo, f"polytope update to {n_new} obstacles")
    # a polytope scene with a stale vertex list of the wrong length fails loudly at first use
    gpu.obs_sets = obs_sets
    with pytest.raises(ValueError, match="one vertex array per obstacle"):
        gpu.find_set_around_point(seeds[0], fixed_mid=True)


def test_compute_polytope_vertices_on_device(poly_scene):
    """compute_polytope_vertices (util_functions.py:66-79) on the GPU: the same vertex SET as the scene generator's
    qhull enumeration, the reference's ValueError for an unbounded set, and a PolytopeScene built without vertex
    lists gives the same sets as one built with them."""
    import boundplanner_b200 as bp
    from boundplanner_b200 import utils

    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    got = utils.obstacle_points_sets(obs_sets)
    for j, (g, want) in enumerate(zip(got, obs_points)):
        assert g.shape == want.shape, f"obstacle {j}: {g.shape[0]} vertices, expected {want.shape[0]}"
        d = np.linalg.norm(g[:, None, :] - want[None, :, :], axis=2)
        assert d.min(axis=1).max() < 1e-9 and d.min(axis=0).max() < 1e-9
    # a unit cube: 8 vertices, each met by exactly three planes; a cube with a redundant plane through a vertex
    box = np.vstack((np.eye(3), -np.eye(3)))
    v = np.array(bp.compute_polytope_vertices(box, np.array([1, 1, 1, 0, 0, 0.0])))
    assert v.shape == (8, 3) and set(map(tuple, np.round(v, 12))) == {(x, y, z) for x in (0, 1) for y in (0, 1) for z in (0, 1)}
    touch = np.vstack((box, [[1, 1, 1]])) 
    v = np.array(bp.compute_polytope_vertices(touch, np.array([1, 1, 1, 0, 0, 0, 3.0])))
    assert v.shape == (8, 3)                                  # four planes through (1,1,1): still one vertex
    with pytest.raises(ValueError, match="not a polytope"):
        bp.compute_polytope_vertices(box[:5], np.array([1, 1, 1, 0, 0.0]))      # open towards -z
    auto = geo.PolytopeScene(obs_sets)                        # vertices enumerated on the device
    a = geo.build_sets_point(auto, seeds[:6], ws_min, ws_max, fixed_mid=True)
    w = geo.build_sets_point(scene, seeds[:6], ws_min, ws_max, fixed_mid=True)
    assert np.array_equal(a.m.cpu().numpy(), w.m.cpu().numpy())
    assert np.abs(a.A.cpu().numpy() - w.A.cpu().numpy()).max() < 1e-9


def test_large_polytope_scene_without_the_shared_memory_cache(poly_scene):
    """More than 3072 polytope obstacles: the closest points no longer fit in shared memory and the winner of every
    pick is re-solved.  4000 boxes handed over as polytopes (rows + 8 corners) must give the box path's sets, for
    the IRIS loop (fused kernel), one polyhedron pass and the segment sets."""
    geo, *_ = poly_scene
    import torch

    from boundplanner_b200 import scenes
    from oracle.obstacles import obstacle_reps

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(4000, 12, seed=5)
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    sp = geo.PolytopeScene(obs_sets, pts)
    sb = geo.Scene(boxes, inflate)
    assert sp.n == 4000
    for optimize in (True, False):
        op = geo.build_sets_point(sp, seeds, ws_min, ws_max, fixed_mid=True, optimize=optimize)
        ob = geo.build_sets_point(sb, seeds, ws_min, ws_max, fixed_mid=True, optimize=optimize)
        assert np.array_equal(op.status.cpu().numpy(), ob.status.cpu().numpy())
        assert np.array_equal(op.m.cpu().numpy(), ob.m.cpu().numpy())
        assert np.array_equal(op.iters.cpu().numpy(), ob.iters.cpu().numpy())
        assert np.abs(op.A.cpu().numpy() - ob.A.cpu().numpy()).max() < 1e-9
        assert np.abs(op.b.cpu().numpy() - ob.b.cpu().numpy()).max() < 1e-9
    rng = np.random.default_rng(2)
    d = rng.normal(size=seeds.shape)
    d *= 0.04 / np.linalg.norm(d, axis=1)[:, None]
    lp = geo.build_sets_line(sp, seeds, seeds + d, ws_min, ws_max, compute_ellipsoid=True)
    lb = geo.build_sets_line(sb, seeds, seeds + d, ws_min, ws_max, compute_ellipsoid=True)
    assert np.array_equal(lp.m.cpu().numpy(), lb.m.cpu().numpy())
    assert np.array_equal(lp.collision.cpu().numpy(), lb.collision.cpu().numpy())
    assert np.abs(lp.A.cpu().numpy() - lb.A.cpu().numpy()).max() < 1e-9
    assert np.abs(lp.b.cpu().numpy() - lb.b.cpu().numpy()).max() < 1e-9
    # the cached and the re-solving form agree bit for bit: the first 3000 obstacles alone vs the same 3000 plus
    # 1000 far-away ones that never matter
    far = boxes[:1000].copy()
    far[:, :3] += 50.0
    far[:, 3:] += 50.0
    os_small, pts_small, _ = obstacle_reps(boxes[:3000], inflate)
    os_big, pts_big, _ = obstacle_reps(np.vstack((boxes[:3000], far)), inflate)
    a = geo.build_sets_point(geo.PolytopeScene(os_small, pts_small), seeds, ws_min, ws_max, fixed_mid=True)
    b = geo.build_sets_point(geo.PolytopeScene(os_big, pts_big), seeds, ws_min, ws_max, fixed_mid=True)
    assert torch.equal(a.A, b.A) and torch.equal(a.b, b.b) and torch.equal(a.q_ellipse, b.q_ellipse)


def test_polytope_scene_batch_equals_single_scenes(poly_scene):
    """A batch of polytope scenes (one per planning query) gives, seed by seed, the sets of the single scenes."""
    geo, scene, ora, obs_sets, obs_points, seeds, ws_min, ws_max = poly_scene
    import torch

    from boundplanner_b200 import scenes

    parts = [(obs_sets, obs_points)]
    for k, n in ((1, 120), (2, 57)):
        parts.append(scenes.random_polytope_scene(n, np.random.default_rng(40 + k), 0.04, 0.16))
    batch = geo.PolytopeSceneBatch(parts)
    sd, item = [], []
    for k, (os_k, _) in enumerate(parts):
        pts = seeds[:5] if k == 0 else scenes.polytope_free_points(5, os_k, 0.03, np.random.default_rng(50 + k))
        sd.append(pts)
        item += [k] * len(pts)
    sd = np.vstack(sd)
    order = np.random.default_rng(0).permutation(len(item))          # seeds of different scenes interleaved
    sd, item = sd[order], np.asarray(item, np.int32)[order]
    got = geo.build_sets_point(batch, sd, ws_min, ws_max, fixed_mid=True, optimize=True, item_scene=item)
    d = np.tile(np.array([0.03, -0.02, 0.04]), (len(item), 1))
    got_l = geo.build_sets_line(batch, sd, sd + d, ws_min, ws_max, compute_ellipsoid=True, item_scene=item)
    cand = np.random.default_rng(3).uniform(ws_min, ws_max, (len(item), 16, 3))
    got_f = geo.sample_filter(batch, cand, item_scene=item)
    for k, (os_k, op_k) in enumerate(parts):
        single = geo.PolytopeScene(os_k, op_k)
        sel = np.flatnonzero(item == k)
        want = geo.build_sets_point(single, sd[sel], ws_min, ws_max, fixed_mid=True, optimize=True)
        t = torch.as_tensor(sel, device="cuda")
        assert torch.equal(got.A[t], want.A) and torch.equal(got.b[t], want.b) and torch.equal(got.m[t], want.m)
        assert torch.equal(got.q_ellipse[t], want.q_ellipse) and torch.equal(got.status[t], want.status)
        want_l = geo.build_sets_line(single, sd[sel], sd[sel] + d[sel], ws_min, ws_max, compute_ellipsoid=True)
        assert torch.equal(got_l.A[t], want_l.A) and torch.equal(got_l.b[t], want_l.b)
        assert torch.equal(got_l.collision[t], want_l.collision)
        want_f = geo.sample_filter(single, cand[sel])
        assert torch.equal(got_f[t], want_f)
