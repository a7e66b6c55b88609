"""CPU: synthetic scenes, sharding arithmetic and packing helpers (host logic)."""
import numpy as np

from boundplanner_b200 import distributed as bpd
from boundplanner_b200 import scenes


def test_scenes_are_deterministic_and_collision_free():
    b1, infl, s1, wmin, wmax = scenes.config_c2(200, 32)
    b2, _, s2, _, _ = scenes.config_c2(200, 32)
    assert np.array_equal(b1, b2) and np.array_equal(s1, s2)
    assert b1.shape == (200, 6) and np.all(b1[:, 3:] > b1[:, :3])
    assert not scenes.in_collision(s1, b1, infl).any()
    assert np.all(s1 >= wmin) and np.all(s1 <= wmax)
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    assert boxes.shape == (12, 6) and inflate == 0.08
    c4 = scenes.shelf_scene(300, np.random.default_rng(2))
    assert c4.shape == (300, 6) and np.all(c4[:, 3:] > c4[:, :3])


def test_shard_range_and_row_blocks():
    for n, w in ((256, 8), (10, 3), (7, 8), (0, 4)):
        spans = [bpd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    for S, w in ((2048, 8), (256, 2), (33, 4), (5, 8)):
        blocks = bpd.balanced_row_blocks(S, w)
        assert len(blocks) == w and blocks[0][0] == 0 and blocks[-1][1] == S
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
        if S >= 256:
            work = [sum(S - 1 - i for i in range(lo, hi)) for lo, hi in blocks]
            assert max(work) <= 1.1 * (S * (S - 1) / 2) / w + S


def test_pack_sets_padding():
    from boundplanner_b200.set_graph import pack_sets

    sets = [[np.eye(3), np.ones(3)], [np.vstack((np.eye(3), -np.eye(3))), np.arange(6.0)]]
    A, b, m = pack_sets(sets)
    assert A.shape == (2, 6, 3) and m.tolist() == [3, 6]
    assert np.all(A[0, 3:] == 0) and np.all(b[0, 3:] == 10.0)      # normalize_set_size padding


def test_planner_vectorised_host_tests_equal_reference_loops():
    """SetSequencePlanner._in_safe / _min_node_distance / _in_collision (vectorised) against the reference's
    per-node loops (BoundPlanner.py:467-476, :505-510) as restated in oracle.planner_graph."""
    from boundplanner_b200 import scenes
    from boundplanner_b200.planner import SetSequencePlanner, obstacle_sets
    from oracle.planner_graph import dedupe_distance, sample_flags

    rng = np.random.default_rng(3)
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(5)
    pl = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=object(), rng=rng)
    pl._nodes_reset()
    obs_sets = obstacle_sets(ob, infl)
    box = np.vstack((np.eye(3), -np.eye(3)))
    nodes, qs = [], []
    for k in range(7):                                         # ragged row counts
        c = rng.uniform(wmin + 0.2, wmax - 0.2)
        half = rng.uniform(0.05, 0.4, 3)
        extra = rng.normal(size=(int(rng.integers(0, 6)), 3))
        extra /= np.linalg.norm(extra, axis=1)[:, None]
        a_set = np.vstack((box, extra))
        b_set = np.concatenate((c + half, -(c - half), extra @ c + rng.uniform(0.05, 0.3, extra.shape[0])))
        q = np.diag(rng.uniform(1, 50, 3))
        nodes.append([a_set, b_set])
        qs.append((q, c))
        pl._nodes_add(a_set, b_set, q, c)
    n_safe = n_coll = 0
    for _ in range(400):
        x = rng.uniform(wmin, wmax, 3)
        coll, safe = sample_flags(obs_sets, nodes, x)
        assert pl._in_collision(x) == coll
        # the reference's loop breaks at the first obstacle / set; both tests are evaluated independently there
        assert pl._in_safe(x) == safe
        n_safe += safe
        n_coll += coll
    assert n_safe > 10 and n_coll > 10
    for _ in range(50):
        q = np.diag(rng.uniform(1, 50, 3)) + 1e-3 * rng.normal(size=(3, 3))
        p = rng.uniform(wmin, wmax, 3)
        assert abs(pl._min_node_distance(q, p) - dedupe_distance(q, p, qs)) < 1e-12
    q, p = qs[3]
    assert pl._min_node_distance(q, p) == 0.0
