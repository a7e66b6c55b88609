"""GPU: K11-K13 -- the planner loop's own tests batched over queries: sample rejection
(BoundPlanner.py:459-478), duplicate-set distance (:505-512), shortest path over the intersection graph (:434,
against networkx, the reference's own call)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _c3_query_sets(geo, scenes, qid, n_sets):
    """Scene of C3 query qid + a few reduced sets built around free points (what graph.nodes holds)."""
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(qid)
    rng = np.random.default_rng(50 + qid)
    if n_sets == 0:
        return ob, infl, wmin, wmax, [], np.zeros((0, 3, 3)), np.zeros((0, 3))
    sc = geo.Scene(ob, infl)
    seeds = scenes.free_points(n_sets, ob, infl + 0.02, rng)
    out = geo.build_sets_point(sc, seeds, wmin, wmax, fixed_mid=True, optimize=True)
    ok = out.status.cpu().numpy() == 0
    Ar, br, mr, _, _ = geo.reduce_ineqs(out.A, out.b, out.m)
    A, b, m = Ar.cpu().numpy(), br.cpu().numpy(), mr.cpu().numpy()
    sets = [[A[i, : m[i]].copy(), b[i, : m[i]].copy()] for i in range(n_sets) if ok[i]]
    return ob, infl, wmin, wmax, sets, out.q_ellipse.cpu().numpy()[ok], out.p_mid.cpu().numpy()[ok]


def test_sample_filter_matches_reference_loop():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes
    from boundplanner_b200.planner import obstacle_sets
    from boundplanner_b200.set_graph import pack_sets
    from oracle.planner_graph import first_free_sample, sample_flags

    Q, C = 6, 48
    obs_list, all_sets, set_off, cands, refs = [], [], [0], [], []
    for q in range(Q):
        ob, infl, wmin, wmax, sets, _, _ = _c3_query_sets(geo, scenes, q, 3 * q)      # query 0 has no known set
        obs_list.append(ob)
        all_sets += sets
        set_off.append(len(all_sets))
        rng = np.random.default_rng(900 + q)
        cand = rng.uniform(wmin, wmax, (C, 3))                   # the stream of C successive rng.uniform(.., 3) draws
        cands.append(cand)
        obs_sets = obstacle_sets(ob, infl)
        refs.append(([sample_flags(obs_sets, sets, c) for c in cand], first_free_sample(obs_sets, sets, cand)))
    scene = geo.SceneBatch(obs_list, 0.01)
    A, b, m = pack_sets(all_sets) if all_sets else (np.zeros((1, 6, 3)), np.zeros((1, 6)), np.zeros(1, np.int32))
    first, flags = geo.sample_filter(scene, np.array(cands), A, b, m, np.array(set_off, np.int32),
                                     item_scene=np.arange(Q, dtype=np.int32), want_flags=True)
    first2 = geo.sample_filter(scene, np.array(cands), A, b, m, np.array(set_off, np.int32),
                               item_scene=np.arange(Q, dtype=np.int32))
    first, flags, first2 = first.cpu().numpy(), flags.cpu().numpy(), first2.cpu().numpy()
    n_coll = n_safe = 0
    for q in range(Q):
        fl, k = refs[q]
        assert first[q] == k and first2[q] == k
        for c in range(C):
            assert bool(flags[q, c] & 1) == fl[c][0], (q, c)
            if not fl[c][0]:                                     # the reference loop skips nothing: both tests run
                assert bool(flags[q, c] & 2) == fl[c][1], (q, c)
            n_coll += fl[c][0]
            n_safe += fl[c][1]
    assert n_coll > 5 and n_safe > 5                             # both rejection reasons occur
    # single scene, no known sets, nothing free: every candidate inside one big obstacle
    big = geo.Scene(np.array([[-2.0, -2, -1, 2, 2, 2]]), 0.0)
    none = geo.sample_filter(big, np.zeros((1, 8, 3)))
    assert int(none.item()) == -1


def test_dedupe_distance_matches_reference_loop():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes
    from oracle.planner_graph import dedupe_distance

    rng = np.random.default_rng(3)
    _, _, _, _, _, Qs, Ps = _c3_query_sets(geo, scenes, 1, 12)
    n = Qs.shape[0]
    node_off, qn, pn, q_new, p_new, ref = [0], [], [], [], [], []
    for i in range(8):
        k = int(rng.integers(0, n))                              # item 0..: k nodes (k = 0: no node -> inf)
        if i == 0:
            k = 0
        idx = rng.permutation(n)[:k]
        qn += [Qs[j] for j in idx]
        pn += [Ps[j] for j in idx]
        node_off.append(len(qn))
        j = int(rng.integers(0, n))
        dq = 0.0 if i % 3 == 1 and k else 1.0                   # some exact duplicates of a node
        base = idx[0] if (i % 3 == 1 and k) else j
        q_new.append(Qs[base] + dq * 1e-3 * rng.normal(size=(3, 3)))
        p_new.append(Ps[base] + dq * 1e-3 * rng.normal(size=3))
        ref.append(dedupe_distance(q_new[-1], p_new[-1], [(Qs[t], Ps[t]) for t in idx]))
    qn = np.array(qn).reshape(-1, 9) if qn else np.zeros((1, 9))
    pn = np.array(pn).reshape(-1, 3) if pn else np.zeros((1, 3))
    dmin, arg = geo.dedupe_distance(np.array(q_new), np.array(p_new), qn, pn, np.array(node_off, np.int32))
    dmin, arg = dmin.cpu().numpy(), arg.cpu().numpy()
    assert np.isinf(dmin[0]) and arg[0] == -1
    for i in range(1, 8):
        if np.isinf(ref[i]):
            assert np.isinf(dmin[i])
        else:
            assert abs(dmin[i] - ref[i]) <= 1e-12 * max(1.0, ref[i])
            assert (dmin[i] > 0.01) == (ref[i] > 0.01)           # the decision the planner takes (:511)


def test_shortest_path_matches_networkx():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo
    from oracle.planner_graph import shortest_path

    rng = np.random.default_rng(8)
    graphs = []
    for g in range(24):
        n = int(rng.integers(2, 60))
        p_edge = rng.uniform(0.05, 0.4)
        edges = [(u, v, float(rng.uniform(0.01, 1.5))) for u in range(n) for v in range(u + 1, n)
                 if rng.uniform() < p_edge]
        graphs.append((n, edges))
    graphs.append((2, []))                                        # start and end not connected
    graphs.append((3, [(0, 2, 0.4), (2, 1, 0.3), (0, 1, 0.9)]))   # the detour is cheaper than the direct edge
    node_off, edge_off, edge_dst, edge_w = [0], [0], [], []
    for n, edges in graphs:
        adj = [[] for _ in range(n)]
        for u, v, w in edges:
            adj[u].append((v, w))
            adj[v].append((u, w))
        for u in range(n):
            edge_dst += [v for v, _ in adj[u]]
            edge_w += [w for _, w in adj[u]]
            edge_off.append(len(edge_dst))
        node_off.append(node_off[-1] + n)
    G = len(graphs)
    path, plen, cost = geo.shortest_paths(np.array(node_off, np.int32), np.array(edge_off, np.int32),
                                          np.array(edge_dst if edge_dst else [0], np.int32),
                                          np.array(edge_w if edge_w else [0.0]), np.zeros(G, np.int32),
                                          np.ones(G, np.int32), max_len=64)
    path, plen, cost = path.cpu().numpy(), plen.cpu().numpy(), cost.cpu().numpy()
    n_paths = 0
    for g, (n, edges) in enumerate(graphs):
        ref_path, ref_cost = shortest_path(n, edges, 0, 1)
        if ref_path is None:
            assert plen[g] == -1 and np.isinf(cost[g])
            continue
        assert plen[g] == len(ref_path)
        assert list(path[g, : plen[g]]) == ref_path               # same node sequence (weights are generic: no ties)
        assert abs(cost[g] - ref_cost) <= 1e-12
        n_paths += 1
    assert n_paths >= 15
    assert list(path[G - 1, :3]) == [0, 2, 1]
