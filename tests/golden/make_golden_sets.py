"""Generates the BASELINE-size golden fixtures (committed, compressed):

  python tests/golden/make_golden_sets.py [c2] [c4] [c3]

c2_graph_golden.npz  config C2 of BASELINE.json in full: the ORACLE's convex set (find_set_around_point,
                     fixed_mid=True, optimize=True; ConvexSetFinder.py:190-240) for ALL 256 seeds -- rows, row
                     counts, loop iteration counts, final ellipsoid, status -- and the full 32 640-bit set-graph
                     adjacency from the reference's own scipy.optimize.linprog call (BoundPlanner.py:774-798, tol
                     0.01) on those sets, plus the exact margin s* of every pair HiGHS answers within 1e-5 of the
                     decision boundary (near ties, see oracle/set_graph.py).
c4_sets_golden.npz   config C4: the oracle's sets for 258 of the 2048 seeds (every 8th + seeds 700, 2047) and the
                     adjacency among them (33 153 pairs).
Source of every number: the oracle (oracle/convex_set_finder.py, a restatement of the reference whose QP / SOCP
solvers are not installable here) and, for the adjacency, the reference's verbatim HiGHS call.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from boundplanner_b200 import scenes  # noqa: E402

STATUS = {"ok": 0, "RuntimeError": 1, "ValueError": 5}   # BP_OK, BP_ELLIPSE_VIOLATION, BP_ROW_CAP


def _build(args):
    boxes, inflate, ws_min, ws_max, seeds, max_rows = args
    from oracle.convex_set_finder import ConvexSetFinder
    from oracle.obstacles import obstacle_reps

    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    f = ConvexSetFinder(obs_sets, pts, ws_max, ws_min, max_rows=max_rows)
    out = []
    for p in seeds:
        try:
            A, b, Q, c = f.find_set_around_point(p, fixed_mid=True, optimize=True)
            out.append((0, A, b, Q, c, f.last_iters))
        except (RuntimeError, ValueError) as e:
            out.append((STATUS[type(e).__name__], np.zeros((0, 3)), np.zeros(0), np.zeros((3, 3)), np.zeros(3),
                        f.last_iters))
    return out


def _pairs(args):
    sets, pairs, tol = args
    from oracle.set_graph import intersection_margin, set_intersection

    res = []
    for i, j in pairs:
        ok = bool(set_intersection(sets[i], sets[j], tol)[2])
        res.append(ok)
    return res


def _margins(args):
    sets, pairs, tol = args
    from oracle.set_graph import intersection_margin

    return [intersection_margin(sets[i], sets[j], tol) for i, j in pairs]


def build_all(boxes, inflate, ws_min, ws_max, seeds, cores, max_rows=None):
    chunks = np.array_split(np.arange(len(seeds)), cores * 4)
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_build, [(boxes, inflate, ws_min, ws_max, seeds[c], max_rows) for c in chunks if len(c)])
    return [r for p in parts for r in p]


def graph(sets, ok, tol, cores):
    idx = [i for i in range(len(sets)) if ok[i]]
    pairs = [(i, j) for a, i in enumerate(idx) for j in idx[a + 1:]]
    chunks = [pairs[k::cores * 4] for k in range(cores * 4)]
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_pairs, [(sets, c, tol) for c in chunks])
    adj = np.zeros((len(sets), len(sets)), dtype=bool)
    for c, r in zip(chunks, parts):
        for (i, j), v in zip(c, r):
            adj[i, j] = v
    return adj, pairs


def pack(results):
    m = np.array([r[1].shape[0] for r in results], dtype=np.int32)
    rows = np.concatenate([np.hstack((r[1], r[2][:, None])) for r in results]) if m.sum() else np.zeros((0, 4))
    return dict(status=np.array([r[0] for r in results], dtype=np.int32), m=m, rows=rows,
                q_ellipse=np.array([r[3] for r in results]), p_mid=np.array([r[4] for r in results]),
                iters=np.array([r[5] for r in results], dtype=np.int32))


def near_tie_margins(sets, pairs, adj, tol, cores, band=1e-5):
    """exact margins of the pairs a cheap bound cannot place safely on one side: run the margin LP for every pair
    (4-variable LP, same cost as the feasibility call) and keep those with |s*| < band."""
    chunks = [pairs[k::cores * 4] for k in range(cores * 4)]
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_margins, [(sets, c, tol) for c in chunks])
    near, bad = [], []
    for c, r in zip(chunks, parts):
        for (i, j), s in zip(c, r):
            if abs(s) < band:
                near.append((i, j, s))
            if (s <= 0) != bool(adj[i, j]) and abs(s) >= band:
                bad.append((i, j, s))
    assert not bad, f"HiGHS feasibility answer contradicts the exact margin: {bad[:5]}"
    return np.array(near, dtype=np.float64).reshape(-1, 3)


def main():
    which = set(sys.argv[1:]) or {"c2", "c4"}
    cores = os.cpu_count() or 1
    if "c2" in which:
        t0 = time.time()
        boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
        res = build_all(boxes, inflate, ws_min, ws_max, seeds, cores)
        d = pack(res)
        sets = [[r[1], r[2]] for r in res]
        adj, pairs = graph(sets, d["status"] == 0, 0.01, cores)
        near = near_tie_margins(sets, pairs, adj, 0.01, cores)
        np.savez_compressed(os.path.join(HERE, "c2_graph_golden.npz"), seeds=seeds, n_obs=boxes.shape[0],
                            adj_bits=np.packbits(adj, axis=1), near_ties=near, tol=0.01, **d)
        print(f"c2: {len(res)} sets ({(d['status'] == 0).sum()} ok), {len(pairs)} pairs, {int(adj.sum())} edges, "
              f"{near.shape[0]} near ties, {time.time() - t0:.0f} s")
    if "c4" in which:
        t0 = time.time()
        boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
        sel = np.unique(np.concatenate((np.arange(0, 2048, 8), [700, 2047])))
        res = build_all(boxes, inflate, ws_min, ws_max, seeds[sel], cores)
        d = pack(res)
        sets = [[r[1], r[2]] for r in res]
        adj, pairs = graph(sets, d["status"] == 0, 0.01, cores)
        near = near_tie_margins(sets, pairs, adj, 0.01, cores)
        np.savez_compressed(os.path.join(HERE, "c4_sets_golden.npz"), seed_index=sel, seeds=seeds[sel],
                            n_obs=boxes.shape[0], adj_bits=np.packbits(adj, axis=1), near_ties=near, tol=0.01, **d)
        print(f"c4: {len(res)} sets ({(d['status'] == 0).sum()} ok), {len(pairs)} pairs, {int(adj.sum())} edges, "
              f"{near.shape[0]} near ties, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
