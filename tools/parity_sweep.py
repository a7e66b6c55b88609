"""Wider parity sweep than the test suite (GPU box): the fused set build and the pair graph against the oracle on
several scene families and RNG seeds.  Prints one line per scene family; exit code 1 on any mismatch.

  python tools/parity_sweep.py [--seeds-per-scene 24]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
from oracle.set_graph import set_intersection
from tests import util as tu
from tests.util import oracle_finder

ap = argparse.ArgumentParser()
ap.add_argument("--seeds-per-scene", type=int, default=24)
args = ap.parse_args()
K = args.seeds_per_scene
bad = 0
families = []
for sd in (11, 12):
    rng = np.random.default_rng(sd)
    boxes = scenes.random_box_scene(1000, rng, 0.02, 0.08)
    families.append((f"C2-like clutter rng {sd}", boxes, 0.01, rng))
rng = np.random.default_rng(13)
families.append(("dense clutter, big boxes", scenes.random_box_scene(400, rng, 0.05, 0.25), 0.01, rng))
rng = np.random.default_rng(14)
families.append(("shelf plates + clutter (C4-like, 2000)", scenes.shelf_scene(2000, rng), 0.0, rng))
rng = np.random.default_rng(15)
families.append(("C3-like, 200 boxes", scenes.random_box_scene(200, rng, 0.05, 0.2), 0.01, rng))
ws_min, ws_max = scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX
for name, boxes, inflate, rng in families:
    seeds = scenes.free_points(K, boxes, inflate + 0.005, rng)
    sc = geo.Scene(boxes, inflate)
    t0 = time.perf_counter()
    res = {}
    for fixed_mid in (True, False):
        out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=fixed_mid, optimize=True)
        res[fixed_mid] = [t.cpu().numpy() for t in (out.A, out.b, out.m, out.q_ellipse, out.p_mid, out.status, out.iters)]
    A, b, m, Q, P, st, it = res[True]
    bits = geo.pair_feasible(torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(), torch.as_tensor(m).cuda(), 0.01)
    adj = geo.unpack_adjacency(bits, K).cpu().numpy()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    n_cmp = n_err = 0
    worst_row = worst_q = 0.0
    for fixed_mid in (True, False):
        A_, b_, m_, Q_, P_, st_, it_ = res[fixed_mid]
        for s in range(K):
            try:
                a_o, b_o, q_o, p_o = f.find_set_around_point(seeds[s], fixed_mid=fixed_mid, optimize=True)
            except (RuntimeError, ValueError):
                n_err += 1
                if st_[s] == 0:
                    print(f"  MISMATCH {name}: seed {s} fixed_mid={fixed_mid}: oracle raises, GPU status 0"); bad += 1
                continue
            n_cmp += 1
            if st_[s] != 0 or m_[s] != len(b_o) or it_[s] != f.last_iters:
                print(f"  MISMATCH {name}: seed {s} fixed_mid={fixed_mid}: status {st_[s]} rows {m_[s]} vs {len(b_o)} "
                      f"iters {it_[s]} vs {f.last_iters}"); bad += 1
                continue
            try:
                tu.assert_rows_close(A_[s, :m_[s]], b_[s, :m_[s]], a_o, b_o, f"{name} seed {s} fixed_mid={fixed_mid}")
            except AssertionError as e:
                print("  MISMATCH", e); bad += 1
                continue
            if not (tu.REORDERED and tu.REORDERED[-1].startswith(f"{name} seed {s} ")):
                worst_row = max(worst_row, np.abs(A_[s, :m_[s]] - a_o).max(), np.abs(b_[s, :m_[s]] - b_o).max())
            worst_q = max(worst_q, np.abs(Q_[s] - q_o).max() / np.abs(q_o).max(), np.abs(P_[s] - p_o).max())
    n_pairs = n_near = 0
    ok_sets = [s for s in range(K) if st[s] == 0]
    for i in ok_sets:
        for j in ok_sets:
            if j <= i:
                continue
            x, _, hit = set_intersection([A[i, :m[i]], b[i, :m[i]]], [A[j, :m[j]], b[j, :m[j]]], 0.01)
            n_pairs += 1
            if bool(hit) != bool(adj[i, j]):
                # HiGHS works to 1e-7: only a disagreement with a clear margin counts
                from oracle.set_graph import intersection_margin
                mg = intersection_margin([A[i, :m[i]], b[i, :m[i]]], [A[j, :m[j]], b[j, :m[j]]], 0.01)
                if abs(mg) > 1e-6:
                    print(f"  MISMATCH {name}: pair {i},{j}: HiGHS {hit} GPU {adj[i, j]} margin {mg:.3e}"); bad += 1
                else:
                    n_near += 1
    print(f"{name}: {n_cmp} sets compared ({n_err} reference errors reproduced), rows/iters exact, "
          f"max |row diff| {worst_row:.1e}, max rel |q diff| {worst_q:.1e}; {n_pairs} pairs vs HiGHS, "
          f"{int(adj.sum())} edges, {n_near} near ties; {time.perf_counter() - t0:.0f} s", flush=True)
print("sets whose picked rows matched up to a permutation (numerical ties among touching obstacles):", len(tu.REORDERED))
for t in tu.REORDERED:
    print("   ", t)
print("PARITY SWEEP", "FAILED" if bad else "OK", f"({bad} mismatches)")
sys.exit(1 if bad else 0)
