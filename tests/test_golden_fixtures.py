"""CPU: the committed BASELINE-size golden graphs are what the oracle / the reference's HiGHS call produce
(guards the fixtures and the oracle against silent drift; the GPU tests compare the kernels with them)."""
import os

import numpy as np

from boundplanner_b200 import scenes
from oracle.set_graph import set_intersection
from tests.util import oracle_finder

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sets(g):
    off = np.concatenate(([0], np.cumsum(g["m"])))
    return [[g["rows"][off[s]: off[s + 1], :3], g["rows"][off[s]: off[s + 1], 3]] for s in range(len(g["m"]))]


def test_c2_golden_is_the_oracle_and_highs():
    g = np.load(os.path.join(GOLD, "c2_graph_golden.npz"))
    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
    assert np.array_equal(seeds, g["seeds"]) and len(g["m"]) == 256 and (g["status"] == 0).all()
    adj = np.unpackbits(g["adj_bits"], axis=1)[:, :256].astype(bool)
    assert adj.shape == (256, 256) and not np.tril(adj).any() and adj.sum() == 1007
    sets = _sets(g)
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    for s in (0, 97, 255):
        A, b, Q, p = f.find_set_around_point(seeds[s], fixed_mid=True, optimize=True)
        assert A.shape[0] == g["m"][s] and f.last_iters == g["iters"][s]
        assert np.abs(A - sets[s][0]).max() < 1e-12 and np.abs(b - sets[s][1]).max() < 1e-12
        assert np.abs(Q - g["q_ellipse"][s]).max() <= 1e-9 * np.abs(Q).max()
    rng = np.random.default_rng(0)
    for i, j in np.argwhere(adj)[rng.choice(1007, 40, replace=False)]:
        assert set_intersection(sets[i], sets[j], 0.01)[2]
    for _ in range(60):
        i, j = sorted(rng.choice(256, 2, replace=False))
        assert bool(set_intersection(sets[i], sets[j], 0.01)[2]) == bool(adj[i, j])


def test_c4_golden_shape():
    g = np.load(os.path.join(GOLD, "c4_sets_golden.npz"))
    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
    sel = g["seed_index"]
    assert len(sel) == 258 and np.array_equal(seeds[sel], g["seeds"]) and (g["status"] == 0).all()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    sets = _sets(g)
    k = 100
    A, b, Q, p = f.find_set_around_point(seeds[sel[k]], fixed_mid=True, optimize=True)
    assert A.shape[0] == g["m"][k] and np.abs(A - sets[k][0]).max() < 1e-12 and f.last_iters == g["iters"][k]
