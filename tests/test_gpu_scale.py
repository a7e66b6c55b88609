"""GPU: BASELINE-size configurations (C2 1k obstacles / 256 seeds, C4 10k obstacles / 2048 seeds)
through size-independent properties, plus oracle parity on a sample of seeds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import assert_rows_close, oracle_finder  # noqa: E402


@pytest.fixture(scope="module")
def geo():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry

    return geometry


def _check_batch_properties(geo, boxes, inflate, seeds, ws_min, ws_max, out, adj):
    A, b, m = out.A.cpu().numpy(), out.b.cpu().numpy(), out.m.cpu().numpy()
    status = out.status.cpu().numpy()
    P = out.p_mid.cpu().numpy()
    Q = out.q_ellipse.cpu().numpy()
    ok = status == 0
    assert ok.mean() > 0.95
    lb, ub = boxes[:, :3] - inflate, boxes[:, 3:] + inflate
    for s in np.where(ok)[0]:
        As, bs = A[s, : m[s]], b[s, : m[s]]
        assert np.all(As @ seeds[s] - bs < 0)                                   # seed strictly inside
        assert np.abs(np.linalg.norm(As, axis=1) - 1).max() < 1e-12              # unit normals
        assert np.all(A[s, m[s]:] == 0) and np.all(b[s, m[s]:] == 10.0)          # normalize_set_size padding
        # every obstacle is cut off by some picked row: min over its 8 vertices >= -1e-4
        if m[s] > 6:
            rows, rb = As[6:], bs[6:]
            vmin = (np.minimum(rows[None] * lb[:, None, :], rows[None] * ub[:, None, :]).sum(axis=2) - rb[None])
            assert np.all((vmin >= -1e-4).any(axis=1))
        # final ellipsoid (semi-axis matrix q_inv = Q^-1, quirk Q2) is inscribed in the set
        E = np.linalg.inv(Q[s])
        L = np.linalg.cholesky(0.5 * (E + E.T))
        assert np.all(np.linalg.norm(As @ L, axis=1) <= bs - As @ P[s] + 1e-7)
    # adjacency: symmetric-by-construction upper triangle; a set's centre inside another set (with margin)
    # implies an edge; p_mid of both far outside each other's AABB implies none is not guaranteed -> only sufficient checks
    idx = np.where(ok)[0]
    for i in idx[:64]:
        for j in idx[:64]:
            if j > i:
                in_both = np.all(A[i, : m[i]] @ P[j] - b[i, : m[i]] <= -0.011) and \
                          np.all(A[j, : m[j]] @ P[j] - b[j, : m[j]] <= -0.011)
                if in_both:
                    assert adj[i, j]


def test_c2_full_size_properties_and_sample_parity(geo):
    from boundplanner_b200 import scenes

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
    sc = geo.Scene(boxes, inflate)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
    adj = geo.unpack_adjacency(bits, seeds.shape[0]).cpu().numpy()
    _check_batch_properties(geo, boxes, inflate, seeds, ws_min, ws_max, out, adj)
    # idempotence: the same call gives bit-identical outputs (no atomics / race-dependent results)
    out2 = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    assert np.array_equal(out.A.cpu().numpy(), out2.A.cpu().numpy())
    assert np.array_equal(out.q_ellipse.cpu().numpy(), out2.q_ellipse.cpu().numpy())
    bits2 = geo.pair_feasible(out.A, out.b, out.m, 0.01)
    assert np.array_equal(bits.cpu().numpy(), bits2.cpu().numpy())
    # oracle parity on a sample of seeds + their pair checks (reference's linprog)
    from oracle.set_graph import intersection_margin, set_intersection

    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    sets = out.to_sets()
    sample = list(range(0, 256, 16))
    for s in sample:
        Ao, bo, Qo, po = f.find_set_around_point(seeds[s], fixed_mid=True)
        assert out.iters[s].item() == f.last_iters
        assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"seed {s}")
    for i in sample:
        for j in range(i + 1, 256, 7):
            ok = bool(set_intersection(sets[i], sets[j], 0.01)[2])
            if ok != bool(adj[i, j]):
                assert abs(intersection_margin(sets[i], sets[j], 0.01)) < 1e-6


def test_c4_large_scene(geo):
    """10k obstacles (distance table only, no closest-point cache in shared memory), 2048 seeds."""
    from boundplanner_b200 import scenes

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
    assert boxes.shape[0] == 10000 and seeds.shape[0] == 2048
    sc = geo.Scene(boxes, inflate)
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
    adj = geo.unpack_adjacency(bits, seeds.shape[0]).cpu().numpy()
    sub = slice(0, 256)
    import dataclasses

    out_sub = dataclasses.replace(out, A=out.A[sub], b=out.b[sub], m=out.m[sub], q_ellipse=out.q_ellipse[sub],
                                  p_mid=out.p_mid[sub], status=out.status[sub])
    _check_batch_properties(geo, boxes, inflate, seeds[sub], ws_min, ws_max, out_sub, adj[sub, sub])
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    sets = out.to_sets()
    for s in (0, 700, 2047):
        if out.status[s].item() != 0:
            continue
        Ao, bo, Qo, po = f.find_set_around_point(seeds[s], fixed_mid=True)
        assert_rows_close(sets[s][0], sets[s][1], Ao, bo, f"seed {s}")
    # row-block partition of the 2 096 128 pair checks == one call
    parts = [geo.pair_feasible(out.A, out.b, out.m, 0.01, r0, r1).cpu().numpy() for r0, r1 in ((0, 500), (500, 2048))]
    assert np.array_equal(np.vstack(parts), bits.cpu().numpy())


def test_row_overflow_and_reference_cap(geo):
    """Dense clutter: sets with more than 20 rows are produced (the reference would raise, quirk Q5)
    and m_max smaller than needed reports BP_ROW_OVERFLOW instead of writing out of bounds."""
    from boundplanner_b200 import scenes

    rng = np.random.default_rng(9)
    boxes = scenes.random_box_scene(4000, rng, 0.01, 0.03)
    seeds = scenes.free_points(64, boxes, 0.0, rng)
    sc = geo.Scene(boxes, 0.0)
    out = geo.build_sets_point(sc, seeds, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX, fixed_mid=True)
    m = out.m.cpu().numpy()
    assert (m > 20).any() and m.max() <= 48
    small = geo.build_sets_point(sc, seeds, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX, fixed_mid=True, m_max=12)
    st = small.status.cpu().numpy()
    assert (st == 2).any() and small.m.cpu().numpy().max() <= 12


def test_cuda_graph_pipeline_equals_eager(geo):
    """Static-buffer CUDA-graph replay of a step gives bit-identical results to the eager calls,
    also when the seeds change between replays."""
    import torch
    from boundplanner_b200 import scenes
    from boundplanner_b200.pipeline import SetGraphPipeline

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(1000, 128)
    sc = geo.Scene(boxes, inflate)
    pipe = SetGraphPipeline(sc, 64, ws_min, ws_max, fixed_mid=True, optimize=True, tol=0.01)
    for part in (seeds[:64], seeds[64:]):
        host = torch.as_tensor(part).pin_memory()
        A, b, m, q, p, status, bits = [t.clone() for t in pipe.run(host)]
        out = geo.build_sets_point(sc, part, ws_min, ws_max, fixed_mid=True, optimize=True)
        ebits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
        assert np.array_equal(A.numpy(), out.A.cpu().numpy()) and np.array_equal(b.numpy(), out.b.cpu().numpy())
        assert np.array_equal(m.numpy(), out.m.cpu().numpy()) and np.array_equal(q.numpy(), out.q_ellipse.cpu().numpy())
        assert np.array_equal(status.numpy(), out.status.cpu().numpy())
        assert np.array_equal(bits.numpy(), ebits.cpu().numpy())


def test_maximum_scene_size_and_loud_failure(geo):
    """The distance table lives in shared memory: N up to 28960 works (227 KB opt-in minus 768 B static), larger scenes fail loudly."""
    from boundplanner_b200 import _lib, scenes

    rng = np.random.default_rng(4)
    boxes = scenes.random_box_scene(28800, rng, 0.004, 0.012)
    seeds = scenes.free_points(4, boxes, 0.0, rng)
    sc = geo.Scene(boxes, 0.0)
    out = geo.build_sets_point(sc, seeds, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX, fixed_mid=True, optimize=False)
    assert (out.status.cpu().numpy() == 0).all() and (out.m.cpu().numpy() > 6).all()
    too_big = geo.Scene(scenes.random_box_scene(30000, rng, 0.004, 0.012), 0.0)
    with pytest.raises(_lib.BpGeoError, match="too large"):
        geo.build_sets_point(too_big, seeds, scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX)


def test_pipelined_end_to_end_equals_synchronous(geo):
    """PipelinedSetGraph (two steps in flight, D2H on a copy stream) delivers the same results as run()."""
    import torch

    from boundplanner_b200 import scenes
    from boundplanner_b200.pipeline import PipelinedSetGraph, SetGraphPipeline

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(1000, 64)
    sc = geo.Scene(boxes, inflate)
    rng = np.random.default_rng(9)
    batches = [scenes.free_points(64, boxes, inflate, rng) for _ in range(5)]
    hosts = [torch.as_tensor(b).pin_memory() for b in batches]
    ref = SetGraphPipeline(sc, 64, ws_min, ws_max, fixed_mid=True, optimize=True)
    want = [[t.clone() for t in ref.run(h)] for h in hosts]
    ps = PipelinedSetGraph(sc, 64, ws_min, ws_max, depth=2, fixed_mid=True, optimize=True)
    got = []
    for h in hosts:
        r = ps.take()
        if r is not None:
            got.append([t.clone() for t in r])
        ps.put(h)
    got += [[t.clone() for t in r] for r in ps.drain()]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert torch.equal(a, b)
