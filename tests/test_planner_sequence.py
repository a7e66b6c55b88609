"""Planned set sequence: the planner loop (boundplanner_b200/planner.py, restating
BoundPlanner.plan_convex_set_path up to the shortest path) run with the oracle
backend on CPU and, on the GPU box, with the kernel backend -- same sequences."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

from boundplanner_b200 import scenes
from boundplanner_b200.planner import SetSequencePlanner
from tests.util import OracleBackend

P0 = np.array([0.3, 0.0, 0.7])                  # boundplanner_example.py:89-92
P1 = np.array([0.45, -0.5, 0.2])
R0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()


def _plan(backend_cls, seed=0, **kw):
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    backend = backend_cls(boxes, inflate, list(ws_max), list(ws_min)) if backend_cls else None
    planner = SetSequencePlanner(boxes, inflate, list(ws_max), list(ws_min), backend=backend,
                                 rng=np.random.default_rng(seed))
    return planner.plan_set_sequence(P0.copy(), P1.copy(), R0, R0, **kw), planner


def test_example_plan_with_oracle_backend():
    res, planner = _plan(OracleBackend)
    path, ids, p_via = res["path"], res["set_ids"], res["p_via"]
    assert path[0] == 0 and path[-1] == 1 and ids[-1] == 1 and len(ids) == len(p_via) - 1
    assert np.allclose(p_via[0], P0) and np.allclose(p_via[-1], P1)
    g, ig = res["graph"], res["inter_graph"]
    # every via point lies in the two consecutive sets it connects (they are projections onto intersections)
    for k in range(1, len(p_via) - 1):
        for sid in (ids[k - 1], ids[k]):
            A, b = g.nodes[sid]["cset"]
            assert np.max(A @ p_via[k] - b) < 1e-7
    # consecutive sets of the sequence intersect
    from oracle.set_graph import set_intersection

    for a, b_ in zip(ids[:-1], ids[1:]):
        assert set_intersection(g.nodes[a]["cset"], g.nodes[b_]["cset"], 0.01)[2]
    # determinism with a seeded generator (quirk Q4)
    res2, _ = _plan(OracleBackend)
    assert res2["path"] == path and res2["set_ids"] == ids and np.allclose(res2["p_via"], p_via)


@pytest.mark.gpu
def test_example_plan_gpu_matches_oracle_sequence():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200.planner import GpuBackend

    for seed, kw in ((0, {}), (3, {}), (1, {"first_sample": np.array([0.5, -0.2, 0.6])})):
        want, _ = _plan(OracleBackend, seed, **kw)
        got, _ = _plan(GpuBackend, seed, **kw)
        assert got["path"] == want["path"], f"seed {seed}"                       # index work: exact
        assert got["set_ids"] == want["set_ids"]
        assert got["graph"].number_of_nodes() == want["graph"].number_of_nodes()
        assert sorted((d["id0"], d["id1"]) for _, d in got["inter_graph"].nodes.items()) == \
            sorted((d["id0"], d["id1"]) for _, d in want["inter_graph"].nodes.items())     # graph adjacency
        assert np.abs(got["p_via"] - want["p_via"]).max() < 1e-6
        for (ga, gb), (wa, wb) in zip(got["sets_via"], want["sets_via"]):
            assert ga.shape == wa.shape and np.abs(ga - wa).max() < 1e-6 and np.abs(gb - wb).max() < 1e-6


def _plan_c3(backend_cls, i):
    obstacles, inflate, start, end, ws_min, ws_max = scenes.config_c3_query(i)
    backend = backend_cls(obstacles, inflate, list(ws_max), list(ws_min))
    planner = SetSequencePlanner(obstacles, inflate, list(ws_max), list(ws_min), backend=backend,
                                 rng=np.random.default_rng(i))
    try:
        return planner.plan_set_sequence(start.copy(), end.copy(), R0, R0)
    except (RuntimeError, ValueError) as e:          # the reference raises on these queries too
        return {"error": type(e).__name__ + ": " + str(e).split("(")[0]}


def test_c3_queries_with_oracle_backend():
    """C3-style queries (200 obstacles) exercise the random-sampling branch of the loop."""
    n_ok = 0
    for i in range(3):
        res = _plan_c3(OracleBackend, i)
        if "error" not in res:
            n_ok += 1
            assert res["path"][0] == 0 and res["path"][-1] == 1
    assert n_ok >= 1


@pytest.mark.gpu
def test_c3_queries_gpu_matches_oracle_sequence():
    from boundplanner_b200.planner import GpuBackend

    n_ok = 0
    for i in (0, 1, 5, 7, 8, 11):       # 1: "Exceeded max iterations", 5: >20 rows (ValueError), the others plan
        want = _plan_c3(OracleBackend, i)
        got = _plan_c3(GpuBackend, i)
        if "error" in want:
            assert got.get("error") == want["error"], f"query {i}"
            continue
        n_ok += 1
        assert got["path"] == want["path"] and got["set_ids"] == want["set_ids"], f"query {i}"
        assert np.abs(got["p_via"] - want["p_via"]).max() < 1e-6
    assert n_ok >= 3


@pytest.mark.gpu
def test_batched_lockstep_planner_matches_sequential():
    """plan_batch (all queries advance together, one batched kernel call per primitive and round, one
    scene per query in a SceneBatch) == the sequential planner, query by query, incl. the error exits."""
    from boundplanner_b200.planner import GpuBackend, plan_batch

    ids = list(range(12))
    queries = []
    for i in ids:
        obstacles, inflate, start, end, ws_min, ws_max = scenes.config_c3_query(i)
        queries.append(dict(obstacles=obstacles, start=start, end=end, r0=R0, r1=R0))
    results, stats = plan_batch(queries, 0.01, list(ws_max), list(ws_min), rng_seeds=ids)
    assert stats["rounds"] > 10 and stats["kernel_batches"] < 40 * stats["rounds"]
    n_ok = 0
    for i, res in zip(ids, results):
        want = _plan_c3(GpuBackend, i)
        if "error" in want:
            assert isinstance(res, Exception), f"query {i}"
            assert type(res).__name__ + ": " + str(res).split("(")[0] == want["error"], f"query {i}"
            continue
        n_ok += 1
        assert not isinstance(res, Exception), f"query {i}: {res}"
        assert res["path"] == want["path"] and res["set_ids"] == want["set_ids"], f"query {i}"
        assert np.abs(res["p_via"] - want["p_via"]).max() < 1e-9
    assert n_ok >= 3


class _OracleLockstepExecutor:
    """CPU stand-in for planner.BatchedGpuExecutor: answers every pending request with the oracle."""

    def __init__(self, queries, inflate, ws_max, ws_min):
        self.backends = [OracleBackend(q["obstacles"], inflate, ws_max, ws_min) for q in queries]
        self.calls = 0

    def execute(self, pending):
        out = {}
        self.calls += 1
        for qid, req in pending.items():
            try:
                out[qid] = self.backends[qid].execute(req)
            except (RuntimeError, ValueError) as e:
                out[qid] = e
        return out


def test_lockstep_driver_equals_sequential_driver_on_cpu():
    """The lock-step driver (plan_batch) is pure host logic: with the oracle answering the requests it
    must reproduce the sequential driver query by query (mixed scenes, early finishers, error exits)."""
    from boundplanner_b200.planner import plan_batch

    ids = [0, 5, 7]
    queries = []
    for i in ids:
        obstacles, inflate, start, end, ws_min, ws_max = scenes.config_c3_query(i)
        queries.append(dict(obstacles=obstacles, start=start, end=end, r0=R0, r1=R0))
    ex = _OracleLockstepExecutor(queries, 0.01, list(ws_max), list(ws_min))
    results, stats = plan_batch(queries, 0.01, list(ws_max), list(ws_min), rng_seeds=ids, executor=ex)
    assert stats["rounds"] == ex.calls and stats["rounds"] > 5
    for i, res in zip(ids, results):
        want = _plan_c3(OracleBackend, i)
        if "error" in want:
            assert isinstance(res, Exception) and type(res).__name__ + ": " + str(res).split("(")[0] == want["error"]
        else:
            assert res["path"] == want["path"] and res["set_ids"] == want["set_ids"]
            assert np.array_equal(res["p_via"], want["p_via"])


def test_device_loop_protocol_equals_host_loop_on_cpu():
    """The planner's device-loop requests (candidates drawn ahead + generator rewind, set built in the sampling
    round, duplicate distance and shortest path from the backend) give the same plan -- and leave the generator in
    the same state -- as the reference's one-at-a-time host loops.  Both sides use the oracle's primitives."""
    from scipy.spatial.transform import Rotation as R

    from boundplanner_b200 import scenes
    from boundplanner_b200.planner import SetSequencePlanner
    from tests.util import OracleBackend, OracleDeviceLoopBackend

    r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
    cases = []
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    cases.append((boxes, inflate, np.array([0.3, 0.0, 0.7]), np.array([0.45, -0.5, 0.2]), ws_min, ws_max, 3))
    for i in (7, 8):
        ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
        cases.append((ob, infl, st, en, wmin, wmax, i))
    for ob, infl, st, en, wmin, wmax, seed in cases:
        outs = []
        for cls in (OracleBackend, OracleDeviceLoopBackend):
            backend = cls(ob, infl, list(wmax), list(wmin))
            pl = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=backend, rng=np.random.default_rng(seed))
            pl.sample_chunk = 5                       # several chunks per sample: exercises the rewind
            assert pl.device_loop == (cls is OracleDeviceLoopBackend)
            try:
                res = pl.plan_set_sequence(st.copy(), en.copy(), r0, r0)
                outs.append((res["path"], res["set_ids"], res["p_via"], pl.rng.uniform(0, 1, 4)))
            except (RuntimeError, ValueError) as e:
                outs.append((type(e).__name__, str(e), None, pl.rng.uniform(0, 1, 4)))
        a, b = outs
        assert a[0] == b[0] and a[1] == b[1]
        if a[2] is not None:
            assert np.array_equal(a[2], b[2])
        assert np.array_equal(a[3], b[3])             # same number of draws consumed


def test_replanning_branch_and_start_fallbacks_with_oracle_backend():
    """BoundPlanner.py:231-276 / :296-324: after a first plan, replanning from a point on the first segment builds
    the start set around the segment start .. horizon point that the previous sets still cover; the fallbacks for a
    start segment in collision reuse the previous end set (IndexError on a first plan) or rebuild around `start`."""
    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    backend = OracleBackend(boxes, inflate, list(ws_max), list(ws_min))
    pl = SetSequencePlanner(boxes, inflate, list(ws_max), list(ws_min), backend=backend, rng=np.random.default_rng(0))
    first = pl.plan_set_sequence(P0.copy(), P1.copy(), R0, R0)
    assert len(pl.sets_via_prev) == len(first["sets_via"]) >= 1
    p_via = first["p_via"]
    # MPC horizon: points along the first segment(s) of the planned path
    t = np.linspace(0.05, 0.6, 8)
    horizon = np.array([p_via[0] + ti * (p_via[1] - p_via[0]) for ti in t])
    start = p_via[0] + 0.02 * (p_via[1] - p_via[0])
    idx = pl.replanning_horizon_index(start, horizon)
    s0 = pl.sets_via_prev[0]
    assert np.max(s0[0] @ start - s0[1]) < 1e-8
    inside = np.max(s0[0] @ horizon.T - s0[1][:, None], axis=0) < 1e-8
    assert idx == (len(horizon) - 1 if inside.all() else max(1, int(np.argmin(inside)) - 1))
    assert pl.replanning_horizon_index(start, horizon, new_obs=True) == 1
    prev_sets = [[a.copy(), b.copy()] for a, b in pl.sets_via_prev]
    res = pl.plan_set_sequence(start.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=horizon)
    a0, b0 = res["graph"].nodes[0]["cset"]
    assert np.max(a0 @ start - b0) < 1e-6 and np.max(a0 @ pl.p_horizon_max - b0) < 1e-6      # segment inside its set
    assert res["path"][0] == 0 and res["path"][-1] == 1 and np.allclose(res["p_via"][-1], P1)
    # fallbacks: a start segment that runs through an obstacle
    lo, hi = boxes[0][:3], boxes[0][3:]
    through = 0.5 * (lo + hi)
    pl2 = SetSequencePlanner(boxes, inflate, list(ws_max), list(ws_min), backend=backend, rng=np.random.default_rng(0))
    with pytest.raises(IndexError):                       # first plan: no previous end set to reuse
        pl2.plan_set_sequence(P0.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=np.array([P0, through]))
    pl2.sets_via_prev = prev_sets
    res2 = pl2.plan_set_sequence(P0.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=np.array([P0, through]))
    a0, b0 = res2["graph"].nodes[0]["cset"]
    want_a, want_b = backend.reduce_ineqs(np.array(prev_sets[-1][0]), np.array(prev_sets[-1][1]))
    assert np.array_equal(a0, want_a) and np.array_equal(b0, want_b)                          # old end set reused
    assert np.array_equal(res2["graph"].nodes[0]["q_ellipse"], np.eye(3))
    pl3 = SetSequencePlanner(boxes, inflate, list(ws_max), list(ws_min), backend=backend, rng=np.random.default_rng(0))
    res3 = pl3.plan_set_sequence(P0.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=np.array([P0, through]),
                                 new_obs=True)                                                # rebuilt around start
    a0, b0 = res3["graph"].nodes[0]["cset"]
    assert np.max(a0 @ P0 - b0) < 1e-6
    with pytest.raises(AttributeError):                   # quirk Q12: start inside an inflated obstacle + new_obs
        pl3.plan_set_sequence(through.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=np.array([through, P0]),
                              new_obs=True)


@pytest.mark.gpu
def test_replanning_gpu_matches_oracle_sequence():
    """The replanning call (:231-276) through the kernels == through the oracle: same start set, path and sequence."""
    from boundplanner_b200.planner import GpuBackend

    outs = []
    for cls in (OracleBackend, GpuBackend):
        boxes, ws_min, ws_max, inflate = scenes.example_scene()
        backend = cls(boxes, inflate, list(ws_max), list(ws_min))
        pl = SetSequencePlanner(boxes, inflate, list(ws_max), list(ws_min), backend=backend, rng=np.random.default_rng(0))
        first = pl.plan_set_sequence(P0.copy(), P1.copy(), R0, R0)
        p_via = first["p_via"]
        horizon = np.array([p_via[0] + ti * (p_via[1] - p_via[0]) for ti in np.linspace(0.05, 0.6, 8)])
        start = p_via[0] + 0.02 * (p_via[1] - p_via[0])
        res = pl.plan_set_sequence(start.copy(), P1.copy(), R0, R0, replanning=True, p_horizon=horizon)
        outs.append((first, res, pl.p_horizon_max.copy()))
    (f0, r0_, h0), (f1, r1_, h1) = outs
    assert f0["path"] == f1["path"] and f0["set_ids"] == f1["set_ids"]
    assert np.abs(h0 - h1).max() < 1e-6
    assert r0_["path"] == r1_["path"] and r0_["set_ids"] == r1_["set_ids"]
    assert np.abs(r0_["p_via"] - r1_["p_via"]).max() < 1e-6
    a0, b0 = r0_["graph"].nodes[0]["cset"]
    a1, b1 = r1_["graph"].nodes[0]["cset"]
    assert a0.shape == a1.shape and np.abs(a0 - a1).max() < 1e-6 and np.abs(b0 - b1).max() < 1e-6
