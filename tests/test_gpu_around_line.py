"""GPU: find_set_around_line (ConvexSetFinder.py:242-307) and mvie_socp_fixed_r (:564-588) -- SURVEY §8 rows
a11 / a7 -- through the C ABI against the oracle.  The reference planner's call is commented out
(BoundPlanner.py:378-380) but both are part of ConvexSetFinder's surface."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import RTOL, assert_rows_close, oracle_finder  # noqa: E402


def _segments(rng, boxes, inflate, n, ws_min, ws_max, lmin=0.05, lmax=0.4):
    """Random segments whose midpoint is free (the loop is seeded there)."""
    from boundplanner_b200 import scenes

    mids = scenes.free_points(n, boxes, inflate + 0.02, rng, ws_min, ws_max)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    d *= rng.uniform(lmin, lmax, (n, 1))
    return mids - d / 2, d


def _run_oracle(f, p0, dp1, optimize):
    try:
        a, b, q, p = f.find_set_around_line(p0, dp1, optimize=optimize)
        return ("ok", np.array(a), np.array(b), q, p, f.last_iters)
    except RuntimeError as e:
        return ("RuntimeError", str(e))
    except ValueError as e:
        return ("ValueError", str(e))


def test_dropin_find_set_around_line_example_scene():
    import torch

    assert torch.cuda.is_available()
    import boundplanner_b200 as bp
    from boundplanner_b200 import scenes
    from oracle.convex_set_finder import ConvexSetFinder as OracleFinder
    from oracle.obstacles import obstacle_reps

    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    gpu = bp.ConvexSetFinder(obs_sets, pts, list(ws_max), list(ws_min))
    ora = OracleFinder(obs_sets, pts, list(ws_max), list(ws_min))
    rng = np.random.default_rng(21)
    p0s, dps = _segments(rng, boxes, inflate, 14, ws_min, ws_max, 0.05, 0.5)
    n_ok = n_err = 0
    for p0, dp1 in zip(p0s, dps):
        for optimize in (True, False):
            ref = _run_oracle(ora, p0, dp1, optimize)
            if ref[0] != "ok":
                with pytest.raises(RuntimeError if ref[0] == "RuntimeError" else ValueError):
                    gpu.find_set_around_line(p0, dp1, optimize=optimize)
                n_err += 1
                continue
            a, b, q, p = gpu.find_set_around_line(p0, dp1, optimize=optimize)
            assert isinstance(a, list) and isinstance(b, list)          # :307 returns compute_polyhedron's lists
            assert_rows_close(np.array(a), np.array(b), ref[1], ref[2], "around-line set")
            assert np.abs(q - ref[3]).max() <= 1e-5 * np.abs(ref[3]).max()
            assert np.abs(p - ref[4]).max() <= RTOL
            n_ok += 1
    assert n_ok >= 10


def test_batched_around_line_c2_scene():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, scenes

    rng = np.random.default_rng(0)
    boxes = scenes.random_box_scene(1000, rng, 0.02, 0.08)
    inflate = 0.01
    ws_min, ws_max = scenes.WORKSPACE_MIN, scenes.WORKSPACE_MAX
    p0s, dps = _segments(rng, boxes, inflate, 48, ws_min, ws_max, 0.02, 0.15)
    scene = geo.Scene(boxes, inflate)
    out = geo.build_sets_around_line(scene, p0s, dps, ws_min, ws_max, optimize=True)
    status, m, iters = out.status.cpu().numpy(), out.m.cpu().numpy(), out.iters.cpu().numpy()
    A, b, Q, P = out.A.cpu().numpy(), out.b.cpu().numpy(), out.q_ellipse.cpu().numpy(), out.p_mid.cpu().numpy()
    f = oracle_finder(boxes, inflate, ws_min, ws_max)
    n_ok = 0
    for s in range(0, 48, 2):
        ref = _run_oracle(f, p0s[s], dps[s], True)
        if ref[0] == "RuntimeError":
            assert status[s] in (1, 3), (s, status[s], ref[1])     # ellipse violation / infeasible a_lb
            continue
        assert ref[0] == "ok" and status[s] == 0, (s, status[s], ref)
        assert iters[s] == ref[5]                                   # same number of loop passes
        assert_rows_close(A[s, :m[s]], b[s, :m[s]], ref[1], ref[2], f"segment {s}")
        assert (A[s, m[s]:] == 0).all() and (b[s, m[s]:] == 10).all()
        assert np.abs(Q[s] - ref[3]).max() <= 1e-5 * np.abs(ref[3]).max()
        assert np.abs(P[s] - (p0s[s] + dps[s] / 2)).max() <= 1e-15
        n_ok += 1
    assert n_ok >= 12


def test_mvie_socp_fixed_r_dropin():
    import torch

    assert torch.cuda.is_available()
    import boundplanner_b200 as bp
    from boundplanner_b200 import scenes
    from oracle import mvie as omvie
    from oracle.convex_set_finder import line_frame
    from oracle.obstacles import obstacle_reps

    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    gpu = bp.ConvexSetFinder(obs_sets, pts, list(ws_max), list(ws_min))
    rng = np.random.default_rng(5)
    box = np.vstack((np.eye(3), -np.eye(3)))
    for trial in range(12):
        k = rng.integers(2, 13)
        c = rng.uniform(-0.4, 0.4, 3)
        An = rng.normal(size=(k, 3))
        An /= np.linalg.norm(An, axis=1)[:, None]
        A = np.vstack((box, An))
        b = np.concatenate((np.ones(6), An @ c + rng.uniform(0.02, 0.4, k)))
        R, _ = line_frame(rng.normal(size=3))
        a_lb = rng.uniform(0, 0.02)
        qn, qe, eigs = gpu.mvie_socp_fixed_r(A, b, c, R, a_lb)
        qno, qeo, so = omvie.mvie_fixed_r(A, b, c, R, a_lb)
        assert np.abs(eigs - so).max() <= 1e-8 * so.max()
        assert np.abs(qn - qno).max() <= 1e-7 * np.abs(qno).max()
        assert np.abs(qe - qeo).max() <= 1e-7 * np.abs(qeo).max()
    # a_lb that cannot be met: the reference's SOCP is infeasible (x.value is None -> TypeError there)
    with pytest.raises(RuntimeError):
        gpu.mvie_socp_fixed_r(box, 0.2 * np.ones(6), np.zeros(3), np.eye(3), 0.5)
