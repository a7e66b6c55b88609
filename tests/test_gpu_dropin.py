"""GPU: the reference-facing drop-in classes behave like the reference's (as
restated by the oracle) on the C1 example scene (boundplanner_example.py:19-92)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.util import RTOL, assert_rows_close  # noqa: E402


@pytest.fixture(scope="module")
def c1():
    import torch

    assert torch.cuda.is_available()
    import boundplanner_b200 as bp
    from boundplanner_b200 import scenes
    from oracle.convex_set_finder import ConvexSetFinder as OracleFinder
    from oracle.obstacles import obstacle_reps

    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)          # what BoundPlanner.add_obstacle_reps builds
    gpu = bp.ConvexSetFinder(obs_sets, pts, list(ws_max), list(ws_min))
    ora = OracleFinder(obs_sets, pts, list(ws_max), list(ws_min))
    return gpu, ora, obs_sets, pts


def test_attributes_match_reference_surface(c1):
    gpu, ora, obs_sets, pts = c1
    for name in ("obs_sets", "obs_points_sets", "e_max", "e_min", "max_iter", "proj_time", "ell_time",
                 "set_line_time", "rng", "find_set_around_point", "find_set_collision_avoidance",
                 "compute_polyhedron", "compute_set_projs", "compute_set_projs_line", "mvie_socp",
                 "mvie_socp_fixed_mid", "mvie_socp_fixed_r", "init_halfspaces", "init_halfspaces_point"):
        assert hasattr(gpu, name), name
    assert gpu.max_iter == 5 and len(gpu.obs_sets) == 12


def test_example_start_set_and_end_set(c1):
    """The two calls plan_convex_set_path makes first (BoundPlanner.py:278-294, :381-389)."""
    gpu, ora, *_ = c1
    p0 = np.array([0.3, 0.0, 0.7])
    p1 = np.array([0.45, -0.5, 0.2])
    A, b, Q, p = gpu.find_set_around_point(p0, fixed_mid=True)
    Ao, bo, Qo, po = ora.find_set_around_point(p0, fixed_mid=True)
    assert_rows_close(A, b, Ao, bo, "start set")
    assert np.abs(Q - Qo).max() <= 1e-5 * np.abs(Qo).max() and np.abs(p - po).max() <= RTOL
    assert A.dtype == np.float64 and A.flags.writeable          # callers mutate the arrays in place
    l_ee = np.array([0.0, 0.0, 0.05])                            # R_y(90) @ [-0.05, 0, 0]
    out = gpu.find_set_collision_avoidance(p1, p1 + l_ee, True)
    outo = ora.find_set_collision_avoidance(p1, p1 + l_ee, True)
    assert len(out) == 5 and out[4] == outo[4]
    assert_rows_close(out[0], out[1], outo[0], outo[1], "end set")
    assert np.abs(out[2] - outo[2]).max() <= 1e-5 * np.abs(outo[2]).max()
    assert np.abs(out[3] - outo[3]).max() <= RTOL


def test_mpc_call_site(c1):
    """BoundMPC.py:486-488: find_set_collision_avoidance(pl, pf, limit_space=True, e_max=0.7)."""
    gpu, ora, *_ = c1
    rng = np.random.default_rng(2)
    for _ in range(6):
        pl = np.array([0.3, 0.0, 0.7]) + rng.normal(size=3) * 0.05
        pf = pl + rng.normal(size=3) * 0.05
        A, b, coll = gpu.find_set_collision_avoidance(pl, pf, limit_space=True, e_max=0.7)
        Ao, bo, collo = ora.find_set_collision_avoidance(pl, pf, limit_space=True, e_max=0.7)
        assert coll == collo
        assert_rows_close(A, b, Ao, bo, "mpc line set")


def test_component_methods(c1):
    gpu, ora, *_ = c1
    p = np.array([0.3, 0.0, 0.7])
    q_inv = np.diag([1e-4] * 3)
    q_ell = np.diag([1e4] * 3)
    y = gpu.compute_set_projs(gpu.obs_sets, p, q_inv)
    yo = ora.compute_set_projs(ora.obs_sets, p, q_inv)
    assert np.abs(y - yo).max() < 1e-9
    a0, b0 = gpu.init_halfspaces()
    a_set, b_set = gpu.compute_polyhedron(q_inv, q_ell, p, a0, b0)
    ao, bo = ora.compute_polyhedron(q_inv, q_ell, p, *ora.init_halfspaces())
    assert isinstance(a_set, list) and isinstance(b_set, list)
    assert_rows_close(np.array(a_set), np.array(b_set), np.array(ao), np.array(bo), "polyhedron")
    A, b = np.array(a_set), np.array(b_set)
    q1, c1_ = gpu.mvie_socp_fixed_mid(A, b, p)
    q1o, _ = ora.mvie_socp_fixed_mid(A, b, p)
    assert c1_ is p and np.abs(q1 - q1o).max() <= RTOL * np.abs(q1o).max()
    q2, c2 = gpu.mvie_socp(A, b)                                 # no hint, like the reference signature
    q2o, c2o = ora.mvie_socp(A, b)
    assert np.abs(q2 - q2o).max() <= RTOL * np.abs(q2o).max() and np.abs(c2 - c2o).max() <= RTOL
    x, phi = gpu.compute_set_projs_line(gpu.obs_sets, p, p + np.array([0.1, -0.2, 0.05]))
    xo, phio = ora.compute_set_projs_line(ora.obs_sets, p, p + np.array([0.1, -0.2, 0.05]))
    assert np.abs(x - xo).max() < 1e-9 and np.abs(phi - phio).max() < 1e-9


def test_error_behaviour(c1):
    gpu, ora, obs_sets, pts = c1
    # a seed inside an (inflated) obstacle: the reference raises RuntimeError("Ellipse violates constraints")
    inside = np.array([0.6, 0.0, -0.05])
    with pytest.raises(RuntimeError, match="Ellipse violates constraints"):
        gpu.find_set_around_point(inside, fixed_mid=True)
    with pytest.raises(RuntimeError, match="Ellipse violates constraints"):
        ora.find_set_around_point(inside, fixed_mid=True)
    # more than 20 rows: the reference's MVIE buffers overflow with a ValueError (quirk Q5)
    A = np.vstack([np.eye(3), -np.eye(3)] * 4)
    b = np.ones(24)
    with pytest.raises(ValueError):
        gpu.mvie_socp_fixed_mid(A, b, np.zeros(3))
    # non-box obstacles need their vertices (general polytopes, tests/test_gpu_polytopes.py); without them: loud
    import boundplanner_b200 as bp

    bad = [[np.vstack((np.eye(3), -np.eye(3), np.ones((1, 3)))), np.ones(7)]]
    with pytest.raises(ValueError, match="obs_points_sets"):
        bp.ConvexSetFinder(bad, [], [1, 1, 1], [-1, -1, 0])
    sixteen = [[np.vstack((np.eye(3), -np.eye(3), np.ones((10, 3)))), np.ones(16)]]
    with pytest.raises(ValueError, match="more than 15 rows"):
        bp.ConvexSetFinder(sixteen, [np.zeros((8, 3))], [1, 1, 1], [-1, -1, 0])


def test_scene_update_through_attribute_assignment(c1):
    """add_obstacle_reps(update=True) assigns set_finder.obs_sets (BoundPlanner.py:150-152)."""
    import boundplanner_b200 as bp
    from boundplanner_b200 import scenes
    from oracle.convex_set_finder import ConvexSetFinder as OracleFinder
    from oracle.obstacles import obstacle_reps

    boxes, ws_min, ws_max, inflate = scenes.example_scene()
    obs_sets, pts, _ = obstacle_reps(boxes[:6], inflate)
    gpu = bp.ConvexSetFinder(obs_sets, pts, list(ws_max), list(ws_min))
    obs2, pts2, _ = obstacle_reps(boxes, inflate)
    gpu.obs_sets = obs2.copy()
    gpu.obs_points_sets = pts2.copy()
    ora = OracleFinder(obs2, pts2, list(ws_max), list(ws_min))
    p0 = np.array([0.3, 0.0, 0.7])
    A, b, _, _ = gpu.find_set_around_point(p0, fixed_mid=True)
    Ao, bo, _, _ = ora.find_set_around_point(p0, fixed_mid=True)
    assert_rows_close(A, b, Ao, bo, "after update")


def test_set_intersection_dropin():
    import boundplanner_b200 as bp
    from oracle.set_graph import set_intersection as ref_intersection

    box = np.vstack((np.eye(3), -np.eye(3)))
    s1 = [box, np.array([1, 1, 1, 0, 0, 0.0])]
    s2 = [box, np.array([1.5, 1.5, 1.5, -0.5, -0.5, -0.5])]
    s3 = [box, np.array([3, 3, 3, -2, -2, -2.0])]
    x, inter, ok = bp.set_intersection(s1, s2, tol=0.01)
    xr, interr, okr = ref_intersection(s1, s2, tol=0.01)
    assert ok and okr and np.array_equal(inter[0], interr[0]) and np.array_equal(inter[1], interr[1])
    assert np.max(inter[0] @ x - (inter[1] - 0.01)) <= 1e-9        # the returned point is inside the shrunk sets
    x, _, ok = bp.set_intersection(s1, s3, tol=0.01)
    assert not ok and x is None and not ref_intersection(s1, s3, tol=0.01)[2]
    adj = bp.adjacency([s1, s2, s3], tol=0.01)
    assert adj.tolist() == [[False, True, False], [True, False, False], [False, False, False]]


def test_robot_model_dropin():
    import boundplanner_b200 as bp
    from oracle import fk_iiwa14 as ofk

    model = bp.RobotModel()
    q = np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6, 0.7])
    assert np.abs(model.fk_pos(q) - ofk.fk_pos(q)).max() < 1e-12
    for i in range(7):
        assert np.abs(model.fk_pos_col(q, i) - ofk.fk_pos_col(q, i)).max() < 1e-12
    assert np.abs(model.hom_transform_endeffector(q) - ofk.hom_transform_endeffector(q)).max() < 1e-12
    assert np.abs(model.fk(q) - ofk.fk(q)).max() < 1e-10
    assert np.abs(model.jacobian_fk(q) - ofk.jacobian_fk(q)).max() < 1e-12
    with pytest.raises(NotImplementedError):
        model.fk_pos([0.0] * 7)                                  # non-ndarray = symbolic branch of the reference


def test_against_committed_golden_fixtures(c1):
    """Kernels vs tests/golden/*.npz (generated by tests/golden/make_golden.py)."""
    import os

    import boundplanner_b200 as bp

    gpu = c1[0]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_sets_golden.npz"))
    A, b, Q, p = gpu.find_set_around_point(g["p0"], fixed_mid=True)
    assert_rows_close(A, b, g["A0"], g["b0"], "golden start set")
    assert np.abs(Q - g["Q0"]).max() <= 1e-5 * np.abs(g["Q0"]).max() and np.abs(p - g["c0"]).max() <= RTOL
    A, b, Q, p, coll = gpu.find_set_collision_avoidance(g["p1"], g["p1"] + g["l_ee"], True)
    assert_rows_close(A, b, g["A1"], g["b1"], "golden end set")
    assert np.abs(Q - g["Q1"]).max() <= 1e-5 * np.abs(g["Q1"]).max() and bool(coll) == bool(g["collision1"])
    # FK kernel vs the reference's own serialized CasADi functions (fk_reference_blobs.npz)
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_reference_blobs.npz"))
    p_ee, p_col, T, J = bp.RobotModel.fk_batch(ref["q"], want_pose=True, want_jacobian=True)
    assert np.abs(p_ee.cpu().numpy() - ref["fk_pos"]).max() < 1e-12
    assert np.abs(p_col.cpu().numpy()[:, :6] - ref["fk_pos_col"]).max() < 1e-12
    assert np.abs(T.cpu().numpy() - ref["hom_trans"]).max() < 1e-12
    assert np.abs(J.cpu().numpy() - ref["jacobian"]).max() < 1e-12
    fk = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_golden.npz"))
    p_ee, p_col, T, J = bp.RobotModel.fk_batch(fk["q"], want_pose=True, want_jacobian=True)
    assert np.abs(p_ee.cpu().numpy() - fk["p_ee"]).max() < 1e-12
    assert np.abs(p_col.cpu().numpy() - fk["p_col"]).max() < 1e-12
    assert np.abs(T.cpu().numpy() - fk["T_ee"]).max() < 1e-12
    assert np.abs(J.cpu().numpy() - fk["jac"]).max() < 1e-12


def test_forward_kinematics_call_shapes_of_the_mpc_loop():
    """RobotModel.forward_kinematics / djacobian_fk / acceleration_ee (RobotModel.py:70-77, 233-262) through the
    shim with the call shapes of MPCNode.py:38,118 and util_functions.py:57-62 (integrate_joint), pinned to the
    reference's jacobian.ca differentiated along dq."""
    import boundplanner_b200 as bp
    from oracle import fk_iiwa14 as ofk

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "fk_reference_blobs.npz"))
    model = bp.RobotModel()
    q0 = np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, 0.0])           # boundplanner_with_mpc_example.py:20-26
    p0, _, _ = model.forward_kinematics(q0, q0)                       # MPCNode.py:38
    assert np.abs(p0 - ofk.fk(q0)).max() < 1e-12
    for i in (1, 5, 17, 40):
        q, dq = ref["q"][i], ref["dq"][i]
        p_lie, jac_fk, djac_fk = model.forward_kinematics(q, dq)      # MPCNode.py:118
        assert p_lie.shape == (6,) and jac_fk.shape == (6, 7) and djac_fk.shape == (6, 7)
        assert np.abs(p_lie - ofk.fk(q)).max() < 1e-12
        assert np.abs(jac_fk - ref["jacobian"][i]).max() < 1e-12
        assert np.abs(djac_fk - ref["djacobian"][i]).max() < 1e-12
        assert np.abs(model.djacobian_fk(q, dq) - ref["djacobian"][i]).max() < 1e-12
        ddq = 0.1 * dq[::-1]
        an = djac_fk @ dq + jac_fk @ ddq                               # util_functions.py:61
        assert np.abs(model.acceleration_ee(q, dq, ddq) - an).max() < 1e-12
        assert np.abs(model.acceleration_ee(q, dq, ddq) - ofk.acceleration_ee(q, dq, ddq)).max() < 1e-12
        vn = np.concatenate((model.velocity_ee(q, dq), model.omega_ee(q, dq)))   # util_functions.py:60
        assert np.abs(vn - ref["jacobian"][i] @ dq).max() < 1e-12
    # batched form, ragged tile (B not a multiple of the CTA size)
    T, J, dJ = bp.RobotModel.forward_kinematics_batch(ref["q"], ref["dq"])
    assert np.abs(T.cpu().numpy() - ref["hom_trans"]).max() < 1e-12
    assert np.abs(J.cpu().numpy() - ref["jacobian"]).max() < 1e-12
    assert np.abs(dJ.cpu().numpy() - ref["djacobian"]).max() < 1e-12
    qq = np.tile(ref["q"], (3, 1))[:150]
    dd = np.tile(ref["dq"], (3, 1))[:150]
    _, _, dJ2 = bp.RobotModel.forward_kinematics_batch(qq, dd)
    assert np.array_equal(dJ2.cpu().numpy()[:64], dJ.cpu().numpy()) and np.array_equal(dJ2.cpu().numpy()[128:150], dJ.cpu().numpy()[:22])
