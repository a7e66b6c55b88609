"""GPU: the per-step geometry of BoundMPC (BoundMPC.py:480-496) -- 12 FK + 6 line sets -- vs the oracle,
on the example scene with obs_size_increase = 0 as BoundMPC builds its planner (BoundMPC.py:265)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mpc_step_geometry_matches_oracle():
    import torch

    assert torch.cuda.is_available()
    from boundplanner_b200 import geometry as geo, mpc_geometry, scenes
    from oracle import fk_iiwa14 as ofk
    from oracle.obstacles import normalize_set_size
    from tests.util import oracle_finder

    boxes, ws_min, ws_max, _ = scenes.example_scene()
    scene = geo.Scene(boxes, 0.0)
    f = oracle_finder(boxes, 0.0, ws_min, ws_max)
    q_start = np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, 0.0])      # boundplanner_with_mpc_example.py:20-26
    T = 5
    q0 = q_start[None] + 0.05 * np.arange(T)[:, None] * np.array([1, -1, 0.5, 1, 0, -1, 0.3])[None]
    qf = q0 + 0.02 * np.array([1, 1, -1, 0.5, 1, 0, 1.0])[None]
    A, b, coll = mpc_geometry.collision_sets(scene, q0, qf, ws_min, ws_max)
    assert A.shape == (T, 6, 15, 3) and b.shape == (T, 6, 15)
    for t in range(T):
        set_joints = []
        for i in range(6):
            pl, pf = ofk.fk_pos_col(q0[t], i), ofk.fk_pos_col(qf[t], i)
            a_c, b_c, c = f.find_set_collision_avoidance(pl, pf, limit_space=True, e_max=0.7)
            assert bool(c) == bool(coll[t, i])
            set_joints.append([a_c, b_c - ofk.COL_JOINT_SIZES[i]])
        sets_normed = normalize_set_size(set_joints, 15)
        for i in range(6):
            assert np.abs(A[t, i] - sets_normed[i][0]).max() < 1e-6
            assert np.abs(b[t, i] - sets_normed[i][1]).max() < 1e-6
