"""Oracle (TEST INFRASTRUCTURE): obstacle representation.

Follows bound_planner/BoundPlanner/BoundPlanner.py:126-152 (``make_box``,
``add_obstacle_reps``) and bound_planner/utils/util_functions.py:119-133
(``normalize_set_size``), :66-79 (``compute_polytope_vertices`` -- pycddlib in
the reference; for the axis-aligned boxes the planner builds, the vertex set is
the 8 corners of [lb - inflate, ub + inflate], enumerated here directly).
"""
from __future__ import annotations

import itertools

import numpy as np


def make_box(lb, ub):
    """H-representation of a box: A = [I; -I], b = [ub; -lb] (BoundPlanner.py:126-129)."""
    a_set = np.concatenate((np.eye(3), -np.eye(3)))
    b_set = np.concatenate((np.asarray(ub, float), -np.asarray(lb, float)))
    return [a_set, b_set]


def normalize_set_size(sets, max_set_size=15):
    """Pad every [A, b] IN PLACE to max_set_size rows with A=0, b=10
    (util_functions.py:119-133; oversize sets are left ragged, with a print)."""
    for s in sets:
        m = s[0].shape[0]
        if m <= max_set_size:
            a_norm = np.zeros((max_set_size, 3))
            b_norm = 10.0 * np.ones(max_set_size)
            a_norm[:m] = s[0]
            b_norm[:m] = s[1]
            s[0], s[1] = a_norm, b_norm
        else:
            print(f"(SetNormalizer) ERROR set size {m} exceeds max set size {max_set_size}")
    return sets


def box_vertices(a_set, b_set):
    """8 corners of an axis-aligned box given as A=[I;-I], b=[ub;-lb]."""
    ub, lb = b_set[:3], -b_set[3:6]
    return np.array([[(lb, ub)[s][k] for k, s in enumerate(sel)]
                     for sel in itertools.product((0, 1), repeat=3)])


def obstacle_reps(obstacles, obs_size_increase=0.08):
    """Return (obs_sets padded to 15 rows, obs_points_sets (8x3 each), obs_sets_orig)
    exactly as BoundPlanner.add_obstacle_reps builds them (:131-152)."""
    obs_sets, obs_points_sets, obs_sets_orig = [], [], []
    for ob in obstacles:
        ob = np.asarray(ob, float)
        set_ob = make_box(ob[:3], ob[3:])
        adapted = [set_ob[0].copy(), set_ob[1] + obs_size_increase]
        obs_points_sets.append(box_vertices(adapted[0], adapted[1]))
        obs_sets_orig.append(set_ob)
        obs_sets.append(adapted)
    obs_sets = normalize_set_size(obs_sets)
    return obs_sets, obs_points_sets, obs_sets_orig
