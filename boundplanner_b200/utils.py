"""Drop-ins for the helpers of bound_planner/utils/util_functions.py that sit on the hot path."""
from __future__ import annotations

import numpy as np

from . import geometry as geo
from .set_graph import pack_sets


def normalize_set_size(sets, max_set_size=15):
    """Pad every [A, b] IN PLACE to max_set_size rows with A = 0, b = 10
    (util_functions.py:119-133; oversize sets are left ragged, with the reference's print)."""
    for s in sets:
        n = s[0].shape[0]
        if n <= max_set_size:
            a_norm = np.zeros((max_set_size, 3))
            b_norm = 10 * np.ones(max_set_size)
            a_norm[:n, :] = s[0]
            b_norm[:n] = s[1]
            s[0], s[1] = a_norm, b_norm
        else:
            print(f"(SetNormalizer) ERROR set size {n} exceeds max set size {max_set_size}")
    return sets


def reduce_ineqs(a_set, b_set):
    """Same signature and return as the reference (util_functions.py:82-88): [A_reduced, b_reduced]."""
    A, b, m = pack_sets([[np.asarray(a_set, float), np.asarray(b_set, float).reshape(-1)]])
    Ao, bo, mo, _, status = geo.reduce_ineqs(A, b, m)
    k = int(mo.item())
    return [Ao[0, :k].cpu().numpy(), bo[0, :k].cpu().numpy()]


def reduce_ineqs_batch(batch):
    """Device path: geometry.SetBatch -> (A_red, b_red, m_red, keep) CUDA tensors."""
    return geo.reduce_ineqs(batch.A, batch.b, batch.m)[:4]


def compute_polytope_vertices(a_set, b_set, vmax=64):
    """Same signature and return as the reference (util_functions.py:66-79): the list of vertices of
    {x : a_set x <= b_set}; ValueError("Polyhedron is not a polytope") for an unbounded (or empty) set.
    The order of the vertices is the kernel's (first row triple through each vertex), not cddlib's; every caller in
    the reference treats the list as a set (obstacle vertex tests ConvexSetFinder.py:449-451, plotting)."""
    from ._lib import STATUS_NOT_A_POLYTOPE, STATUS_ROW_OVERFLOW

    A, b, m = pack_sets([[np.asarray(a_set, float), np.asarray(b_set, float).reshape(-1)]])
    V, nv, status = geo.polytope_vertices(A, b, m, vmax=vmax)
    st = int(status.item())
    if st == STATUS_NOT_A_POLYTOPE:
        raise ValueError("Polyhedron is not a polytope")
    if st == STATUS_ROW_OVERFLOW:
        raise ValueError(f"polytope has more than {vmax} vertices")
    v = V[0, : int(nv.item())].cpu().numpy()
    return [v[i].copy() for i in range(v.shape[0])]


def obstacle_points_sets(obs_sets, vmax=64):
    """Vertex lists of a whole list of obstacle sets in one launch (what add_obstacle_reps builds per obstacle with
    cddlib, BoundPlanner.py:142): [[A, b], ...] -> [vertices (V x 3), ...]."""
    from ._lib import STATUS_OK

    A, b, m = pack_sets([[np.asarray(s[0], float), np.asarray(s[1], float).reshape(-1)] for s in obs_sets])
    V, nv, status = geo.polytope_vertices(A, b, m, vmax=vmax)
    V, nv, status = V.cpu().numpy(), nv.cpu().numpy(), status.cpu().numpy()
    if (status != STATUS_OK).any():
        bad = int(np.where(status != STATUS_OK)[0][0])
        raise ValueError(f"obstacle {bad}: " + ("Polyhedron is not a polytope" if status[bad] == 6
                                                else f"more than {vmax} vertices"))
    return [V[j, : nv[j]].copy() for j in range(len(obs_sets))]
