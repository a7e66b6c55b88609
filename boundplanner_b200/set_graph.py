"""Set-graph side of the hot path: drop-in ``set_intersection`` plus the batched
all-pairs adjacency the planner's ``add_edges`` loop is replaced by.

Mirrors BoundPlanner.set_intersection (bound_planner/BoundPlanner/BoundPlanner.py:774-787)
and the way add_edges calls it (tol = 0.01, :796-798).  To use it in the
unchanged planner:

    planner.set_intersection = boundplanner_b200.set_intersection
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as geo
from ._lib import BP_MAX_ROWS


def pack_sets(sets, m_max=None):
    """Host list of [A (m,3), b (m,)] -> (A [S,m_max,3], b [S,m_max], m [S]) padded with
    A = 0, b = 10 like normalize_set_size (util_functions.py:119-133)."""
    ms = [np.asarray(s[0]).shape[0] for s in sets]
    if m_max is None:
        m_max = max(ms + [1])
    if m_max > BP_MAX_ROWS:
        raise ValueError(f"a set has more than {BP_MAX_ROWS} rows")
    A = np.zeros((len(sets), m_max, 3))
    b = np.full((len(sets), m_max), 10.0)
    for k, s in enumerate(sets):
        A[k, : ms[k]] = s[0]
        b[k, : ms[k]] = s[1]
    return A, b, np.asarray(ms, np.int32)


def set_intersection(set1, set2, tol=0.0):
    """Same signature and return as the reference: (point_inside or None, [A;A'],[b;b'], success)."""
    set_inter = [
        np.concatenate((set1[0], set2[0])),
        np.concatenate((set1[1], set2[1])),
    ]
    A, b, m = pack_sets([set1, set2])
    # one host-to-device copy (rows, offsets and row counts packed), one kernel chain, one copy back
    m_max = A.shape[1]
    packed = torch.as_tensor(np.concatenate((A.reshape(-1), b.reshape(-1), m.astype(np.float64)))).cuda()
    A_d = packed[: 6 * m_max].view(2, m_max, 3)
    b_d = packed[6 * m_max: 8 * m_max].view(2, m_max)
    m_d = packed[8 * m_max:].to(torch.int32)
    ok, x = geo.pairs_feasible_list(A_d, b_d, m_d, np.array([[0, 1]], np.int32), tol)
    res = torch.cat((ok.to(torch.float64), x.reshape(-1))).cpu().numpy()
    success = bool(res[0])
    point = res[1:4].copy() if success else None
    return point, set_inter, success


def adjacency(sets, tol=0.01):
    """Boolean [S,S] symmetric adjacency of host sets: adj[i,j] == set_intersection(sets[i], sets[j], tol)[2]."""
    A, b, m = pack_sets(sets)
    S = len(sets)
    bits = geo.pair_feasible(torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(), torch.as_tensor(m).cuda(), tol=tol)
    upper = geo.unpack_adjacency(bits, S).cpu().numpy()
    return upper | upper.T


def adjacency_from_batch(batch, tol=0.01):
    """Device path: geometry.SetBatch -> upper-triangular bool [S,S] tensor (no host round trip)."""
    bits = geo.pair_feasible(batch.A, batch.b, batch.m, tol=tol)
    return geo.unpack_adjacency(bits, batch.A.shape[0])
