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
