"""Oracle (TEST INFRASTRUCTURE): redundancy removal of a convex set's rows.

Restates bound_planner/utils/util_functions.py:82-88 (``reduce_ineqs``), which
hands [b | -A] to pycddlib==3.0.2 ``matrix_redundancy_remove`` (cddlib's
dd_MatrixRedundancyRemove).  cddlib is a third-party dependency absent from
/root/reference and not installable here -> PARITY UNPINNED against cddlib
itself; its published algorithm is restated:

    for i = m down to 1:                      (rows are visited from the last to the first)
        solve  min  b_i - a_i.x   s.t.  a_j.x <= b_j  (j != i, j still present),  a_i.x <= b_i + 1
        if the optimum is >= 0: row i is redundant -> remove it on the spot

so a row that only touches the polytope (vertex / edge contact, optimum == 0) is
redundant, and of two identical rows the later one is removed.  The kept rows
keep their original order and their original (unnormalised) coefficients.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog


def redundant_row_mask(a_set, b_set, eps=1e-9):
    A = np.asarray(a_set, float)
    b = np.asarray(b_set, float)
    m = A.shape[0]
    keep = np.ones(m, bool)
    for i in range(m - 1, -1, -1):
        if not np.any(A[i]):
            keep[i] = b[i] < 0          # 0 <= b_i is always redundant
            continue
        others = keep.copy()
        others[i] = False
        res = linprog(-A[i], A_ub=np.vstack((A[others], A[i][None])), b_ub=np.concatenate((b[others], [b[i] + 1.0])),
                      bounds=[(None, None)] * 3)
        if res.status != 0:
            raise RuntimeError(f"redundancy LP failed: {res.message}")
        if b[i] - A[i] @ res.x >= -eps * max(1.0, abs(b[i])):
            keep[i] = False
    return ~keep


def reduce_ineqs(a_set, b_set):
    red = redundant_row_mask(a_set, b_set)
    return [np.asarray(a_set, float)[~red].copy(), np.asarray(b_set, float)[~red].copy()]
