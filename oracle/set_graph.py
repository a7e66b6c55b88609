"""Oracle (TEST INFRASTRUCTURE): pairwise set-intersection test of the set graph.

``set_intersection`` executes the reference's own solver call verbatim
(bound_planner/BoundPlanner/BoundPlanner.py:774-787: scipy.optimize.linprog,
HiGHS, c = 0, free bounds), so this row of the hot path is PINNED against the
real third-party solver, run in this container (scipy 1.18.1; the reference
pins 1.13.1, requirements.txt:3).  ``add_edges`` calls it with tol = 0.01
(:796-798).

``intersection_margin`` is the exact quantity behind the yes/no answer,
  s* = min_x max_i (a_i.x - (b_i - tol)) / ||a_i||   over rows with a_i != 0,
intersects <=> s* <= 0.  HiGHS accepts primal infeasibilities up to 1e-7, so
pairs with |s*| below ~1e-6 are "near ties" on which the reference's own answer
depends on solver tolerances; the parity tests report and exclude them.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog


def set_intersection(set1, set2, tol=0.0):
    set_inter = [
        np.concatenate((set1[0], set2[0])),
        np.concatenate((set1[1], set2[1])),
    ]
    sol_lin = linprog(
        np.zeros(3),
        A_ub=set_inter[0],
        b_ub=set_inter[1] - tol,
        bounds=(None, None),
    )
    return sol_lin.x, set_inter, sol_lin.success


def intersection_margin(set1, set2, tol=0.0):
    """Exact Chebyshev-type margin s* (see module docstring) via a 4-variable LP."""
    A = np.concatenate((set1[0], set2[0]))
    b = np.concatenate((set1[1], set2[1])) - tol
    nrm = np.linalg.norm(A, axis=1)
    nz = nrm > 0
    if np.any(b[~nz] < 0):
        return np.inf
    A, b, nrm = A[nz], b[nz], nrm[nz]
    res = linprog(
        np.array([0, 0, 0, 1.0]),
        A_ub=np.hstack((A, -nrm[:, None])),
        b_ub=b,
        bounds=[(None, None)] * 4,
    )
    if res.status == 3:      # unbounded below: trivially intersecting
        return -np.inf
    if not res.success:
        raise RuntimeError(f"margin LP failed: {res.message}")
    return float(res.x[3])


def adjacency(sets, tol=0.01):
    """Upper-triangular boolean adjacency of all set pairs, one reference call
    per pair, in the order add_edges would visit them (BoundPlanner.py:792-798)."""
    n = len(sets)
    adj = np.zeros((n, n), bool)
    for j in range(n):
        for i in range(j):
            adj[i, j] = adj[j, i] = bool(set_intersection(sets[i], sets[j], tol)[2])
    return adj
