"""Oracle (TEST INFRASTRUCTURE): the rest of the set graph -- end-effector fit,
projection points, edge costs and the planner loop up to the planned set sequence.

Restates bound_planner/BoundPlanner/BoundPlanner.py:
  check_intersection  :745-772   (20 qpOASES feasibility QPs, problem
                                  optimization_functions.py:140-183, J = 0)
  add_edges           :789-896   (projection QP :842-864, problem
                                  optimization_functions.py:107-137; edge cost :865-892)
  plan_convex_set_path :174-584  (up to the converged shortest path; the final
                                  via-point NLP, :540-555, stays with Ipopt)
qpOASES is a third-party dependency absent here (casadi==3.6.7) -> PARITY
UNPINNED against it; a feasibility QP with zero objective is an LP feasibility
problem (solved with the reference's own HiGHS, scipy.optimize.linprog) and the
projection of a point onto a polytope is unique (solved by active-set
enumeration after redundancy removal).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog

from .convex_set_finder import min_norm_point_polytopes
from .reduce_ineqs import redundant_row_mask
from .set_graph import set_intersection


def rodrigues_matrix(omega, phi):
    """optimization_functions.py:83-104"""
    k = np.array([[0.0, -omega[2], omega[1]], [omega[2], 0.0, -omega[0]], [-omega[1], omega[0], 0.0]])
    return np.eye(3) + np.sin(phi) * k + (1 - np.cos(phi)) * k @ k


def check_intersection(a_set, b_set, l_ee, sample, omega_normed, omega_norm):
    """:745-772 -> (success, p_inside = [sample, omega_sample])"""
    b_c = b_set - 0.001
    p_inside = np.concatenate((sample, [0]))
    for i in range(20):
        omega_sample = i / 19
        l_eec = rodrigues_matrix(omega_normed, omega_norm * omega_sample) @ l_ee
        res = linprog(np.zeros(3), A_ub=np.vstack((a_set, a_set)), b_ub=np.concatenate((b_c, b_c - a_set @ l_eec)),
                      bounds=(None, None))
        if res.success:
            return True, np.concatenate((sample, [omega_sample]))
    return False, p_inside


def project_point(a_set, b_set, x_d):
    """argmin |x - x_d|^2 s.t. A x <= b  (:842-864)."""
    keep = ~redundant_row_mask(a_set, b_set)
    A, b = a_set[keep], b_set[keep]
    z = min_norm_point_polytopes(A[None], (b - A @ x_d)[None])[0]
    return x_d + z


# ---- plan_convex_set_path's own tests (BoundPlanner.py:459-478, :505-512, :434) ------------------------------
def sample_flags(obs_sets, node_sets, sample):
    """(in_collision, in_safe) of one candidate, the reference's loops verbatim (:467-476)."""
    in_collision = in_safe = False
    for ob in obs_sets:
        if np.max(ob[0] @ sample - ob[1]) < 1e-3:
            in_collision = True
            break
    for a_set, b_set in node_sets:
        if np.max(a_set @ sample - b_set) < 1e-3:
            in_safe = True
            break
    return in_collision, in_safe


def first_free_sample(obs_sets, node_sets, candidates):
    """Index of the first candidate the rejection loop would accept, or -1."""
    for k, c in enumerate(candidates):
        coll, safe = sample_flags(obs_sets, node_sets, c)
        if not coll and not safe:
            return k
    return -1


def dedupe_distance(q_ellipse, p_mid, nodes):
    """dvertex of :505-510; nodes = [(q_ellipse_v, p_mid_v), ...]."""
    dvertex = np.inf
    for qv, pv in nodes:
        d = np.linalg.norm(q_ellipse - qv) + np.linalg.norm(p_mid - pv)
        dvertex = min(dvertex, d)
    return dvertex


def shortest_path(n_nodes, edges, src, dst):
    """The reference's own call (:434): networkx shortest_path with weights.  edges = [(u, v, w), ...]."""
    import networkx as nx

    g = nx.Graph()
    g.add_nodes_from(range(n_nodes))
    for u, v, w in edges:
        g.add_edge(u, v, weight=w)
    try:
        path = nx.shortest_path(g, src, dst, weight="weight")
    except nx.NetworkXNoPath:
        return None, np.inf
    return path, nx.path_weight(g, path, weight="weight")
