"""Shared helpers for the parity tests."""
import numpy as np

from oracle.convex_set_finder import ConvexSetFinder as OracleFinder
from oracle.obstacles import obstacle_reps

# north_star: halfspace coefficients agree within 1e-6 relative (fp64)
RTOL = 1e-6


def oracle_finder(boxes, inflate, ws_min, ws_max, max_rows=None):
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    return OracleFinder(obs_sets, pts, ws_max, ws_min, max_rows=max_rows)


REORDERED = []      # (tag) of every comparison that only matched up to a permutation of the picked rows


def assert_rows_close(A, b, Ao, bo, tag=""):
    """Rows agree within RTOL, in order.  The picked rows (index >= 6) may come in a different ORDER when the
    greedy loop met a numerical tie: the obstacles an inscribed ellipsoid touches are all at distance exactly 1 in
    its metric (ConvexSetFinder.py:429-431), so which of them np.argmin returns first is decided by the last bits
    of the MVIE solve -- in the reference (OSQP / Clarabel tolerances) as much as here.  Such a permutation is
    accepted and recorded in REORDERED; the set of halfspaces must still be the same."""
    assert A.shape == Ao.shape, f"{tag}: row count {A.shape[0]} vs oracle {Ao.shape[0]}"
    scale = max(1.0, np.abs(bo).max())
    if np.abs(A - Ao).max() <= RTOL and np.abs(b - bo).max() <= RTOL * scale:
        return
    assert np.abs(A[:6] - Ao[:6]).max() <= RTOL and np.abs(b[:6] - bo[:6]).max() <= RTOL * scale, f"{tag}: init rows"
    free = list(range(6, Ao.shape[0]))
    for r in range(6, A.shape[0]):
        d = [max(np.abs(A[r] - Ao[k]).max(), abs(b[r] - bo[k]) / scale) for k in free]
        k = int(np.argmin(d))
        assert d[k] <= RTOL, f"{tag}: row {r} has no counterpart in the oracle's set (closest differs by {d[k]})"
        free.pop(k)
    REORDERED.append(tag)


class OracleBackend:
    """Planner-loop primitives served by the oracle (same interface as planner.GpuBackend)."""

    def __init__(self, obstacles, obs_size_increase, workspace_max, workspace_min):
        from oracle.obstacles import obstacle_reps

        self.obs_sets, pts, _ = obstacle_reps(obstacles, obs_size_increase)
        self.set_finder = OracleFinder(self.obs_sets, pts, list(workspace_max), list(workspace_min))

    def find_set_around_point(self, p, fixed_mid, optimize):
        return self.set_finder.find_set_around_point(p, fixed_mid=fixed_mid, optimize=optimize)

    def find_set_collision_avoidance(self, p0, p1, compute_ellipsoid):
        return self.set_finder.find_set_collision_avoidance(p0, p1, compute_ellipsoid)

    def reduce_ineqs(self, a_set, b_set):
        from oracle.reduce_ineqs import reduce_ineqs

        return reduce_ineqs(a_set, b_set)

    def set_intersection(self, set1, set2, tol):
        from oracle.set_graph import set_intersection

        return set_intersection(set1, set2, tol)

    def check_intersection(self, a_set, b_set, l_ee, sample, omega_normed, omega_norm):
        from oracle.planner_graph import check_intersection

        return check_intersection(a_set, b_set, l_ee, sample, omega_normed, omega_norm)

    def project(self, a_set, b_set, x_d, x0=None):
        from oracle.planner_graph import project_point

        return project_point(a_set, b_set, x_d)

    def execute(self, req):
        """Request protocol of boundplanner_b200/planner.py, answered by the oracle."""
        kind = req[0]
        if kind == "set_point":
            A, b, Q, p = self.find_set_around_point(req[1], req[2], req[3])
            Ar, br = self.reduce_ineqs(A, b)
            return A, b, Q, p, Ar, br
        if kind == "set_line":
            A, b, Q, p, coll = self.find_set_collision_avoidance(req[1], req[2], True)
            Ar, br = self.reduce_ineqs(A, b)
            return A, b, Q, p, coll, Ar, br
        if kind == "edges":
            l_ee, omega_normed, omega_norm = req[4]
            out = []
            for setc in req[1]:
                x, _, ok = self.set_intersection(setc, req[2], req[3])
                fits, via = False, None
                if ok:
                    a_set = np.concatenate((setc[0], req[2][0]))
                    b_set = np.concatenate((setc[1], req[2][1]))
                    fits, via = self.check_intersection(a_set, b_set, l_ee, x, omega_normed, omega_norm)
                out.append((x, ok, fits, via))
            return out
        if kind == "intersect_many":
            out = []
            for setc in req[1]:
                x, _, ok = self.set_intersection(setc, req[2], req[3])
                out.append((x, ok))
            return out
        if kind == "fit_many":
            l_ee, omega_normed, omega_norm = req[2]
            return [self.check_intersection(a, b, l_ee, s, omega_normed, omega_norm) for a, b, s in req[1]]
        if kind == "project":
            return self.project(req[1], req[2], req[3])
        if kind == "reduce":
            return self.reduce_ineqs(req[1], req[2])
        raise ValueError(kind)


class OracleDeviceLoopBackend(OracleBackend):
    """OracleBackend that also answers the device-loop requests of the planner (sample_set, shortest_path, set_point
    with the planner attached) -- with the REFERENCE'S host loops, so that the planner-side protocol (candidates
    drawn ahead in chunks, generator rewind) can be checked against the plain host loop without a GPU."""

    device_loop = True

    def execute(self, req):
        import networkx as nx

        kind = req[0]
        if kind == "sample_set":
            cand, optimize, pl = np.asarray(req[1], float).reshape(-1, 3), req[2], req[3]
            for i, c in enumerate(cand):                      # BoundPlanner.py:459-478, one candidate at a time
                if not pl._in_collision(c) and not pl._in_safe(c):
                    try:
                        A, b, Q, p = self.find_set_around_point(c, True, optimize)
                    except (RuntimeError, ValueError) as e:
                        return i, e
                    Ar, br = self.reduce_ineqs(A, b)
                    return i, (A, b, Q, p, Ar, br, pl._min_node_distance(Q, p))
            return -1, None
        if kind == "shortest_path":
            return nx.shortest_path(req[1], 0, 1, weight="weight")
        if kind == "set_point" and len(req) > 4:
            out = super().execute(req[:4])
            return out + (req[4]._min_node_distance(out[2], out[3]),)
        return super().execute(req)
