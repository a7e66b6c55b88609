"""Set-sequence planner: the reference's ``plan_convex_set_path`` up to the planned
set sequence, with every geometric primitive served by the GPU kernels.

Restates the control flow of bound_planner/BoundPlanner/BoundPlanner.py
(non-replanning branch):

  plan_convex_set_path  :174-584   start / end sets, sampling loop, dedupe, convergence test
  compute_via_points    :586-743   with_rot=False branch (the via points are the p_proj's)
  add_edges             :789-896   intersection nodes, projection points, edge costs (quirk Q6)

What is NOT here: the final via-point NLP with rotations (Ipopt, :540-555) and the
replanning branch (:231-276) -- both stay with the reference.  The output is what
precedes them: the shortest path through the intersection graph, the sets it
visits ("planned set sequence") and the position via points.

The geometric primitives come from a ``backend``; the default ``GpuBackend``
calls libbpgeo through the drop-in classes.  (The parity tests run the same loop
with an oracle-backed backend and compare the sequences.)
"""
from __future__ import annotations

import networkx as nx
import numpy as np
from scipy.spatial.transform import Rotation as R


class GpuBackend:
    """Primitives of the planner loop, batch-of-one calls into the kernels."""

    def __init__(self, obstacles, obs_size_increase, workspace_max, workspace_min):
        from . import geometry as geo
        from .convex_set_finder import ConvexSetFinder
        from .set_graph import pack_sets

        self._geo = geo
        self._pack = pack_sets
        boxes = np.asarray(obstacles, float).reshape(-1, 6)
        box = np.concatenate((np.eye(3), -np.eye(3)))
        # add_obstacle_reps (:131-152): inflated boxes padded to 15 rows
        self.obs_sets = []
        for ob in boxes:
            a = np.zeros((15, 3))
            b = 10.0 * np.ones(15)
            a[:6] = box
            b[:6] = np.concatenate((ob[3:], -ob[:3])) + obs_size_increase
            self.obs_sets.append([a, b])
        self.set_finder = ConvexSetFinder(self.obs_sets, [None] * len(self.obs_sets), workspace_max, workspace_min)

    def find_set_around_point(self, p, fixed_mid, optimize):
        return self.set_finder.find_set_around_point(p, fixed_mid=fixed_mid, optimize=optimize)

    def find_set_collision_avoidance(self, p0, p1, compute_ellipsoid):
        return self.set_finder.find_set_collision_avoidance(p0, p1, compute_ellipsoid)

    def reduce_ineqs(self, a_set, b_set):
        from .utils import reduce_ineqs

        return reduce_ineqs(a_set, b_set)

    def set_intersection(self, set1, set2, tol):
        from .set_graph import set_intersection

        return set_intersection(set1, set2, tol)

    def _one_set(self, a_set, b_set):
        import torch

        A, b, m = self._pack([[a_set, b_set]])
        return torch.as_tensor(A).cuda(), torch.as_tensor(b).cuda(), torch.as_tensor(m).cuda()

    def check_intersection(self, a_set, b_set, l_ee, sample, omega_normed, omega_norm):
        A, b, m = self._one_set(a_set, b_set)
        fits, omega = self._geo.check_fit(A, b, m, np.array([[0, 0]], np.int32), l_ee, omega_normed, omega_norm,
                                          x0=np.asarray(sample, float)[None])
        ok = bool(fits.item())
        return ok, np.concatenate((sample, [float(omega.item()) if ok else 0.0]))

    def project(self, a_set, b_set, x_d, x0=None):
        A, b, m = self._one_set(a_set, b_set)
        x, _ = self._geo.project_points(A, b, m, np.array([[0, 0]], np.int32), np.asarray(x_d, float)[None])
        return x[0].cpu().numpy()


class SetSequencePlanner:
    def __init__(self, obstacles=(), obs_size_increase=0.08, workspace_max=(1.0, 1.0, 1.2),
                 workspace_min=(-1.0, -1.0, 0.0), backend=None, rng=None):
        self.obs_size_increase = obs_size_increase
        self.workspace_max = list(workspace_max)
        self.workspace_min = list(workspace_min)
        # :47-58
        self.w_size = 0.1
        self.c_fit = 1.0
        self.w_bias = 0.01
        self.max_set_size = 20
        self.length_ee = 0.05
        self.max_iters = 20
        self.nr_optimized = 10
        self.nr_free_mid = 5
        self.max_samples = 500
        self.rng = rng if rng is not None else np.random.default_rng()      # unseeded in the reference (quirk Q4)
        self.backend = backend if backend is not None else GpuBackend(obstacles, obs_size_increase, workspace_max,
                                                                      workspace_min)
        self.obs_sets = self.backend.obs_sets
        self.verbose = False

    # ---- :789-896 -----------------------------------------------------------
    def add_edges(self, id_new, graph, inter_graph, end, start):
        connected = False
        set_new = graph.nodes[id_new]["cset"]
        for vertex in list(graph.nodes.items()):
            if vertex[0] != id_new:
                setc = vertex[1]["cset"]
                idc = vertex[0]
                p_intersect, set_inter, intersects = self.backend.set_intersection(setc, set_new, 0.01)
            else:
                intersects = False
            if not intersects:
                continue
            fits, via = self.backend.check_intersection(set_inter[0], set_inter[1], self.l_ee, p_intersect,
                                                        self.omega_normed, self.omega_norm)
            self.id_inter += 1
            inter_graph.add_node(self.id_inter, cset=set_inter, id0=idc, id1=id_new, conn_to_start=False,
                                 conn_to_end=False, p_proj=None, p_via=via, fits=fits)
            self.nr_inter_set += 2
            for edge in list(inter_graph.nodes.items()):
                v0, v1 = edge[1]["id0"], edge[1]["id1"]
                cond1 = v0 == vertex[0] or v1 == vertex[0]
                cond2 = v0 == id_new or v1 == id_new
                if cond1:
                    size = vertex[1]["size"]
                elif cond2:
                    size = graph.nodes[id_new]["size"]
                if self.id_inter != edge[0] and (cond1 or cond2):
                    self.nr_edges += 2
                    p_proj = edge[1]["p_proj"]
                    if p_proj is None:
                        p_proj = end
                    me = inter_graph.nodes[self.id_inter]
                    if me["p_proj"] is None:
                        me["p_proj"] = self.backend.project(set_inter[0], set_inter[1], p_proj, p_intersect)
                    dist = np.linalg.norm(me["p_proj"] - p_proj)
                    conn_to_start = me["conn_to_start"] or edge[1]["conn_to_start"]
                    conn_to_end = me["conn_to_end"] or edge[1]["conn_to_end"]
                    me["conn_to_start"] = conn_to_start
                    me["conn_to_end"] = conn_to_end
                    edge[1]["conn_to_start"] = conn_to_start
                    edge[1]["conn_to_end"] = conn_to_end
                    connected = bool(conn_to_start and conn_to_end)          # last edge wins (quirk Q6)
                    c_size = np.tanh(0.25 - np.cbrt(size))
                    cost = dist * (1 + self.w_size * c_size) + self.w_bias
                    if not fits:
                        cost += self.c_fit
                    inter_graph.add_edge(self.id_inter, edge[0], weight=cost)
        return connected

    # ---- :586-743 (with_rot=False) ---------------------------------------------
    def compute_via_points(self, path, start, end, graph, inter_graph):
        x0 = np.empty(0)
        sets_inter = []
        for edge in path[1:-1]:
            sets_inter.append(inter_graph.nodes[edge]["cset"])
            x0 = np.concatenate((x0, inter_graph.nodes[edge]["p_proj"], [0.5]))
            idx = np.linalg.norm(sets_inter[-1][0], axis=1) > 1e-4
            sets_inter[-1][1][idx] -= 0.001                    # in-place, every call (quirk Q7)
        sets, seq = [], []
        last_id = None
        for i in range(len(path)):
            node = inter_graph.nodes[path[i]]
            if i == 0:
                last_id = node["id0"]
            else:
                id0, id1 = node["id0"], node["id1"]
                if id0 != last_id:
                    last_id = id0
                elif id1 != last_id:
                    last_id = id1
            sets.append(graph.nodes[last_id]["cset"])
            seq.append(last_id)
        sets_via, seq_via = [], []
        p_via = [start]
        for i in range(len(sets_inter)):
            p_via_opt = x0[4 * i: 4 * i + 3]
            if np.linalg.norm(p_via_opt - p_via[-1]) > 1e-4:
                p_via.append(p_via_opt)
                sets_via.append(sets[i])
                seq_via.append(seq[i])
        p_via.append(end)
        sets_via.append(sets[-1])
        seq_via.append(seq[-1])
        return np.array(p_via), p_via, sets_via, seq_via

    # ---- :174-534 -------------------------------------------------------------
    def plan_set_sequence(self, start, end, r0, r1, first_sample=None):
        """Returns dict(path, set_ids, sets_via, p_via, graph, inter_graph)."""
        start = np.array(start, float)
        end = np.array(end, float)
        sampled_first = False
        for ob in self.obs_sets:                                  # :199-204
            viol = ob[0] @ end - ob[1]
            if not np.any(viol > 0):
                idx = np.argmax(viol)
                end -= (viol[idx] - self.obs_size_increase) * ob[0][idx, :]
        self.omega = R.from_matrix(r1 @ r0.T).as_rotvec()        # :207-219
        self.omega_norm = np.linalg.norm(self.omega)
        self.omega_normed = self.omega / self.omega_norm if self.omega_norm > 1e-6 else np.array([0, 0, 1.0])
        self.l_ee = r0 @ np.array([-self.length_ee, 0, 0])
        self.l_ee_end = r1 @ np.array([-self.length_ee, 0, 0])
        graph, inter_graph = nx.Graph(), nx.Graph()
        self.nr_sets = self.nr_edges = self.nr_inter_set = 0

        a_set, b_set, q_start, p_mid_start = self.backend.find_set_around_point(start, True, True)   # :278-283
        collision = False
        if np.max(a_set @ (start + self.l_ee) - b_set) > 1e-8:
            a_set, b_set, q_start, p_mid_start, collision = self.backend.find_set_collision_avoidance(
                start, start + self.l_ee, True)
        if collision:
            raise RuntimeError("start point in collision (the replanning fallbacks of :296-324 are not restated)")
        a_set, b_set = self.backend.reduce_ineqs(a_set, b_set)
        set_start = [a_set, b_set]
        self.id_inter = 0
        self.id_graph = 0
        graph.add_node(0, cset=set_start, size=1 / np.linalg.det(q_start), q_ellipse=q_start, p_mid=p_mid_start,
                       a_set=np.array(a_set), b_set=np.array(b_set))
        inter_graph.add_node(0, cset=set_start, id0=0, id1=0, conn_to_start=True, conn_to_end=False, p_proj=start,
                             p_via=np.concatenate((start, [0.0])), fits=True)
        self.nr_sets += 1
        connected = self.add_edges(0, graph, inter_graph, end, start)
        if np.max(a_set @ end - b_set) < 1e-8 and np.max(a_set @ (end + self.l_ee_end) - b_set) < 1e-8:   # :361-375
            return dict(path=[0], set_ids=[0], sets_via=[set_start], p_via=np.array([start, end]), graph=graph,
                        inter_graph=inter_graph)
        a_set, b_set, q_end, p_mid_end, collision = self.backend.find_set_collision_avoidance(
            end, end + self.l_ee_end, True)                                                     # :381-389
        a_set, b_set = self.backend.reduce_ineqs(a_set, b_set)
        set_end = [a_set, b_set]
        self.id_graph += 1
        self.id_inter += 1
        graph.add_node(1, cset=set_end, size=1 / np.linalg.det(q_end), q_ellipse=q_end, p_mid=p_mid_end,
                       a_set=np.array(a_set), b_set=np.array(b_set))
        inter_graph.add_node(1, cset=set_end, id0=1, id1=1, conn_to_start=False, conn_to_end=True, p_proj=end,
                             p_via=np.concatenate((end, [1.0])), fits=True)
        self.nr_sets += 1
        conn = self.add_edges(1, graph, inter_graph, end, start)
        connected = conn or connected

        j = 0
        nr_samples = 0
        p_via_old = None
        path = None
        while True:                                               # :430-534
            via_sample = False
            if connected:
                path = nx.shortest_path(inter_graph, 0, 1, weight="weight")
                p_via, p_via_list, sets_via, seq_via = self.compute_via_points(path, start, end, graph, inter_graph)
                if p_via_old is not None and p_via_old.shape == p_via.shape and \
                        np.linalg.norm(p_via_old - p_via) < 1e-4:
                    break                                         # "Found path solution"
                samples = p_via_list[1:-1]
                via_sample = True
                p_via_old = np.copy(p_via)
            elif not sampled_first and first_sample is not None:
                samples = [first_sample]
            else:
                in_collision = in_safe = True
                nr_sampled = 0
                while (in_collision or in_safe) and nr_sampled <= self.max_samples:
                    in_collision = in_safe = False
                    sample = self.rng.uniform(self.workspace_min, self.workspace_max, 3)
                    nr_sampled += 1
                    for ob in self.obs_sets:
                        if np.max(ob[0] @ sample - ob[1]) < 1e-3:
                            in_collision = True
                            break
                    for setc in graph.nodes.items():
                        if np.max(setc[1]["a_set"] @ sample - setc[1]["b_set"]) < 1e-3:
                            in_safe = True
                            break
                if nr_sampled >= self.max_samples:
                    raise RuntimeError("(PosPath) Could not find collision-free sample")
                samples = [sample]
                nr_samples += 1
                if nr_samples > self.max_iters:
                    raise RuntimeError("(PosPath) Exceeded max iterations")
            for sample in samples:
                j += 1
                optimize = not (nr_samples >= self.nr_optimized)
                # fixed_mid = (via_sample or (not sampled_first),) is a 1-tuple: always truthy (quirk Q3)
                a_set, b_set, q_ellipse, p_mid = self.backend.find_set_around_point(np.asarray(sample, float), True,
                                                                                    optimize)
                a_set, b_set = self.backend.reduce_ineqs(a_set, b_set)
                sampled_first = True
                dvertex = np.inf
                for vertex in graph.nodes.items():
                    d = np.linalg.norm(q_ellipse - vertex[1]["q_ellipse"]) + np.linalg.norm(p_mid - vertex[1]["p_mid"])
                    dvertex = min(dvertex, d)
                if dvertex > 0.01:
                    self.id_graph += 1
                    graph.add_node(self.id_graph, cset=[a_set, b_set], size=1 / np.linalg.det(q_ellipse),
                                   q_ellipse=q_ellipse, p_mid=p_mid, a_set=np.array(a_set), b_set=np.array(b_set))
                    self.nr_sets += 1
                    conn = self.add_edges(self.id_graph, graph, inter_graph, end, start)
                    connected = conn or connected
        self.graph, self.inter_graph = graph, inter_graph
        return dict(path=path, set_ids=seq_via, sets_via=sets_via, p_via=p_via, graph=graph, inter_graph=inter_graph)
