"""Drop-in for the reference's ``ConvexSetFinder`` backed by libbpgeo.so.

Mirrors bound_planner/BoundPlanner/ConvexSetFinder.py:102-766 -- same
constructor, method names, argument meaning, return types (freshly allocated
float64 NumPy arrays) and error behaviour -- so that the unchanged planner
(``BoundPlanner.set_finder``, BoundPlanner.py:119-124) and BoundMPC
(BoundMPC.py:486-488) can use it in place of the CasADi/CVXPY implementation:

    planner.set_finder = boundplanner_b200.ConvexSetFinder(
        planner.obs_sets, planner.obs_points_sets, planner.workspace_max, planner.workspace_min)

Every method is a batch-of-one call into the batched device API in
``geometry.py``; use ``find_sets_around_points`` / ``find_sets_collision_avoidance``
for real batches.  Obstacles are normally the axis-aligned boxes the planner builds
(A = [I; -I], BoundPlanner.py:126-129); general convex polytopes (<= 15 rows, with
their vertices in ``obs_points_sets``) are supported by every method as well.
There is no CPU fallback.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import geometry as geo
from ._lib import (BP_MAX_ROWS, STATUS_ELLIPSE_VIOLATION, STATUS_MVIE_NO_INTERIOR, STATUS_MVIE_NOT_CONVERGED,
                   STATUS_OK, STATUS_ROW_CAP, STATUS_ROW_OVERFLOW)

_BOX = np.concatenate((np.eye(3), -np.eye(3)))


def boxes_from_obs_sets(obs_sets):
    """[A (15x3 padded), b] list as built by add_obstacle_reps -> [N,6] (lb | ub), already inflated."""
    boxes = np.empty((len(obs_sets), 6))
    for k, (a_set, b_set) in enumerate(obs_sets):
        a_set = np.asarray(a_set, float)
        b_set = np.asarray(b_set, float)
        if a_set.shape[0] < 6 or not np.array_equal(a_set[:6], _BOX) or np.any(a_set[6:] != 0.0):
            raise NotImplementedError("not an axis-aligned box (A = [I; -I], BoundPlanner.make_box): general polytope")
        boxes[k, :3] = -b_set[3:6]
        boxes[k, 3:] = b_set[:3]
    return boxes


class ConvexSetFinder:
    REFERENCE_MAX_ROWS = 20      # MVIE buffers of the reference (ConvexSetFinder.py:126-128, quirk Q5)

    def __init__(self, obs_sets, obs_points_sets, e_max, e_min, strict_rows=True):
        self.region = []
        self.rng = np.random.default_rng(0)          # :105
        self.ell_time = 0.0
        self.set_line_time = 0.0
        self.proj_time = 0.0
        self.e_max = e_max
        self.e_min = e_min
        self.max_iter = 5                             # :133
        # strict_rows=True reproduces the reference's ValueError on sets with more than 20 rows
        self.strict_rows = strict_rows
        self.verbose = False
        self._scene_obj = None
        self._scene_dirty = True
        self._polytopes = False
        self._obs_sets = []
        self._obs_points_sets = []
        self.obs_sets = obs_sets
        self.obs_points_sets = obs_points_sets
        self._scene                                   # upload the scene now: a missing GPU fails here

    # add_obstacle_reps(update=True) ASSIGNS these two attributes, obs_sets first and obs_points_sets second
    # (BoundPlanner.py:150-152).  Both setters only mark the device scene stale; it is (re)built on first use, from
    # whatever pair of lists is current then -- so the order of the two assignments does not matter, and a polytope
    # update never keeps the old vertex lists.
    @property
    def obs_sets(self):
        return self._obs_sets

    @obs_sets.setter
    def obs_sets(self, value):
        self._obs_sets = list(value).copy()
        self._scene_dirty = True

    @property
    def obs_points_sets(self):
        return self._obs_points_sets

    @obs_points_sets.setter
    def obs_points_sets(self, value):
        # for boxes the vertices are implied (8 corners); polytope scenes take their vertex lists from here
        self._obs_points_sets = list(value).copy() if value is not None else []
        self._scene_dirty = True

    @property
    def _scene(self):
        if not self._scene_dirty:
            return self._scene_obj
        try:
            boxes = boxes_from_obs_sets(self._obs_sets)
        except NotImplementedError:
            # general polytopes: the rows of obs_sets + the vertices of obs_points_sets
            if len(self._obs_points_sets) != len(self._obs_sets):
                raise ValueError(f"polytope obstacles need one vertex array per obstacle: {len(self._obs_sets)} "
                                 f"obs_sets but {len(self._obs_points_sets)} obs_points_sets") from None
            self._scene_obj = geo.PolytopeScene(self._obs_sets, self._obs_points_sets)
            self._polytopes = True
        else:
            self._polytopes = False
            if self._scene_obj is None or isinstance(self._scene_obj, geo.PolytopeScene):
                self._scene_obj = geo.Scene(boxes, 0.0)       # obs_sets are already inflated (:141)
            else:
                self._scene_obj.update(boxes, 0.0)
        self._scene_dirty = False
        return self._scene_obj

    # ---- errors ----------------------------------------------------------
    def _raise_for_status(self, status, m=None):
        if status == STATUS_ELLIPSE_VIOLATION:
            print("(Polyhedron) ERROR point is inside ellipse but should be outside.")
            raise RuntimeError("Ellipse violates constraints")           # :438
        if status == STATUS_ROW_OVERFLOW:
            raise ValueError(f"convex set needs more than {BP_MAX_ROWS} rows")
        if status == STATUS_ROW_CAP:
            # d2[: a_set.shape[0]] = b_set with more than 20 rows (:516)
            raise ValueError(
                f"could not broadcast input array from shape ({m},) into shape ({self.REFERENCE_MAX_ROWS},)")
        if status in (STATUS_MVIE_NO_INTERIOR, STATUS_MVIE_NOT_CONVERGED):
            raise RuntimeError(f"MVIE failed (status {status})")
        if self.strict_rows and m is not None and m > self.REFERENCE_MAX_ROWS:
            # d2[: a_set.shape[0]] = b_set with more than 20 rows (:516)
            raise ValueError(
                f"could not broadcast input array from shape ({m},) into shape ({self.REFERENCE_MAX_ROWS},)")

    def _ws(self):
        return np.asarray(self.e_min, float), np.asarray(self.e_max, float)

    # ---- :377-421 ---------------------------------------------------------
    def init_halfspaces(self):
        a_set_init, b_set_init = [], []
        for i in range(3):
            a_set_init.append(np.eye(3)[i, :])
            b_set_init.append(self.e_max[i])
            a_set_init.append(-np.eye(3)[i, :])
            b_set_init.append(-self.e_min[i])
        return a_set_init, b_set_init

    def init_halfspaces_point(self, p, e_max=0.3):
        a_set_init, b_set_init = [], []
        for i in range(3):
            a_set_init.append(np.eye(3)[i, :])
            b_set_init.append(p[i] + e_max)
            a_set_init.append(-np.eye(3)[i, :])
            b_set_init.append(-p[i] + e_max)
        return a_set_init, b_set_init

    # ---- :465-510 ---------------------------------------------------------
    def compute_set_projs(self, obs_sets, p0, ellipse_mat):
        start = time.perf_counter()
        scene = self._scene_for(obs_sets)
        y, _ = geo.closest_points(scene, np.asarray(p0, float)[None], np.asarray(ellipse_mat, float)[None])
        out = y[0].cpu().numpy()
        self.proj_time += time.perf_counter() - start
        return out

    def compute_set_projs_line(self, obs_sets, p0, p1):
        start = time.perf_counter()
        scene = self._scene_for(obs_sets)
        x, phi = geo.closest_points_line(scene, np.asarray(p0, float)[None], np.asarray(p1, float)[None])
        out = x[0].cpu().numpy(), phi[0].cpu().numpy()
        self.proj_time += time.perf_counter() - start
        return out

    def _scene_for(self, obs_sets):
        if obs_sets is self._obs_sets or len(obs_sets) == len(self._obs_sets) and all(
                a is b for a, b in zip(obs_sets, self._obs_sets)):
            return self._scene
        return geo.Scene(boxes_from_obs_sets(obs_sets), 0.0)

    # ---- :423-463 ---------------------------------------------------------
    def compute_polyhedron(self, q_inv, q_ellipse, p_seed, a_set_init, b_set_init):
        init = np.hstack((np.asarray(a_set_init, float), np.asarray(b_set_init, float)[:, None]))
        if init.shape != (6, 4):
            raise NotImplementedError("compute_polyhedron expects the 6 rows of init_halfspaces*")
        A, b, m, status = geo.polyhedron(self._scene, np.asarray(p_seed, float)[None], np.asarray(q_inv, float)[None],
                                         np.asarray(q_ellipse, float)[None], init[None])
        self._raise_for_status(int(status.item()))
        k = int(m.item())
        A, b = A[0, :k].cpu().numpy(), b[0, :k].cpu().numpy()
        return [A[i].copy() for i in range(k)], [float(b[i]) for i in range(k)]

    # ---- :512-562 ---------------------------------------------------------
    def _mvie(self, a_set, b_set, centre, free):
        a_set = np.asarray(a_set, float)
        b_set = np.asarray(b_set, float)
        m = a_set.shape[0]
        self._raise_for_status(STATUS_OK, m)
        if m > BP_MAX_ROWS:
            raise ValueError(f"convex set needs more than {BP_MAX_ROWS} rows")
        A = np.zeros((1, BP_MAX_ROWS, 3))
        b = np.full((1, BP_MAX_ROWS), 10.0)
        A[0, :m], b[0, :m] = a_set, b_set
        q_inv, _, c, status, _ = geo.mvie(A, b, np.array([m], np.int32), np.asarray(centre, float)[None], free)
        self._raise_for_status(int(status.item()))
        return q_inv[0].cpu().numpy(), c[0].cpu().numpy()

    def mvie_socp(self, a_set, b_set, p_hint=None):
        if p_hint is None:
            # no hint: the kernel finds an interior start itself (phase-I LP on the set's rows)
            p_hint = np.full(3, np.nan)
        return self._mvie(a_set, b_set, p_hint, True)

    def mvie_socp_fixed_mid(self, a_set, b_set, p_mid):
        q_new, _ = self._mvie(a_set, b_set, p_mid, False)
        return q_new, p_mid

    # ---- :564-588 ---------------------------------------------------------
    def mvie_socp_fixed_r(self, a_set, b_set, p_mid, r_ellipse, a_lb):
        a_set = np.asarray(a_set, float)
        b_set = np.asarray(b_set, float)
        m = a_set.shape[0]
        self._raise_for_status(STATUS_OK, m)
        if m > BP_MAX_ROWS:
            raise ValueError(f"convex set needs more than {BP_MAX_ROWS} rows")
        A = np.zeros((1, BP_MAX_ROWS, 3))
        b = np.full((1, BP_MAX_ROWS), 10.0)
        A[0, :m], b[0, :m] = a_set, b_set
        q_new, q_ell, eigs, status, _ = geo.mvie_fixed_r(A, b, np.array([m], np.int32), np.asarray(p_mid, float)[None],
                                                         np.asarray(r_ellipse, float)[None], np.array([float(a_lb)]))
        self._raise_for_status(int(status.item()))
        return q_new[0].cpu().numpy(), q_ell[0].cpu().numpy(), eigs[0].cpu().numpy()

    # ---- :190-240 ---------------------------------------------------------
    @staticmethod
    def _fetch_one(out, with_collision=False):
        """Everything a batch-of-one call hands back, in ONE device-to-host copy: rows, offsets, ellipsoid, centre and
        the integer outputs (as doubles) packed on the device first."""
        ints = [out.m, out.status, out.rows_peak if out.rows_peak is not None else out.m]
        if with_collision:
            ints.append(out.collision)
        flat = torch.cat([out.A[0].reshape(-1), out.b[0].reshape(-1), out.q_ellipse[0].reshape(-1),
                          out.p_mid[0].reshape(-1)] + [t[:1].to(torch.float64) for t in ints]).cpu().numpy()
        m_max = out.A.shape[1]
        o = 0
        A = flat[o: o + 3 * m_max].reshape(m_max, 3); o += 3 * m_max
        b = flat[o: o + m_max]; o += m_max
        q = flat[o: o + 9].reshape(3, 3).copy(); o += 9
        p = flat[o: o + 3].copy(); o += 3
        m, status, peak = int(flat[o]), int(flat[o + 1]), int(flat[o + 2])
        coll = bool(flat[o + 3]) if with_collision else False
        return A[:m].copy(), b[:m].copy(), q, p, m, status, peak, coll

    def find_set_around_point(self, p_seed, fixed_mid=False, optimize=True):
        out = self.find_sets_around_points(np.asarray(p_seed, float)[None], fixed_mid=fixed_mid, optimize=optimize)
        A, b, q, p, m, status, peak, _ = self._fetch_one(out)
        self._raise_for_status(status, peak if optimize else None)
        return A, b, q, p

    def find_sets_around_points(self, seeds, fixed_mid=False, optimize=True, m_max=BP_MAX_ROWS):
        """Batched form: S seeds -> geometry.SetBatch (device tensors)."""
        start = time.perf_counter()
        ws_min, ws_max = self._ws()
        # the reference overflows its 20-row MVIE buffers in whichever pass first exceeds them (quirk Q5)
        out = geo.build_sets_point(self._scene, seeds, ws_min, ws_max, fixed_mid=bool(fixed_mid),
                                   optimize=bool(optimize), max_iter=self.max_iter, m_max=m_max,
                                   row_cap=self.REFERENCE_MAX_ROWS if self.strict_rows else 0)
        torch.cuda.current_stream().synchronize()
        self.ell_time += time.perf_counter() - start      # projections and MVIE are fused on the device
        return out

    # ---- :242-307 (the reference planner's call is commented out, BoundPlanner.py:378-380) ----
    def find_set_around_line(self, p0, dp1, optimize=True):
        out = self.find_sets_around_lines(np.asarray(p0, float)[None], np.asarray(dp1, float)[None], optimize=optimize)
        status = int(out.status.item())
        m = int(out.m.item())
        self._raise_for_status(status, int(out.rows_peak.item()))
        A, b = out.A[0, :m].cpu().numpy(), out.b[0, :m].cpu().numpy()
        # the reference returns the Python lists of compute_polyhedron here (:307)
        return ([A[i].copy() for i in range(m)], [float(b[i]) for i in range(m)], out.q_ellipse[0].cpu().numpy(),
                out.p_mid[0].cpu().numpy())

    def find_sets_around_lines(self, p0, dp1, optimize=True, m_max=BP_MAX_ROWS):
        """Batched form: S segments -> geometry.SetBatch (device tensors)."""
        start = time.perf_counter()
        ws_min, ws_max = self._ws()
        out = geo.build_sets_around_line(self._scene, p0, dp1, ws_min, ws_max, optimize=bool(optimize),
                                         max_iter=self.max_iter, m_max=m_max,
                                         row_cap=self.REFERENCE_MAX_ROWS if self.strict_rows else 0)
        torch.cuda.current_stream().synchronize()
        self.ell_time += time.perf_counter() - start
        return out

    # ---- :309-375 ---------------------------------------------------------
    def find_set_collision_avoidance(self, p0, p1, compute_ellipsoid=False, limit_space=False, e_max=0.3):
        out = self.find_sets_collision_avoidance(np.asarray(p0, float)[None], np.asarray(p1, float)[None],
                                                 compute_ellipsoid, limit_space, e_max)
        A, b, q, p, m, status, _, collision = self._fetch_one(out, with_collision=True)
        if collision:
            print("(LineSet) [WARNING] Line is touching an obstacle")       # :337
        self._raise_for_status(status, m if compute_ellipsoid else None)
        if compute_ellipsoid:
            return A, b, q, p, collision
        return A, b, collision

    def find_sets_collision_avoidance(self, p0, p1, compute_ellipsoid=False, limit_space=False, e_max=0.3,
                                      m_max=BP_MAX_ROWS):
        start = time.perf_counter()
        ws_min, ws_max = self._ws()
        out = geo.build_sets_line(self._scene, p0, p1, ws_min, ws_max, compute_ellipsoid=bool(compute_ellipsoid),
                                  limit_space=bool(limit_space), e_max=float(e_max), m_max=m_max)
        torch.cuda.current_stream().synchronize()
        self.proj_time += time.perf_counter() - start
        return out
