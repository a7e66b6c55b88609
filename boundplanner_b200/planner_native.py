"""Native lock-step planner driver: ``plan_batch`` with the per-query loop in C++ (csrc/bp_planner.h) instead of
Python generators, and every round's requests answered by one batched kernel chain inside libbpgeo.so
(``bp_plan_batch``; BASELINE config C3).

Same loop, same results as ``planner.plan_batch`` -- BoundPlanner.plan_convex_set_path (:174-584) up to the planned
set sequence, add_edges (:789-896), compute_via_points (:586-743, no rotations) -- but no Python between the
rounds: the queries' graphs, generators (numpy's PCG64 stream, restated) and bookkeeping live on the host in C++,
their graph nodes in device tables.  What Python still does is the per-query set-up that involves rotations
(scipy's ``as_rotvec`` and the 20 Rodrigues samples of check_intersection, :745-772) and the result objects.
"""
from __future__ import annotations

import ctypes

import numpy as np

MAX_NODES, NODE_ROWS, SET_ROWS, MAX_PATH, FIT_SAMPLES = 64, 24, 48, 64, 20
LENGTH_EE = 0.05            # BoundPlanner.__init__ (:54)

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_up = ctypes.POINTER(ctypes.c_ulonglong)


class BpPlanIn(ctypes.Structure):
    _fields_ = [("Q", ctypes.c_int), ("boxes", _dp), ("box_off", _ip), ("inflate", ctypes.c_double), ("ws_min", _dp),
                ("ws_max", _dp), ("starts", _dp), ("ends", _dp), ("l_ee", _dp), ("l_ee_end", _dp), ("ee_samples", _dp),
                ("rng", _up), ("has_first", _ip), ("first_sample", _dp), ("sample_chunk", ctypes.c_int),
                ("max_rounds", ctypes.c_int)]


class BpPlanOut(ctypes.Structure):
    _fields_ = [("err_kind", _ip), ("err_msg", ctypes.c_char_p), ("path", _ip), ("path_len", _ip), ("set_ids", _ip),
                ("n_ids", _ip), ("p_via", _dp), ("n_via", _ip), ("rng_out", _up), ("n_nodes", _ip), ("n_inter", _ip),
                ("n_edges", _ip), ("finish_round", _ip), ("finish_ms", _dp), ("node_A", _dp), ("node_b", _dp), ("node_m", _ip),
                ("stats", ctypes.POINTER(ctypes.c_longlong))]


def _rodrigues(omega, phi):
    k = np.array([[0.0, -omega[2], omega[1]], [omega[2], 0.0, -omega[0]], [-omega[1], omega[0], 0.0]])
    return np.eye(3) + np.sin(phi) * k + (1.0 - np.cos(phi)) * (k @ k)


def ee_setup(r0, r1, length_ee=LENGTH_EE):
    """plan_convex_set_path :207-219 for one (r0, r1): (l_ee, l_ee_end, the 20 rotated offsets of
    check_intersection)."""
    from scipy.spatial.transform import Rotation as R

    omega = R.from_matrix(r1 @ r0.T).as_rotvec()
    omega_norm = np.linalg.norm(omega)
    omega_normed = omega / omega_norm if omega_norm > 1e-6 else np.array([0, 0, 1.0])
    l_ee = r0 @ np.array([-length_ee, 0, 0])
    l_ee_end = r1 @ np.array([-length_ee, 0, 0])
    samples = np.ascontiguousarray([_rodrigues(omega_normed, omega_norm * k / (FIT_SAMPLES - 1)) @ l_ee
                                    for k in range(FIT_SAMPLES)])
    return l_ee, l_ee_end, samples


def rng_state_words(rng):
    """(state_hi, state_lo, inc_hi, inc_lo) of a numpy Generator over PCG64."""
    st = rng.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise ValueError("the native planner driver restates numpy's PCG64 stream only")
    s, inc = int(st["state"]["state"]), int(st["state"]["inc"])
    m = (1 << 64) - 1
    return [s >> 64, s & m, inc >> 64, inc & m]


class PackedQueries:
    """Host arrays of a batch in the layout of bp_plan_in / bp_plan_out (kept alive with the ctypes views)."""

    def __init__(self, queries, obs_size_increase, workspace_max, workspace_min, rng_seeds=None, sample_chunk=32,
                 want_nodes=True):
        Q = len(queries)
        self.Q = Q
        boxes = [np.asarray(q["obstacles"], float).reshape(-1, 6) for q in queries]
        self.box_off = np.zeros(Q + 1, np.int32)
        self.box_off[1:] = np.cumsum([b.shape[0] for b in boxes])
        self.boxes = np.ascontiguousarray(np.vstack(boxes)) if Q else np.zeros((0, 6))
        self.ws_min = np.ascontiguousarray(np.asarray(workspace_min, float).reshape(3))
        self.ws_max = np.ascontiguousarray(np.asarray(workspace_max, float).reshape(3))
        self.starts = np.ascontiguousarray([np.asarray(q["start"], float) for q in queries]).reshape(Q, 3)
        self.ends = np.ascontiguousarray([np.asarray(q["end"], float) for q in queries]).reshape(Q, 3)
        self.l_ee = np.zeros((Q, 3))
        self.l_ee_end = np.zeros((Q, 3))
        self.ee_samples = np.zeros((Q, FIT_SAMPLES, 3))
        cache = {}
        for i, q in enumerate(queries):
            r0, r1 = np.asarray(q["r0"], float), np.asarray(q["r1"], float)
            key = (r0.tobytes(), r1.tobytes())
            if key not in cache:
                cache[key] = ee_setup(r0, r1)
            self.l_ee[i], self.l_ee_end[i], self.ee_samples[i] = cache[key]
        self.rng = np.zeros((Q, 4), np.uint64)
        for i in range(Q):
            g = np.random.default_rng(rng_seeds[i] if rng_seeds is not None else None)
            self.rng[i] = rng_state_words(g)
        self.has_first = np.zeros(Q, np.int32)
        self.first_sample = np.zeros((Q, 3))
        for i, q in enumerate(queries):
            if q.get("first_sample") is not None:
                self.has_first[i] = 1
                self.first_sample[i] = np.asarray(q["first_sample"], float)
        self.inp = BpPlanIn(Q, self.boxes.ctypes.data_as(_dp), self.box_off.ctypes.data_as(_ip), float(obs_size_increase),
                            self.ws_min.ctypes.data_as(_dp), self.ws_max.ctypes.data_as(_dp),
                            self.starts.ctypes.data_as(_dp), self.ends.ctypes.data_as(_dp), self.l_ee.ctypes.data_as(_dp),
                            self.l_ee_end.ctypes.data_as(_dp), self.ee_samples.ctypes.data_as(_dp),
                            self.rng.ctypes.data_as(_up), self.has_first.ctypes.data_as(_ip),
                            self.first_sample.ctypes.data_as(_dp), int(sample_chunk), 0)
        # outputs
        self.err_kind = np.zeros(Q, np.int32)
        self.err_msg = ctypes.create_string_buffer(160 * max(Q, 1))
        self.path = np.full((Q, MAX_PATH), -1, np.int32)
        self.path_len = np.zeros(Q, np.int32)
        self.set_ids = np.full((Q, MAX_PATH), -1, np.int32)
        self.n_ids = np.zeros(Q, np.int32)
        self.p_via = np.zeros((Q, MAX_PATH + 2, 3))
        self.n_via = np.zeros(Q, np.int32)
        self.rng_out = np.zeros((Q, 4), np.uint64)
        self.n_nodes = np.zeros(Q, np.int32)
        self.n_inter = np.zeros(Q, np.int32)
        self.n_edges = np.zeros(Q, np.int32)
        self.finish_round = np.full(Q, -1, np.int32)
        self.finish_ms = np.zeros(Q)
        self.node_A = np.zeros((Q, MAX_NODES, NODE_ROWS, 3)) if want_nodes else None
        self.node_b = np.zeros((Q, MAX_NODES, NODE_ROWS)) if want_nodes else None
        self.node_m = np.zeros((Q, MAX_NODES), np.int32) if want_nodes else None
        self.stats = np.zeros(8, np.int64)
        self.out = BpPlanOut(self.err_kind.ctypes.data_as(_ip), ctypes.cast(self.err_msg, ctypes.c_char_p),
                             self.path.ctypes.data_as(_ip), self.path_len.ctypes.data_as(_ip),
                             self.set_ids.ctypes.data_as(_ip), self.n_ids.ctypes.data_as(_ip),
                             self.p_via.ctypes.data_as(_dp), self.n_via.ctypes.data_as(_ip),
                             self.rng_out.ctypes.data_as(_up), self.n_nodes.ctypes.data_as(_ip),
                             self.n_inter.ctypes.data_as(_ip), self.n_edges.ctypes.data_as(_ip),
                             self.finish_round.ctypes.data_as(_ip), self.finish_ms.ctypes.data_as(_dp),
                             self.node_A.ctypes.data_as(_dp) if want_nodes else None,
                             self.node_b.ctypes.data_as(_dp) if want_nodes else None,
                             self.node_m.ctypes.data_as(_ip) if want_nodes else None,
                             self.stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))

    def results(self):
        """Per query: dict(path, set_ids, sets_via, p_via, n_nodes, n_inter, n_edges) or the exception the Python
        planner raises for it."""
        out = []
        raw = self.err_msg.raw
        for q in range(self.Q):
            if self.err_kind[q]:
                msg = raw[160 * q: 160 * (q + 1)].split(b"\0", 1)[0].decode()
                out.append((RuntimeError if self.err_kind[q] == 1 else ValueError)(msg))
                continue
            ids = [int(v) for v in self.set_ids[q, : self.n_ids[q]]]
            res = dict(path=[int(v) for v in self.path[q, : self.path_len[q]]], set_ids=ids,
                       p_via=self.p_via[q, : self.n_via[q]].copy(), n_nodes=int(self.n_nodes[q]),
                       n_inter=int(self.n_inter[q]), n_edges=int(self.n_edges[q]))
            if self.node_A is not None:
                res["sets_via"] = [[self.node_A[q, s, : self.node_m[q, s]].copy(), self.node_b[q, s, : self.node_m[q, s]].copy()]
                                   for s in ids]
            out.append(res)
        return out

    def stats_dict(self):
        return dict(rounds=int(self.stats[0]), set_requests=int(self.stats[1]), pair_tests=int(self.stats[2]),
                    projections=int(self.stats[3]), shortest_paths=int(self.stats[4]),
                    kernel_chains=int(self.stats[5]), device_wait_ms=self.stats[6] / 1e3, lanes=int(self.stats[7]),
                    finish_round=self.finish_round.copy(), finish_ms=self.finish_ms.copy())


class NativePlanner:
    """A batch of independent planning queries, one scene each, planned in lock step by libbpgeo's native driver.
    The scene batch and the device tables are created once; ``run`` can be called again (same scenes)."""

    def __init__(self, queries, obs_size_increase=0.01, workspace_max=(1.0, 1.0, 1.2), workspace_min=(-1.0, -1.0, 0.0)):
        import torch

        from . import _lib
        from . import geometry as geo

        if not torch.cuda.is_available():
            raise _lib.BpGeoError("boundplanner_b200 needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.queries = queries
        self.infl, self.ws_max, self.ws_min = float(obs_size_increase), list(workspace_max), list(workspace_min)
        self.scene = geo.SceneBatch([q["obstacles"] for q in queries], obs_size_increase)
        self._h = ctypes.c_void_p(0)
        self._packed = {}
        _lib.check(self._lib.bp_plan_create(self.scene._h, len(queries), ctypes.byref(self._h)))

    def run(self, rng_seeds=None, sample_chunk=32, want_nodes=True):
        import torch

        from . import _lib

        # the packed host arrays of the batch (inputs incl. every query's generator state, output buffers) are kept
        # per seed list: planning the same batch again only runs the driver
        key = (None if rng_seeds is None else tuple(int(v) for v in rng_seeds), int(sample_chunk), bool(want_nodes))
        pk = self._packed.get(key) if rng_seeds is not None else None
        if pk is None:
            pk = PackedQueries(self.queries, self.infl, self.ws_max, self.ws_min, rng_seeds, sample_chunk, want_nodes)
            if rng_seeds is not None:
                self._packed = {key: pk}
        _lib.check(self._lib.bp_plan_run(self._h, ctypes.byref(pk.inp), ctypes.byref(pk.out),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return pk.results(), pk.stats_dict()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bp_plan_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_batch_native(queries, obs_size_increase=0.01, workspace_max=(1.0, 1.0, 1.2), workspace_min=(-1.0, -1.0, 0.0),
                      rng_seeds=None, sample_chunk=32):
    """Drop-in for ``planner.plan_batch``: (results, stats)."""
    pl = NativePlanner(queries, obs_size_increase, workspace_max, workspace_min)
    try:
        return pl.run(rng_seeds, sample_chunk)
    finally:
        pl.close()
