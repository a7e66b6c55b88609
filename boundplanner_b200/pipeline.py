"""Set-graph pipeline with static buffers and CUDA-graph replay.

One "step" of the hot path (BASELINE.json configs[1]): S seeds -> S convex sets
(find_set_around_point) -> all-pairs intersection graph (set_intersection at
tol 0.01).  The kernel launches of a step (k_iris_fused, k_set_aabb, k_pair_filter,
k_pair_lp) are captured once into a CUDA graph; a step is then one H2D copy of the
seeds, one graph launch and one D2H copy of the results, with no per-step allocation
and no Python between kernels.

SetGraphPipeline / PipelinedSetGraph: one GPU (the latter keeps two steps in flight).
PeerSetGraphPipeline: multi-GPU, sets and adjacency rows exchanged by peer stores over
NVLink inside one CUDA graph per step.  ShardedSetGraphPipeline: the NCCL all-gather variant.
"""
from __future__ import annotations

import torch

from . import geometry as geo


class SetGraphPipeline:
    def __init__(self, scene, n_seeds, ws_min, ws_max, fixed_mid=True, optimize=True, max_iter=5, tol=0.01,
                 m_max=geo.BP_MAX_ROWS, use_cuda_graph=True, tail=None):
        import os

        # tail: the pair tests run in the tail of the set-build kernel (finished CTAs test their set against the
        # sets that arrive later) instead of as k_pair_filter / k_pair_lp launches; needs every CTA resident
        # (n_seeds <= 2 per SM).  Opt-in (BPGEO_TAIL=1): measured SLOWER than the separate pair stage on C2 (one B200:
        # 0.567 vs 0.525 ms per step; two: 0.730 vs 0.653 ms -- the sets of the slowest seed chains arrive together at
        # the very end, so most pair work cannot start early, and the tail runs it on 8 warps per SM instead of 16)
        if tail is None:
            tail = os.environ.get("BPGEO_TAIL", "0") == "1" and int(n_seeds) <= 2 * torch.cuda.get_device_properties(
                torch.cuda.current_device()).multi_processor_count
        self.tail = bool(tail)
        self.scene, self.S = scene, int(n_seeds)
        self.ws_min, self.ws_max = ws_min, ws_max
        self.kw = dict(fixed_mid=fixed_mid, optimize=optimize, max_iter=max_iter)
        self.tol = tol
        self.seeds_dev = torch.zeros((self.S, 3), dtype=torch.float64, device="cuda")
        # every result of a step lives in ONE device buffer (A | b | q_ellipse | p_mid | m | status | adjacency bits),
        # so that the end-to-end path copies it back with a single D2H transfer instead of seven
        S, words = self.S, (self.S + 31) // 32
        n64 = [S * m_max * 3, S * m_max, S * 9, S * 3]
        n32 = [S, S, S * words]
        self._out_dev = torch.zeros((sum(n64) * 8 + sum(n32) * 4,), dtype=torch.uint8, device="cuda")
        v64 = self._out_dev[: sum(n64) * 8].view(torch.float64)
        v32 = self._out_dev[sum(n64) * 8:].view(torch.int32)
        o64 = [0, n64[0], n64[0] + n64[1], n64[0] + n64[1] + n64[2]]
        A = v64[o64[0]: o64[0] + n64[0]].view(S, m_max, 3)
        b = v64[o64[1]: o64[1] + n64[1]].view(S, m_max)
        q = v64[o64[2]: o64[2] + n64[2]].view(S, 3, 3)
        p = v64[o64[3]: o64[3] + n64[3]].view(S, 3)
        b.fill_(10.0)                                   # normalize_set_size padding
        m = v32[:S]
        status = v32[S: 2 * S]
        self.bits = v32[2 * S:].view(S, words)
        own = geo.alloc_set_batch(self.S, m_max)        # iters / rows_peak / loop workspace
        self.batch = geo.SetBatch(A, b, m, q, p, status, iters=own.iters, rows_peak=own.rows_peak, work=own.work)
        self.pair_buf = (self.bits, geo.alloc_pair_buffers(self.S)[1])
        self.aabb = torch.empty((self.S, 6), dtype=torch.float64, device="cuda")   # written by the build's epilogue
        self._tail = None
        if self.tail:
            self._count = torch.zeros((1,), dtype=torch.int32, device="cuda")
            self._log = torch.zeros((self.S,), dtype=torch.int64, device="cuda")
            self._epoch = torch.zeros((1,), dtype=torch.int32, device="cuda")
            self._tail = geo.make_tail(A, b, m, self.aabb, self._count, self._log, self.bits, self._epoch, tol)
        self._views = (A, b, m, q, p, status, self.bits)
        self._host = None
        self._graph = None
        self._stream = torch.cuda.Stream()
        if use_cuda_graph:
            self._capture()

    def _enqueue(self):
        if self._tail is not None:
            geo.step_begin(self._tail, False)
            geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, aabb=self.aabb,
                                 tail=self._tail, **self.kw)
            return
        geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, aabb=self.aabb,
                             **self.kw)
        geo.pair_feasible(self.batch.A, self.batch.b, self.batch.m, self.tol, out=self.pair_buf, aabb=self.aabb)

    def _capture(self):
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(2):                      # warm-up: function attributes, lazy module load
                self._enqueue()
        self._stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._enqueue()
        self._graph = g

    def stage_times(self, reps=5):
        """Per-stage device time of one step in ms (median of `reps` eager steps, CUDA events between the kernels on
        the launch stream): k_iris_fused (with the bounding boxes from its epilogue; "aabb" is 0 then) |
        k_pair_filter | k_pair_lp."""
        import statistics

        if self._tail is not None:
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self._enqueue()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return {"build_and_pairs": statistics.median(ts), "aabb": 0.0, "filter": 0.0, "lp": 0.0}
        acc = {"build": [], "aabb": [], "filter": [], "lp": []}
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, aabb=self.aabb,
                                 **self.kw)
            e1.record()
            st = geo.pair_feasible_stages(self.batch.A, self.batch.b, self.batch.m, self.tol, out=self.pair_buf,
                                          aabb=self.aabb)
            acc["build"].append(e0.elapsed_time(e1))
            for k, v in st.items():
                acc[k].append(v)
        return {k: statistics.median(v) for k, v in acc.items()}

    def run_device(self):
        """Inputs already in self.seeds_dev; enqueue one step on the current stream."""
        if self._graph is not None:
            self._graph.replay()
        else:
            self._enqueue()
        return self.batch, self.bits

    def run(self, seeds_host_pinned):
        """End-to-end step from pinned host seeds [S,3]: H2D, step, D2H of every result
        (A, b, m, q_ellipse, p_mid, status, adjacency bits) into pinned host buffers."""
        if self._host is None:
            self._out_host = torch.empty(self._out_dev.shape, dtype=torch.uint8, pin_memory=True)
            # host views with the layout of the device buffer
            host = []
            for t in self._views:
                off = t.data_ptr() - self._out_dev.data_ptr()
                host.append(self._out_host[off: off + t.numel() * t.element_size()].view(t.dtype).view(t.shape))
            self._host = host
        self.seeds_dev.copy_(seeds_host_pinned, non_blocking=True)
        self.run_device()
        self._out_host.copy_(self._out_dev, non_blocking=True)      # one D2H transfer for all results
        torch.cuda.current_stream().synchronize()
        return self._host

    # ---- asynchronous form: submit() enqueues H2D + step + D2H and returns at once, wait() delivers the results ----
    def submit(self, seeds_host_pinned, copy_stream=None):
        """Enqueue one end-to-end step.  The D2H transfer goes to `copy_stream` (behind an event), so that the next
        step of ANOTHER pipeline object can start while this one's results travel back."""
        if self._host is None:
            self.run(seeds_host_pinned)                               # allocates the pinned mirror (synchronous once)
        cur = torch.cuda.current_stream()
        # a previous submit()'s D2H on the copy stream may still be reading the device buffers this step overwrites
        # (and the pinned mirror): order this step behind it
        pending = getattr(self, "_copy_done_for_compute", None)
        if pending is not None:
            cur.wait_event(pending)
            self._copy_done_for_compute = None
        self.seeds_dev.copy_(seeds_host_pinned, non_blocking=True)
        self.run_device()
        if copy_stream is None:
            self._out_host.copy_(self._out_dev, non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record(cur)
        else:
            ready = torch.cuda.Event()
            ready.record(cur)
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                self._out_host.copy_(self._out_dev, non_blocking=True)
                self._done = torch.cuda.Event()
                self._done.record(copy_stream)
            # the next step on `cur` overwrites the device buffers: it has to wait for this copy
            self._copy_done_for_compute = self._done

    def wait(self):
        """Block until the results of the last submit() are in the pinned host buffers; returns them."""
        self._done.synchronize()
        return self._host


class PipelinedSetGraph:
    """`depth` SetGraphPipeline objects used round-robin: while the results of step k travel back to the host on a
    copy stream, step k+1 already runs on the compute stream (its own buffers).  Every step still copies its seeds
    in and ALL its results out; only the host no longer idles on a synchronize per step."""

    def __init__(self, scene, n_seeds, ws_min, ws_max, depth=2, **kw):
        self.pipes = [SetGraphPipeline(scene, n_seeds, ws_min, ws_max, **kw) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream()
        self._k = 0
        self._busy = [False] * depth

    def take(self):
        """Results (pinned host views) of the step that used the next slot `depth` submissions ago, or None.  They
        stay valid until the following put(): consume them first."""
        i = self._k % len(self.pipes)
        if not self._busy[i]:
            return None
        self._busy[i] = False
        return self.pipes[i].wait()             # the slot's D2H has long finished: its buffers may be reused now

    def put(self, seeds_host_pinned):
        i = self._k % len(self.pipes)
        if self._busy[i]:
            raise RuntimeError("take() the slot's results before reusing it")
        self.pipes[i].submit(seeds_host_pinned, self.copy_stream)
        self._busy[i] = True
        self._k += 1

    def drain(self):
        """Results of the steps still in flight, oldest first."""
        out = []
        n = len(self.pipes)
        for j in range(n):
            i = (self._k + j) % n
            if self._busy[i]:
                out.append(self.pipes[i].wait())
                self._busy[i] = False
        return out


class ShardedSetGraphPipeline:
    """Multi-GPU step (one process per GPU, NCCL): every rank builds its S_loc seeds (CUDA graph 1),
    the packed halfspace tensors are all-gathered, every rank tests a balanced row block of the global
    pair matrix (CUDA graph 2 on the gathered static buffers) and the bit rows are all-gathered.
    Two collectives per step, nothing else crosses NVLink."""

    def __init__(self, scene, n_seeds_local, ws_min, ws_max, fixed_mid=True, optimize=True, max_iter=5, tol=0.01,
                 m_max=geo.BP_MAX_ROWS, group=None):
        import torch.distributed as dist

        from . import distributed as bpd

        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.scene, self.S_loc, self.S = scene, int(n_seeds_local), int(n_seeds_local) * self.world
        self.ws_min, self.ws_max, self.tol = ws_min, ws_max, tol
        self.kw = dict(fixed_mid=fixed_mid, optimize=optimize, max_iter=max_iter)
        dev = "cuda"
        self.seeds_dev = torch.zeros((self.S_loc, 3), dtype=torch.float64, device=dev)
        self.batch = geo.alloc_set_batch(self.S_loc, m_max)
        self.m_max = m_max
        # packed rows (a0,a1,a2,b) + three extra rows carrying the row count and the set's bounding box
        # (computed by its owner, so no rank recomputes all S boxes): one all-gather message per step
        self.packed = torch.zeros((self.S_loc, m_max + 3, 4), dtype=torch.float64, device=dev)
        self.gathered = torch.empty((self.S, m_max + 3, 4), dtype=torch.float64, device=dev)
        self.aabb_loc = torch.empty((self.S_loc, 6), dtype=torch.float64, device=dev)
        self.aabb_g = torch.empty((self.S, 6), dtype=torch.float64, device=dev)
        self.Ag = torch.empty((self.S, m_max, 3), dtype=torch.float64, device=dev)
        self.bg = torch.empty((self.S, m_max), dtype=torch.float64, device=dev)
        self.mg = torch.empty((self.S,), dtype=torch.int32, device=dev)
        self.blocks = bpd.balanced_row_blocks(self.S, self.world)
        self.r0, self.r1 = self.blocks[self.rank]
        self.max_rows = max(hi - lo for lo, hi in self.blocks)
        self.words = (self.S + 31) // 32
        self.pair_buf = geo.alloc_pair_buffers(self.S, self.r1 - self.r0)
        self.bits_padded = torch.zeros((self.max_rows, self.words), dtype=torch.int32, device=dev)
        self.bits_gathered = torch.empty((self.world * self.max_rows, self.words), dtype=torch.int32, device=dev)
        self._stream = torch.cuda.Stream()
        self._g1 = self._capture(self._build)
        self._g2 = None                      # captured after the first all-gather filled the static buffers

    def _build(self):
        geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, **self.kw)
        self.packed[:, : self.m_max, :3] = self.batch.A
        self.packed[:, : self.m_max, 3] = self.batch.b
        self.packed[:, self.m_max, :] = self.batch.m.to(torch.float64).unsqueeze(1)
        geo.set_aabb(self.batch.A, self.batch.b, self.batch.m, out=self.aabb_loc)
        self.packed[:, self.m_max + 1, :3] = self.aabb_loc[:, :3]
        self.packed[:, self.m_max + 2, :3] = self.aabb_loc[:, 3:]

    def _pairs(self):
        self.Ag.copy_(self.gathered[:, : self.m_max, :3])
        self.bg.copy_(self.gathered[:, : self.m_max, 3])
        self.mg.copy_(self.gathered[:, self.m_max, 0].to(torch.int32))
        self.aabb_g[:, :3] = self.gathered[:, self.m_max + 1, :3]
        self.aabb_g[:, 3:] = self.gathered[:, self.m_max + 2, :3]
        if self.r1 > self.r0:
            geo.pair_feasible(self.Ag, self.bg, self.mg, self.tol, self.r0, self.r1, out=self.pair_buf,
                              aabb=self.aabb_g)
            self.bits_padded[: self.r1 - self.r0].copy_(self.pair_buf[0])

    def _capture(self, fn):
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(2):
                fn()
        self._stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            fn()
        return g

    def run_device(self):
        self._g1.replay()
        self.dist.all_gather_into_tensor(self.gathered, self.packed, group=self.group)
        if self._g2 is None:
            torch.cuda.current_stream().synchronize()
            self._g2 = self._capture(self._pairs)
        self._g2.replay()
        self.dist.all_gather_into_tensor(self.bits_gathered, self.bits_padded, group=self.group)
        return self.batch, self.bits_gathered

    def adjacency_bits(self):
        """Global bit matrix [S, words] assembled from the gathered, padded row blocks."""
        return torch.cat([self.bits_gathered[r * self.max_rows: r * self.max_rows + (hi - lo)]
                          for r, (lo, hi) in enumerate(self.blocks)])


class PeerSetGraphPipeline:
    """Multi-GPU step with the exchange done by PEER STORES over NVLink instead of NCCL collectives.

    Every rank owns one symmetric allocation (torch symmetric memory: the same layout mapped into every
    process) holding the global tables  A[S,m_max,3] | b[S,m_max] | aabb[S,6] | m[S] | adjacency[S,words].
    A step is ONE CUDA graph per rank:

      k_iris_fused (own S_loc seeds; its epilogue computes each finished set's bounding box and stores rows,
      row count and box into the tables of EVERY rank through the peers' mapped addresses -- no pack /
      all-gather / unpack, no separate box / scatter kernels) -> signal-pad barrier -> pair kernels on the own row block of the global pair matrix, reading the
      local tables -> k_scatter_rows_peers: the adjacency rows go to every rank -> signal-pad barrier.

    The second barrier also protects the tables: no rank starts the next step's scatter before every rank
    has finished reading.  Same results as ShardedSetGraphPipeline (tools/check_sharded.py --peer)."""

    def __init__(self, scene, n_seeds_local, ws_min, ws_max, fixed_mid=True, optimize=True, max_iter=5, tol=0.01,
                 m_max=geo.BP_MAX_ROWS, group=None, tail=None):
        import os

        import numpy as np
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        # tail: pair tests in the tail of the set-build kernel, driven by per-set arrival flags in every rank's
        # tables (no barrier between build and pairs, no pair kernels); needs every CTA resident.  Opt-in
        # (BPGEO_TAIL=1), see SetGraphPipeline
        if tail is None:
            tail = os.environ.get("BPGEO_TAIL", "0") == "1" and int(n_seeds_local) <= 2 * torch.cuda.get_device_properties(
                torch.cuda.current_device()).multi_processor_count and int(n_seeds_local) * dist.get_world_size(group) <= 8192
        self.tail = bool(tail)

        from . import _lib
        from . import distributed as bpd

        self._lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.scene, self.S_loc, self.S = scene, int(n_seeds_local), int(n_seeds_local) * self.world
        self.ws_min, self.ws_max, self.tol, self.m_max = ws_min, ws_max, tol, m_max
        self.kw = dict(fixed_mid=fixed_mid, optimize=optimize, max_iter=max_iter)
        S, words = self.S, (self.S + 31) // 32
        self.words = words
        # layout of the symmetric allocation (bytes, 256-aligned regions)
        sizes = [S * m_max * 3 * 8, S * m_max * 8, S * 6 * 8, S * 4, 2 * S * words * 4, S * 8, 256]
        offs, o = [], 0
        for sz in sizes:
            offs.append(o)
            o += (sz + 255) // 256 * 256
        self.off_A, self.off_b, self.off_aabb, self.off_m, self.off_bits, self.off_log, self.off_count = offs
        dev = torch.device("cuda", torch.cuda.current_device())
        self.symm = symm_mem.empty((o,), dtype=torch.uint8, device=dev)
        self.symm.zero_()
        self.hdl = symm_mem.rendezvous(self.symm, group.group_name)
        self.peer_base = torch.as_tensor(np.asarray(self.hdl.buffer_ptrs, dtype=np.uint64).view(np.int64), device=dev)

        def view(off, nbytes, dtype, shape):
            return self.symm[off: off + nbytes].view(dtype).view(shape)

        self.Ag = view(self.off_A, sizes[0], torch.float64, (S, m_max, 3))
        self.bg = view(self.off_b, sizes[1], torch.float64, (S, m_max))
        self.aabb_g = view(self.off_aabb, sizes[2], torch.float64, (S, 6))
        self.mg = view(self.off_m, sizes[3], torch.int32, (S,))
        self.bits2_g = view(self.off_bits, sizes[4], torch.int32, (2, S, words))     # tail mode: alternate by epoch
        self.bits_g = self.bits2_g[0]                                                # classic mode: buffer 0
        self.log_g = view(self.off_log, sizes[5], torch.int64, (S,))
        self.count_g = view(self.off_count, 4, torch.int32, (1,))
        self._epoch = torch.zeros((1,), dtype=torch.int32, device=dev)
        self._tail = geo.make_tail(self.Ag, self.bg, self.mg, self.aabb_g, self.count_g, self.log_g, self.bits2_g,
                                   self._epoch, tol, off_count=self.off_count, off_log=self.off_log,
                                   off_bits=self.off_bits) if self.tail else None
        self.seeds_dev = torch.zeros((self.S_loc, 3), dtype=torch.float64, device=dev)
        self.batch = geo.alloc_set_batch(self.S_loc, m_max)
        self.aabb_loc = torch.empty((self.S_loc, 6), dtype=torch.float64, device=dev)
        self.blocks = bpd.balanced_row_blocks(S, self.world)
        self.r0, self.r1 = self.blocks[self.rank]
        self.pair_buf = geo.alloc_pair_buffers(S, max(self.r1 - self.r0, 1))
        self._stream = torch.cuda.Stream()
        self._graph = None
        self._steps = 0
        self._capture()
        self._steps = int(self._epoch.item())            # warm-up + capture runs have advanced the epoch

    def _peers(self):
        return dict(base=self.peer_base, world=self.world, slot0=self.rank * self.S_loc, off_A=self.off_A,
                    off_b=self.off_b, off_m=self.off_m, off_aabb=self.off_aabb)

    def _enqueue(self):
        lib, st = self._lib, geo._stream()
        if self._tail is not None:
            # one kernel per step and rank: set build, exchange (peer stores + arrival flags) and pair tests; the
            # only rank synchronisation is the barrier at the end of the step
            geo.step_begin(self._tail, True)
            geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, aabb=self.aabb_loc,
                                 peers=self._peers(), tail=self._tail, **self.kw)
            self.hdl.barrier(channel=1)
            return
        # the owner's stores of every finished set (rows, row count, bounding box) into all ranks' tables happen in
        # the epilogue of the set-build kernel: no separate box / scatter launches
        geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, aabb=self.aabb_loc,
                             peers=self._peers(), **self.kw)
        self.hdl.barrier(channel=0)
        rows = self.r1 - self.r0
        if rows > 0:
            bits = self.pair_buf[0][:rows]
            geo.pair_feasible(self.Ag, self.bg, self.mg, self.tol, self.r0, self.r1, out=(bits, self.pair_buf[1]),
                              aabb=self.aabb_g)
            geo.check(lib.bp_scatter_rows_peers(geo._ptr(bits), rows, self.words, self.r0, geo._ptr(self.peer_base),
                                                self.world, self.off_bits, st))
        self.hdl.barrier(channel=1)

    def stage_times(self, reps=5):
        """Per-stage device time of one step on THIS rank in ms (median of `reps` eager steps with CUDA events
        between the stages).  barrier0 = waiting for the slowest rank's sets, barrier1 = for its adjacency rows."""
        import statistics

        names = ["build", "aabb", "scatter_sets", "barrier0", "filter", "lp", "scatter_rows", "barrier1"]
        acc = {k: [] for k in names}
        lib, st = self._lib, geo._stream()
        if self._tail is not None:
            for _ in range(reps):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                self.hdl.barrier(channel=0)
                ev[0].record()
                geo.step_begin(self._tail, True)
                geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch,
                                     aabb=self.aabb_loc, peers=self._peers(), tail=self._tail, **self.kw)
                ev[1].record()
                self.hdl.barrier(channel=1)
                ev[2].record()
                torch.cuda.synchronize()
                acc["build"].append(ev[0].elapsed_time(ev[1]))
                acc["barrier1"].append(ev[1].elapsed_time(ev[2]))
            self._steps = int(self._epoch.item())
            return {"build_exchange_pairs": statistics.median(acc["build"]), "barrier1": statistics.median(acc["barrier1"])}
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            self.hdl.barrier(channel=0)                  # start the ranks together, like back-to-back steps do
            ev[0].record()
            geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch,
                                 aabb=self.aabb_loc, peers=self._peers(), **self.kw)
            ev[1].record()
            ev[2].record()                               # (boxes and peer stores: epilogue of the build kernel)
            ev[3].record()
            self.hdl.barrier(channel=0)
            ev[4].record()
            rows = self.r1 - self.r0
            pst = {"filter": 0.0, "lp": 0.0}
            bits = self.pair_buf[0][:max(rows, 0)]
            if rows > 0:
                pst = geo.pair_feasible_stages(self.Ag, self.bg, self.mg, self.tol, self.r0, self.r1,
                                               out=(bits, self.pair_buf[1]), aabb=self.aabb_g)
            ev[5].record()
            if rows > 0:
                geo.check(lib.bp_scatter_rows_peers(geo._ptr(bits), rows, self.words, self.r0, geo._ptr(self.peer_base),
                                                    self.world, self.off_bits, st))
            ev[6].record()
            self.hdl.barrier(channel=1)
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            torch.cuda.synchronize()
            acc["build"].append(ev[0].elapsed_time(ev[1]))
            acc["aabb"].append(ev[1].elapsed_time(ev[2]))
            acc["scatter_sets"].append(ev[2].elapsed_time(ev[3]))
            acc["barrier0"].append(ev[3].elapsed_time(ev[4]))
            acc["filter"].append(pst["filter"])
            acc["lp"].append(pst["lp"])
            acc["scatter_rows"].append(ev[5].elapsed_time(ev[6]))
            acc["barrier1"].append(ev[6].elapsed_time(end))
        return {k: statistics.median(v) for k, v in acc.items()}

    def _capture(self):
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(2):
                self._enqueue()
        self._stream.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._enqueue()
            self._graph = g
        except Exception as e:                       # the barrier op may refuse capture on some builds: run eagerly
            self._graph = None
            self.capture_error = repr(e)
            torch.cuda.synchronize()

    def run_device(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._enqueue()
        self._steps += 1                                 # (== the device epoch: every step bumps it once)
        return self.batch, (self.bits2_g[self._steps & 1] if self._tail is not None else self.bits_g)

    def adjacency_bits(self):
        """Global bit matrix [S, words]: already complete on every rank (tail mode: the buffer of the last epoch)."""
        if self._tail is not None:
            return self.bits2_g[self._steps & 1]
        return self.bits_g
