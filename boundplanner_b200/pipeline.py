"""Set-graph pipeline with static buffers and CUDA-graph replay.

One "step" of the hot path (BASELINE.json configs[1]): S seeds -> S convex sets
(find_set_around_point) -> all-pairs intersection graph (set_intersection at
tol 0.01).  The 16 kernel launches of a step are captured once into a CUDA
graph; a step is then one H2D copy of the seeds, one graph launch and the D2H
copies of the results, with no per-step allocation and no Python between
kernels.  Single GPU; the multi-GPU path (boundplanner_b200/distributed.py)
runs the same kernels eagerly around its NCCL collectives.
"""
from __future__ import annotations

import torch

from . import geometry as geo


class SetGraphPipeline:
    def __init__(self, scene, n_seeds, ws_min, ws_max, fixed_mid=True, optimize=True, max_iter=5, tol=0.01,
                 m_max=geo.BP_MAX_ROWS, use_cuda_graph=True):
        self.scene, self.S = scene, int(n_seeds)
        self.ws_min, self.ws_max = ws_min, ws_max
        self.kw = dict(fixed_mid=fixed_mid, optimize=optimize, max_iter=max_iter)
        self.tol = tol
        self.seeds_dev = torch.zeros((self.S, 3), dtype=torch.float64, device="cuda")
        self.batch = geo.alloc_set_batch(self.S, m_max)
        self.pair_buf = geo.alloc_pair_buffers(self.S)
        self.bits = self.pair_buf[0]
        self._host = None
        self._graph = None
        self._stream = torch.cuda.Stream()
        if use_cuda_graph:
            self._capture()

    def _enqueue(self):
        geo.build_sets_point(self.scene, self.seeds_dev, self.ws_min, self.ws_max, out=self.batch, **self.kw)
        geo.pair_feasible(self.batch.A, self.batch.b, self.batch.m, self.tol, out=self.pair_buf)

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
            srcs = (self.batch.A, self.batch.b, self.batch.m, self.batch.q_ellipse, self.batch.p_mid,
                    self.batch.status, self.bits)
            self._host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in srcs]
            self._srcs = srcs
        self.seeds_dev.copy_(seeds_host_pinned, non_blocking=True)
        self.run_device()
        for h, d in zip(self._host, self._srcs):
            h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host
