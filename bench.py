#!/usr/bin/env python
"""bench.py -- headline benchmark of the BoundPlanner geometry core on B200.

Workload (BASELINE.json configs[1], "C2"): synthetic 1k random box obstacles,
256 seeds per GPU, set inflation (find_set_around_point, fixed_mid=True,
optimize=True) + all-pairs set-graph build (tol 0.01).  One "step" = one pass of
that hot path over the seed batch.  Metric: convex sets built per second
(whole job), with set-pair checks/s alongside.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL): every rank builds its
own 256 seeds over the replicated scene (weak scaling), the halfspace tensors
are all-gathered and the pair matrix is split in balanced row blocks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_OBS = 1000
N_SEEDS = 256           # per GPU
TOL = 0.01
METRIC = "convex_sets_built_per_sec"
UNIT = "sets/s"
WORKLOAD = "C2: 1000 random box obstacles (edge U[0.02,0.08], inflate 0.01), 256 seeds/GPU, IRIS loop fixed_mid + all-pairs set graph tol 0.01"


# ----------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's Python path
# ----------------------------------------------------------------------------
def _oracle_worker(args):
    boxes, inflate, ws_min, ws_max, seeds = args[:5]
    per_obstacle_loop = len(args) > 5 and args[5]
    from oracle.convex_set_finder import ConvexSetFinder
    from oracle.obstacles import obstacle_reps

    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    f = ConvexSetFinder(obs_sets, pts, ws_max, ws_min, max_rows=None)
    if per_obstacle_loop:
        # the reference's own loop structure: one closest-point QP solve per obstacle and pass
        # (ConvexSetFinder.py:468-486), instead of the oracle's NumPy batch over all obstacles
        f.compute_set_projs = f.compute_set_projs_loop
    sets = []
    for p in seeds:
        try:
            A, b, _, _ = f.find_set_around_point(p, fixed_mid=True, optimize=True)
            sets.append([A, b])
        except RuntimeError:
            pass
    return sets


def cpu_sample(boxes, inflate, ws_min, ws_max, seeds, cores):
    """Build len(seeds) sets with the oracle on `cores` processes and test all their
    pairs with the reference's linprog call; returns (seconds, n_sets, n_pairs)."""
    from oracle.set_graph import set_intersection

    t0 = time.perf_counter()
    if cores == 1:
        sets = _oracle_worker((boxes, inflate, ws_min, ws_max, seeds))
    else:
        import multiprocessing as mp

        chunks = [seeds[i::cores] for i in range(cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            parts = pool.map(_oracle_worker, [(boxes, inflate, ws_min, ws_max, c) for c in chunks if len(c)])
        sets = [s for p in parts for s in p]
    npairs = 0
    for i in range(len(sets)):
        for j in range(i):
            set_intersection(sets[i], sets[j], TOL)
            npairs += 1
    return time.perf_counter() - t0, len(seeds), npairs


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The true
    reference cannot run here (casadi/cvxpy/cdd absent, no network), so this is
    the oracle port (kind "port"), split over all host cores, on a bounded sample
    of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from boundplanner_b200 import scenes

    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(N_OBS, N_SEEDS)
    cores = os.cpu_count() or 1
    cores = min(cores, 32)
    per_step = max(cores * 2, 8)
    for _ in range(args.warmup):
        cpu_sample(boxes, inflate, ws_min, ws_max, seeds[: max(2, cores)], cores)
    tot_t, tot_sets, tot_pairs = 0.0, 0, 0
    for k in range(args.steps):
        sel = seeds[(k * per_step) % N_SEEDS: (k * per_step) % N_SEEDS + per_step]
        t, ns, npairs = cpu_sample(boxes, inflate, ws_min, ws_max, sel, cores)
        tot_t += t
        tot_sets += ns
        tot_pairs += npairs
    value = tot_sets / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{per_step} seeds + their pair checks per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} of the {N_SEEDS} C2 seeds per step, vectorised NumPy oracle "
                                   f"(oracle/convex_set_finder.py) over {cores} processes + scipy linprog per pair"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pair_checks_per_sec": tot_pairs / tot_t,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from boundplanner_b200 import _lib, distributed as bpd, geometry as geo, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.load()

    # same scene on every rank, per-rank seeds (weak scaling: 256 seeds per GPU)
    boxes, inflate, seeds0, ws_min, ws_max = scenes.config_c2(N_OBS, N_SEEDS)
    if rank == 0:
        seeds = seeds0
    else:
        seeds = scenes.free_points(N_SEEDS, boxes, inflate, np.random.default_rng(100 + rank), ws_min, ws_max)
    scene = geo.Scene(boxes, inflate)
    seeds_host = torch.as_tensor(seeds).pin_memory()
    seeds_dev = seeds_host.cuda()
    S_total = N_SEEDS * world

    def pair_fn(A, b, m, tol, r0, r1):
        return geo.pair_feasible(A, b, m, tol, r0, r1)

    # single GPU: static buffers + one CUDA graph per step (boundplanner_b200/pipeline.py);
    # multi GPU: the same kernels eagerly around the NCCL all-gathers
    pipe = None
    if world == 1 and not args.no_cuda_graph:
        from boundplanner_b200.pipeline import SetGraphPipeline

        pipe = SetGraphPipeline(scene, N_SEEDS, ws_min, ws_max, fixed_mid=True, optimize=True, tol=TOL)
        pipe.seeds_dev.copy_(seeds_dev)

    spipe = None
    if world > 1 and not args.no_cuda_graph:
        from boundplanner_b200.pipeline import PeerSetGraphPipeline, ShardedSetGraphPipeline

        # exchange by peer stores over NVLink (symmetric memory) in one CUDA graph per step; BPGEO_PEER=0 or a
        # box without peer mappings -> the NCCL all-gather pipeline
        peer_note = None
        if os.environ.get("BPGEO_PEER", "1") != "0":
            try:
                spipe = PeerSetGraphPipeline(scene, N_SEEDS, ws_min, ws_max, fixed_mid=True, optimize=True, tol=TOL)
            except Exception as e:  # noqa: BLE001
                peer_note = f"symmetric memory unavailable ({type(e).__name__}): NCCL pipeline"
                spipe = None
        if spipe is None:
            spipe = ShardedSetGraphPipeline(scene, N_SEEDS, ws_min, ws_max, fixed_mid=True, optimize=True, tol=TOL)
        spipe.seeds_dev.copy_(seeds_dev)

    def step_device(seeds_d):
        if pipe is not None:
            return pipe.run_device()
        if spipe is not None:
            if seeds_d is not spipe.seeds_dev and seeds_d is not seeds_dev:
                spipe.seeds_dev.copy_(seeds_d)
            return spipe.run_device()
        out = geo.build_sets_point(scene, seeds_d, ws_min, ws_max, fixed_mid=True, optimize=True)
        bits, _ = bpd.sharded_adjacency(out.A, out.b, out.m, pair_fn, TOL)
        return out, bits

    e2e_host = []

    def step_e2e():
        if pipe is not None:
            return pipe.run(seeds_host)                              # H2D seeds, graph replay, D2H results
        sd = seeds_host.cuda(non_blocking=True)                      # H2D of the step's inputs
        out, bits = step_device(sd)
        srcs = (out.A, out.b, out.m, out.q_ellipse, out.p_mid, out.status, bits)
        if not e2e_host:                                             # pinned mirrors, allocated once
            e2e_host.extend(torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in srcs)
        for h, t in zip(e2e_host, srcs):                             # D2H of every result, one synchronize
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return e2e_host

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out, bits = step_device(seeds_dev)
    barrier()

    # ---- timed region: K steps, device-resident inputs, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))
        ev[k][0].record()
        out, bits = step_device(seeds_dev)
        ev[k][1].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    e2e_depth = 1
    if pipe is not None and os.environ.get("BPGEO_E2E_DEPTH", "2") != "1":
        # two pipeline objects used round-robin: every step still copies its seeds in and all its results out, but
        # the D2H of step k (copy stream) overlaps step k+1 and the host does not synchronize per step
        from boundplanner_b200.pipeline import PipelinedSetGraph

        e2e_depth = 2
        ps = PipelinedSetGraph(scene, N_SEEDS, ws_min, ws_max, depth=e2e_depth, fixed_mid=True, optimize=True, tol=TOL)
        for _ in range(4):
            ps.take()
            ps.put(seeds_host)
        ps.drain()
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            got = ps.take()                      # results of step k - 2, delivered in pinned host memory
            if got is not None:
                res = got
            ps.put(seeds_host)
        res = ps.drain()[-1]
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            res = step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
    h2d = seeds_host.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in res)

    # ---- per-kernel timing of the dominant kernels (roofline), CUDA events on the launch stream
    prof = kernel_profile(geo, scene, seeds_dev, ws_min, ws_max, out, torch)

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = tt[0].item(), tt[1].item()
    status = out.status.cpu().numpy()
    mrows = out.m.cpu().numpy()
    n_pairs = S_total * (S_total - 1) // 2
    value = S_total * args.steps / (dev_ms * 1e-3)
    fused = os.environ.get("BPGEO_FUSED", "1") != "0" and N_OBS <= 4096
    # fused: k_iris_fused + aabb + filter + lp;  else state_init, 5x(poly,mvie), final mvie, export, aabb+filter+lp
    launches_per_step = (1 + 3) if fused else (1 + 5 * 2 + 1 + 1 + 3)
    if type(spipe).__name__ == "PeerSetGraphPipeline":
        launches_per_step += 2                      # the two peer-scatter kernels (barriers are torch's)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_obstacles": N_OBS, "seeds_per_gpu": N_SEEDS, "pairs": n_pairs,
                       "l2": "flushed between timed steps (256 MiB fill)",
                       "launch": ("one CUDA graph per step" if pipe is not None else
                                  ("one CUDA graph per step; sets and adjacency rows exchanged by peer stores over "
                                   "NVLink + two signal-pad barriers (no NCCL on the data path)"
                                   + ("" if getattr(spipe, "_graph", None) is not None else " [eager: capture refused]"))
                                  if type(spipe).__name__ == "PeerSetGraphPipeline" else
                                  "two CUDA graphs + two NCCL all-gathers per step" if spipe is not None else
                                  "eager launches"),
                       "frac_sets_over_20_rows": float((mrows > 20).mean()),
                       "frac_status_ok": float((status == 0).mean()),
                       "adjacency_density": float(geo.unpack_adjacency(
                           spipe.adjacency_bits() if spipe is not None else bits, S_total).sum().item())
                       / max(n_pairs, 1)},
            "pair_checks_per_sec": n_pairs * args.steps / (dev_ms * 1e-3),
            "e2e": {"value": S_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps_in_flight": e2e_depth},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": prof["roofline"](hbm_peak, "measured" if "hbm_gbs" in peaks else "fallback"),
            "roofline_fk": prof["roofline_fk"](hbm_peak, "measured" if "hbm_gbs" in peaks else "fallback"),
            "kernels": prof["kernels"],
        }
        if world == 1 and not args.no_plan_latency:
            line["plan_latency"] = plan_latency(not args.no_cpu_baseline)
        # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 24
            t, ns, npairs = cpu_sample(boxes, inflate, ws_min, ws_max, seeds0[:n_cpu], 1)
            line["cpu_baseline"] = {
                "value": ns / t, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {n_cpu} of the {N_SEEDS} C2 seeds + their {npairs} pair checks, vectorised NumPy "
                          "oracle (stronger than the reference's per-obstacle solver loop), 1 core"}
            # the same port with the reference's loop structure (one QP solve per obstacle and pass), 2 seeds
            t0 = time.perf_counter()
            n_loop = len(_oracle_worker((boxes, inflate, ws_min, ws_max, seeds0[:2], True)))
            line["cpu_baseline"]["per_obstacle_loop_sets_per_sec"] = n_loop / (time.perf_counter() - t0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def plan_latency(with_cpu):
    """p50 plan latency (BASELINE.json metric, third part): SetSequencePlanner.plan_set_sequence -- the
    reference's plan_convex_set_path up to the planned set sequence -- on the C1 example and on the
    C3-style queries that plan successfully, every primitive a batch-of-one kernel call."""
    import statistics

    from scipy.spatial.transform import Rotation as R

    from boundplanner_b200 import scenes
    from boundplanner_b200.planner import SetSequencePlanner

    r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()

    def cases():
        boxes, ws_min, ws_max, inflate = scenes.example_scene()
        yield "C1", boxes, inflate, np.array([0.3, 0.0, 0.7]), np.array([0.45, -0.5, 0.2]), ws_min, ws_max, 0
        for i in (0, 7, 8, 11):
            ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
            yield f"C3[{i}]", ob, infl, st, en, wmin, wmax, i

    def run(backend_factory, names=None):
        lat = {}
        for name, ob, infl, st, en, wmin, wmax, seed in cases():
            if names is not None and name not in names:
                continue
            backend = backend_factory(ob, infl, list(wmax), list(wmin))
            for rep in range(3 if names is None else 1):
                planner = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=backend,
                                             rng=np.random.default_rng(seed))
                t0 = time.perf_counter()
                res = planner.plan_set_sequence(st.copy(), en.copy(), r0, r0)
                lat[name] = (time.perf_counter() - t0) * 1e3, len(res["set_ids"]), res["graph"].number_of_nodes()
        return lat

    from boundplanner_b200.planner import GpuBackend

    gpu = run(GpuBackend)
    vals = sorted(v[0] for v in gpu.values())
    out = {"unit": "ms", "p50": statistics.median(vals), "max": vals[-1],
           "queries": {k: {"ms": v[0], "sets_in_sequence": v[1], "sets_built": v[2]} for k, v in gpu.items()},
           "what": "plan_convex_set_path up to the planned set sequence (no final Ipopt NLP), host loop + "
                   "batch-of-one kernel calls"}
    if with_cpu:
        from tests.util import OracleBackend

        cpu = run(OracleBackend, names=("C1", "C3[7]"))
        out["cpu_port_ms"] = {k: v[0] for k, v in cpu.items()}
    return out


def kernel_profile(geo, scene, seeds_dev, ws_min, ws_max, out, torch):
    """Time the kernels of one step in isolation (CUDA events on the launch stream,
    several back-to-back launches per measurement) and build the roofline entries."""
    import ctypes
    import statistics

    from boundplanner_b200 import _lib
    from boundplanner_b200.robot_model import Q_LIM_UPPER

    def timeit(fn, reps=5, inner=4):
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / inner)
        return statistics.median(ts)

    S = seeds_dev.shape[0]
    N = scene.n
    # one polyhedron pass (K3) with the initial sphere, one fixed-mid and one free MVIE (K4) on the built rows
    q0 = torch.eye(3, dtype=torch.float64, device="cuda").repeat(S, 1, 1)
    init_rows = torch.zeros((S, 6, 4), dtype=torch.float64, device="cuda")
    for i in range(3):
        init_rows[:, 2 * i, i] = 1.0
        init_rows[:, 2 * i, 3] = float(ws_max[i])
        init_rows[:, 2 * i + 1, i] = -1.0
        init_rows[:, 2 * i + 1, 3] = -float(ws_min[i])
    qi, qe = q0 * 1e-4, q0 * 1e4
    t_poly = timeit(lambda: geo.polyhedron(scene, seeds_dev, qi, qe, init_rows))
    t_mvie_fm = timeit(lambda: geo.mvie(out.A, out.b, out.m, seeds_dev, False))
    t_mvie_free = timeit(lambda: geo.mvie(out.A, out.b, out.m, seeds_dev, True))
    newton_free = float(geo.mvie(out.A, out.b, out.m, seeds_dev, True)[4].double().mean().item())
    newton_fm = float(geo.mvie(out.A, out.b, out.m, seeds_dev, False)[4].double().mean().item())
    t_pair = timeit(lambda: geo.pair_feasible(out.A, out.b, out.m, TOL))
    sb = geo.alloc_set_batch(S)
    t_build = timeit(lambda: geo.build_sets_point(scene, seeds_dev, ws_min, ws_max, fixed_mid=True, optimize=True,
                                                  out=sb))
    passes_mean = float(torch.clamp(sb.iters.double(), max=5).mean().item())
    m_mean = float(out.m.double().mean().item())

    # FP64 pipe peak of this part (dependent-free DFMA chains, full chip)
    lib = _lib.load()
    probe_out = torch.empty(296 * 256, dtype=torch.float64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    iters = 20000
    t_probe = timeit(lambda: lib.bp_probe_fp64(8, 296, 256, iters, ctypes.c_void_p(probe_out.data_ptr()), st),
                     reps=3, inner=1)
    fp64_peak_tflops = 296 * 256 * 8 * iters * 2 / (t_probe * 1e-3) / 1e12

    # K7 FK micro-benchmark (C5): B = 2^20 configurations, inputs larger than L2 are flushed between runs
    B = 1 << 20
    lim = torch.as_tensor(Q_LIM_UPPER, device="cuda")
    qq = (torch.rand((B, 7), dtype=torch.float64, device="cuda") * 2 - 1) * lim
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    ts = []
    for k in range(8):
        flush.fill_(float(k))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        geo.fk_iiwa14(qq)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t_fk = statistics.median(ts[2:])
    fk_bytes = B * (56 + 24 + 168)

    kernels = {
        "build_sets_point_ms": t_build, "iris_passes_mean": passes_mean,
        "k_poly_point_ms": t_poly, "k_mvie_fixed_mid_ms": t_mvie_fm, "k_mvie_free_ms": t_mvie_free,
        "pair_pipeline_ms": t_pair, "k_fk_1M_ms": t_fk, "newton_iters_fixed_mid": newton_fm,
        "newton_iters_free": newton_free, "mean_rows": m_mean, "fp64_peak_tflops_measured": fp64_peak_tflops,
    }

    def roofline(hbm_peak, which):
        # dominant kernel of the step: the set build (k_iris_fused: one launch = the whole IRIS loop of S seeds;
        # with BPGEO_FUSED=0 the same work is the (k_poly_point, k_mvie) launch sequence).
        # Algorithmic bytes per launch: seeds in, the scene once, one padded set + ellipsoid
        # out per seed.  Algorithmic flops: passes x [N x (1.1 kflop QP + 64 flop x picked rows)] per seed for
        # the polyhedron passes + Newton iterations x (m x 200 + 250) flop for the MVIEs (passes + 1 per seed).
        # (SURVEY 8d's per-unit figures; the kernel itself solves far fewer QPs than N per pass -- lazy lower
        # bounds -- so this is the reference algorithm's work per second, not executed instructions.)
        picks = max(m_mean - 6.0, 0.0)
        bytes_alg = S * 24 + N * 48 + S * (m_mean * 32 + 12 * 8 + 12)
        flops = S * (passes_mean * N * (1100.0 + 64.0 * picks)
                     + (passes_mean * newton_fm + newton_free) * (m_mean * 200.0 + 250.0))
        dur = t_build
        ach = bytes_alg / (dur * 1e-3) / 1e9
        return {"kernel": "k_iris_fused (set build: whole find_set_around_point loop, one CTA per seed)",
                "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": 191744, "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_ncu_final2_iris_fused.txt",
                "peak_source": which,
                "note": "latency-bound fp64 kernel: 256 independent chains of ~5 x (polyhedron pass + ~35 dependent "
                        "Newton iterations); it moves only its compulsory bytes (operands live in L2 / shared "
                        "memory), so the HBM fraction is tiny by construction -- see fp64 below, and roofline_fk "
                        "for the HBM-bound kernel of the path",
                "fp64": {"achieved_gflops": flops / (dur * 1e-3) / 1e9, "peak_gflops": fp64_peak_tflops * 1e3,
                         "frac": flops / (dur * 1e-3) / 1e12 / fp64_peak_tflops}}

    def roofline_fk(hbm_peak, which):
        ach = fk_bytes / (t_fk * 1e-3) / 1e9
        return {"kernel": "k_fk<false,false> (B = 2^20 configurations, 248 B each)", "bound": "hbm", "achieved": ach,
                "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": 201795328,
                "traffic_source": "ncu dram__bytes_read+write per launch (below the 260 MB algorithmic bytes: part "
                                  "of the output is still dirty in the 126 MB L2 at kernel end), "
                                  "profiles/r01_ncu_final2_pair_fk.txt",
                "peak_source": which, "poses_per_sec": B / (t_fk * 1e-3)}

    return {"kernels": kernels, "roofline": roofline, "roofline_fk": roofline_fk}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plan-latency", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
