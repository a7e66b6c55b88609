"""Where the fused IRIS kernel spends its cycles, per seed: builds libbpgeo_prof.so (-DBPGEO_PROFILE: clock64
counters around the polyhedron pass and the MVIE solves) and runs the C2 set build once.
Usage (GPU box): BPGEO_LIB is set by this script; run `python tools/prof_phases.py`."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXTRA = [a for a in sys.argv[1:] if a.startswith("-D")]
PROF = os.path.join(ROOT, "boundplanner_b200", os.environ.get("BPGEO_PROF_NAME", "libbpgeo_prof.so"))
if "--build" in sys.argv or not os.path.exists(PROF):
    import __graft_entry__ as g
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DBPGEO_PROFILE"] + EXTRA + ["-o", PROF,
                                                                        os.path.join(g.CSRC, "bpgeo.cu")], cwd=ROOT)
    if "--build" in sys.argv:
        sys.exit(0)
os.environ["BPGEO_LIB"] = PROF
import numpy as np, torch
from boundplanner_b200 import _lib, geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
lib = _lib.load()
S = seeds.shape[0]
buf = np.zeros((S, 4), dtype=np.int64)
mbuf = np.zeros((S, 8), dtype=np.int64)
pbuf = np.zeros((S, 8), dtype=np.int64)
out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
lib.bp_prof_read(buf.ctypes.data_as(ctypes.c_void_p), S, 1)
lib.bp_prof_read_mvie(mbuf.ctypes.data_as(ctypes.c_void_p), S, 1)
lib.bp_prof_read_poly(pbuf.ctypes.data_as(ctypes.c_void_p), S, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
ev1.record()
torch.cuda.synchronize()
lib.bp_prof_read(buf.ctypes.data_as(ctypes.c_void_p), S, 1)
it = out.iters.cpu().numpy()
tot = buf[:, :3].sum(1)
print(f"slowest CTA {buf[:, 3].max()} cycles ({buf[:, 3].max() / 1.965e3:.0f} us at 1965 MHz), mean {buf[:, 3].mean():.0f}")
print(f"kernel {ev0.elapsed_time(ev1) * 1e3:.0f} us; per seed cycles (poly, mvie-in-loop, mvie-final): "
      f"mean {buf[:, 0].mean():.0f} {buf[:, 1].mean():.0f} {buf[:, 2].mean():.0f}")
k = np.argsort(-buf[:, 3])[:8]
for s in k:
    print(f"seed {s}: passes {it[s]} poly {buf[s, 0]} mvie {buf[s, 1]} final {buf[s, 2]} sum {tot[s]} "
          f"CTA total {buf[s, 3]} ({buf[s, 3] / 1.965e3:.0f} us at 1965 MHz) rows {int(out.m[s])}")
print("per pass: poly", (buf[:, 0] / np.minimum(it, 5)).mean(), "mvie6", (buf[:, 1] / np.minimum(it, 5)).mean())
lib.bp_prof_read_poly(pbuf.ctypes.data_as(ctypes.c_void_p), S, 0)
print("polyhedron pass of those seeds (cycles over all passes: bounds | collect+QP | picks | sweep; picks, rounds, QPs):")
for s in k:
    r = pbuf[s]
    print(f"  seed {s}: {r[0]} | {r[2]} | {r[3]} | {r[4]}; picks {r[5]} rounds {r[6]} qps {r[7]}")

lib.bp_prof_read_mvie(mbuf.ctypes.data_as(ctypes.c_void_p), S, 1)
names = ["rows", "dots+H", "ldl", "linesearch", "predictor", "newton", "armijo", "backtracks"]
n = mbuf[:, 5].sum()
print("MVIE inside the fused kernel, cycles per Newton iteration (all seeds):")
for k in range(4):
    print(f"  {names[k]:10s} {mbuf[:, k].sum() / n:8.0f}")
print(f"  newton iterations per seed {mbuf[:, 5].mean():.1f}, armijo evals per iteration {mbuf[:, 6].sum() / n:.2f}, "
      f"line-search trials per iteration {mbuf[:, 7].sum() / n:.2f}")

lib.bp_prof_read_poly(pbuf.ctypes.data_as(ctypes.c_void_p), S, 1)
npass = np.minimum(it, 5).sum()
pn = ["A bounds", "(unused)", "B collect + QPs", "C picks (warp 0)", "D sweep"]
print("polyhedron pass, cycles per pass (thread 0's view):")
for k in range(5):
    print(f"  {pn[k]:18s} {pbuf[:, k].sum() / npass:8.0f}")
print(f"  picks per pass {pbuf[:, 5].sum() / npass:.1f}, rounds per pass {pbuf[:, 6].sum() / npass:.1f}, "
      f"exact QPs per pass {pbuf[:, 7].sum() / npass:.1f}")

# pair LP statistics on the same sets
qbuf = np.zeros(128, dtype=np.int64)
geo.pair_feasible(out.A, out.b, out.m, 0.01)
lib.bp_prof_read_pair(qbuf.ctypes.data_as(ctypes.c_void_p), 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
geo.pair_feasible(out.A, out.b, out.m, 0.01)
e1.record()
torch.cuda.synchronize()
lib.bp_prof_read_pair(qbuf.ctypes.data_as(ctypes.c_void_p), 1)
print(f"pair pipeline {e0.elapsed_time(e1) * 1e3:.0f} us: {qbuf[64]} LPs of {S * (S - 1) // 2} pairs, {qbuf[67]} intersect, "
      f"mean Newton iterations {qbuf[65] / max(qbuf[64], 1):.1f}, slowest LP {qbuf[66]} cycles, "
      f"{qbuf[68]} more pairs rejected by the margin test before the LP")
print("  iterations histogram (bucket of 2):", [int(v) for v in qbuf[:40]])
