"""Pair stage of ONE rank of an N-GPU C2 step, emulated on one GPU: N x 256 C2 seeds are built here, then the
balanced row block of rank r is tested (k_pair_filter + k_pair_lp) with the -DBPGEO_PROFILE counters.
  python tools/prof_pairs_block.py [world=8]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF = os.path.join(ROOT, "boundplanner_b200", "libbpgeo_prof.so")
if "--build" in sys.argv or not os.path.exists(PROF):
    import __graft_entry__ as g
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DBPGEO_PROFILE", "-o", PROF, os.path.join(g.CSRC, "bpgeo.cu")], cwd=ROOT)
    if "--build" in sys.argv:
        sys.exit(0)
os.environ["BPGEO_LIB"] = PROF
import numpy as np, torch
from boundplanner_b200 import _lib, geometry as geo, scenes
from boundplanner_b200.distributed import balanced_row_blocks
world = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 8
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(n_seeds=256 * world)
sc = geo.Scene(boxes, inflate)
S = seeds.shape[0]
aabb = torch.empty((S, 6), dtype=torch.float64, device="cuda")
out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True, aabb=aabb)
lib = _lib.load()
blocks = balanced_row_blocks(S, world)
for r in (0, world // 2, world - 1):
    r0, r1 = blocks[r]
    bufs = geo.alloc_pair_buffers(S, r1 - r0)
    geo.pair_feasible(out.A, out.b, out.m, 0.01, r0, r1, out=bufs, aabb=aabb)
    qbuf = np.zeros(128, dtype=np.int64)
    lib.bp_prof_read_pair(qbuf.ctypes.data_as(ctypes.c_void_p), 1)
    geo.pair_feasible(out.A, out.b, out.m, 0.01, r0, r1, out=bufs, aabb=aabb)
    torch.cuda.synchronize()
    lib.bp_prof_read_pair(qbuf.ctypes.data_as(ctypes.c_void_p), 1)
    st = geo.pair_feasible_stages(out.A, out.b, out.m, 0.01, r0, r1, out=bufs, aabb=aabb)
    pairs = sum(S - 1 - i for i in range(r0, r1))
    print(f"rank {r}: rows [{r0},{r1}) pairs {pairs}: {qbuf[64]} LPs, {qbuf[67]} intersect, {qbuf[68]} margin rejects, mean Newton "
          f"{qbuf[65] / max(qbuf[64], 1):.1f}, slowest LP {qbuf[66]} cycles ({qbuf[66] / 1.965e3:.0f} us); stages(ms, profile build) {st}")
    print("   iterations histogram (bucket of 2):", [int(v) for v in qbuf[:24]])
