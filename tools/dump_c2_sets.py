"""Dump the C2 set batch (GPU) to gpurun_out/c2_sets.npz for offline analysis."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
bits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
adj = geo.unpack_adjacency(bits, 256)
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/c2_sets.npz", A=out.A.cpu().numpy(), b=out.b.cpu().numpy(), m=out.m.cpu().numpy(),
         q=out.q_ellipse.cpu().numpy(), p=out.p_mid.cpu().numpy(), status=out.status.cpu().numpy(),
         iters=out.iters.cpu().numpy(), adj=adj.cpu().numpy())
print("ok", out.iters.cpu().numpy()[:32], adj.sum().item())
