"""Workload for tools/ncu_capture.sh: the C2 step kernels, the FK kernel and the saturated set build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
aabb = torch.empty((256, 6), dtype=torch.float64, device="cuda")
for _ in range(3):
    out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True, aabb=aabb)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01, aabb=aabb)
q = torch.rand((1 << 20, 7), dtype=torch.float64, device="cuda")
for _ in range(3):
    geo.fk_iiwa14(q)
seeds8 = scenes.free_points(2048, boxes, inflate, np.random.default_rng(7), ws_min, ws_max)
sd8 = torch.as_tensor(seeds8).cuda()
for _ in range(2):
    geo.build_sets_point(sc, sd8, ws_min, ws_max, fixed_mid=True, optimize=True)
torch.cuda.synchronize()
