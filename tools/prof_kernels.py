"""Small driver for ncu captures: builds the C2 sets once, then launches each hot kernel a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
torch.cuda.synchronize()
for _ in range(2):
    geo.mvie(out.A, out.b, out.m, sd, False)
    geo.mvie(out.A, out.b, out.m, sd, True)
    geo.pair_feasible(out.A, out.b, out.m, 0.01)
torch.cuda.synchronize()
q = torch.rand((1 << 20, 7), dtype=torch.float64, device="cuda") * 4 - 2
for _ in range(2):
    geo.fk_iiwa14(q)
torch.cuda.synchronize()
print("done")
