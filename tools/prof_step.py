"""Driver for ncu captures of the set-build kernels: one C2 build_sets_point call after warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
for _ in range(2):
    out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
torch.cuda.synchronize()
print("done")
