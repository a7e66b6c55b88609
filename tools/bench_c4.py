"""C4 (10k obstacles, 2048 seeds, 2 096 128 pair checks) timing on one GPU -- not a bench.py line."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
def step():
    out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
    return out, bits
for _ in range(3): out, bits = step()
ts, tp = [], []
for _ in range(5):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize(); a.record()
    out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True); b.record()
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01); c.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b)); tp.append(b.elapsed_time(c))
S = seeds.shape[0]; npairs = S * (S - 1) // 2
m = out.m.cpu().numpy(); st = out.status.cpu().numpy()
adj = geo.unpack_adjacency(bits, S)
print(json.dumps({"config": "C4", "n_obstacles": int(boxes.shape[0]), "seeds": S, "set_build_ms": float(np.median(ts)),
                  "pair_ms": float(np.median(tp)), "sets_per_s": S / (np.median(ts) + np.median(tp)) * 1e3,
                  "pair_checks_per_s": npairs / np.median(tp) * 1e3, "mean_rows": float(m.mean()),
                  "frac_over_20_rows": float((m > 20).mean()), "frac_ok": float((st == 0).mean()),
                  "adjacency_density": float(adj.sum().item()) / npairs}))
