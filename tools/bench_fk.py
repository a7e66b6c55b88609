"""FK micro-benchmark (C5): B = 2^20 random configurations within joint limits."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from boundplanner_b200 import geometry as geo
from boundplanner_b200.robot_model import Q_LIM_UPPER
B = 1 << 20
g = torch.Generator(device="cuda").manual_seed(0)
lim = torch.as_tensor(Q_LIM_UPPER, device="cuda")
q = (torch.rand((B, 7), dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * lim
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def run(pose, jac, reps=10):
    ts = []
    for _ in range(3): geo.fk_iiwa14(q, pose, jac)
    for k in range(reps):
        flush.fill_(float(k))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); geo.fk_iiwa14(q, pose, jac); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for pose, jac, by in ((False, False, 56 + 24 + 168), (True, False, 56 + 24 + 168 + 128), (True, True, 56 + 24 + 168 + 128 + 336)):
    ms = run(pose, jac)
    print(json.dumps({"pose": pose, "jac": jac, "ms": ms, "Mq_per_s": B / ms / 1e3, "GBps": B * by / ms / 1e6, "bytes_per_q": by}))
