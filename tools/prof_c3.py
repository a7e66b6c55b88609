"""cProfile of the lock-step batched planner on C3 queries (where does the host time go)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import plan_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
queries = []
for i in range(n):
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
    queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
plan_batch(queries[:8], 0.01, list(wmax), list(wmin), rng_seeds=list(range(8)))
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
results, stats = plan_batch(queries, 0.01, list(wmax), list(wmin), rng_seeds=list(range(n)))
pr.disable()
print("seconds", time.perf_counter() - t0, stats)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
