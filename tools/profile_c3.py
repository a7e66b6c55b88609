"""cProfile of the lock-step C3 driver (GPU box): where the host time of plan_batch goes.
  python tools/profile_c3.py [queries]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import plan_batch
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 128
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
ids = list(range(nq))
queries = []
for i in ids:
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
    queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
wmin, wmax = list(wmin), list(wmax)
plan_batch(queries[:8], 0.01, wmax, wmin, rng_seeds=ids[:8])
torch.cuda.synchronize()
t0 = time.perf_counter()
results, stats = plan_batch(queries, 0.01, wmax, wmin, rng_seeds=ids)
torch.cuda.synchronize()
print(f"{nq} queries in {time.perf_counter() - t0:.3f} s, rounds {stats['rounds']}, batches {stats['kernel_batches']}")
pr = cProfile.Profile()
pr.enable()
results, stats = plan_batch(queries, 0.01, wmax, wmin, rng_seeds=ids)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
