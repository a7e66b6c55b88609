"""cProfile of one sequential plan (batch-of-one kernel calls) -- the p50 plan latency path of bench.py."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import GpuBackend, SetSequencePlanner
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
qid = int(sys.argv[1]) if len(sys.argv) > 1 else 7
ob, infl, st, en, wmin, wmax = scenes.config_c3_query(qid)
backend = GpuBackend(ob, infl, list(wmax), list(wmin))
for rep in range(3):
    planner = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=backend, rng=np.random.default_rng(qid))
    t0 = time.perf_counter()
    res = planner.plan_set_sequence(st.copy(), en.copy(), r0, r0)
    print("ms", (time.perf_counter() - t0) * 1e3, "sets built", res["graph"].number_of_nodes(), "inter nodes",
          res["inter_graph"].number_of_nodes())
planner = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=backend, rng=np.random.default_rng(qid))
pr = cProfile.Profile()
pr.enable()
planner.plan_set_sequence(st.copy(), en.copy(), r0, r0)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

# per-request-kind wall time of the batch-of-one backend
import collections
acc = collections.defaultdict(lambda: [0, 0.0])
orig = backend.execute
def timed(req):
    t0 = time.perf_counter()
    try:
        return orig(req)
    finally:
        a = acc[req[0]]
        a[0] += 1
        a[1] += time.perf_counter() - t0
backend.execute = timed
planner = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=backend, rng=np.random.default_rng(qid))
t0 = time.perf_counter()
planner.plan_set_sequence(st.copy(), en.copy(), r0, r0)
tot = time.perf_counter() - t0
print("total ms", tot * 1e3, {k: (v[0], round(v[1] * 1e3, 2)) for k, v in acc.items()})
