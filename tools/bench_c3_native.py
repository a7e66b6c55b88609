"""C3 with the native lock-step driver (bp_plan_run) against the Python lock-step driver (planner.plan_batch):
queries/s on one GPU.  python tools/bench_c3_native.py [queries] [--no-python]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes, planner_native as pn
from boundplanner_b200.planner import plan_batch

nq = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 128
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
ids = list(range(nq))
queries = []
for i in ids:
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
    queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
wmin, wmax = list(wmin), list(wmax)
pl = pn.NativePlanner(queries, infl, wmax, wmin)
pl.run(ids)
torch.cuda.synchronize()
if "--once" in sys.argv:          # (ncu launch lists: one warmed-up run is enough)
    sys.exit(0)
for rep in range(3):
    t0 = time.perf_counter()
    res, st = pl.run(ids, want_nodes=False)
    dt = time.perf_counter() - t0
    ok = sum(not isinstance(r, Exception) for r in res)
    print(f"native: {nq} queries in {dt * 1e3:.1f} ms = {nq / dt:.0f} queries/s, rounds {st['rounds']}, planned {ok}, "
          f"set requests {st['set_requests']}, pair tests {st['pair_tests']}, projections {st['projections']}, "
          f"paths {st['shortest_paths']}, waiting for the device {st['device_wait_ms']:.1f} ms")
t0 = time.perf_counter()
pl2 = pn.NativePlanner(queries, infl, wmax, wmin)
res, st = pl2.run(ids)
print(f"native incl. scene upload + table allocation: {nq / (time.perf_counter() - t0):.0f} queries/s")
if "--no-python" not in sys.argv:
    plan_batch(queries[:8], infl, wmax, wmin, rng_seeds=ids[:8])
    t0 = time.perf_counter()
    want, stats = plan_batch(queries, infl, wmax, wmin, rng_seeds=ids)
    dt = time.perf_counter() - t0
    print(f"python: {nq} queries in {dt * 1e3:.1f} ms = {nq / dt:.0f} queries/s, rounds {stats['rounds']}")
    same = 0
    for i, (w, g) in enumerate(zip(want, res)):
        if isinstance(w, Exception):
            ok = isinstance(g, Exception) and type(g) is type(w)
        else:
            ok = (not isinstance(g, Exception)) and g["path"] == w["path"] and g["set_ids"] == w["set_ids"] and \
                np.abs(g["p_via"] - w["p_via"]).max() < 1e-9
        same += ok
        if not ok:
            print("MISMATCH query", i, "python:", repr(w) if isinstance(w, Exception) else (w["path"], w["set_ids"]),
                  "native:", repr(g) if isinstance(g, Exception) else (g["path"], g["set_ids"]))
    print(f"identical results: {same} of {nq}")
