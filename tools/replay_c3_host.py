"""Host-loop cost of the lock-step planner WITHOUT a GPU: record the request/answer traces of a few C3 queries with
the oracle backend (slow, once; cached under /tmp), then replay them through plan_batch with an executor that only
hands back the recorded answers.  What is timed is the planner's own Python (generators, sampling, graph
bookkeeping, shortest paths) -- the part that bounds C3 throughput.

  python tools/replay_c3_host.py [--queries 32] [--replicas 8] [--profile]"""
import argparse, cProfile, os, pickle, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import SetSequencePlanner, plan_batch

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=32)
ap.add_argument("--replicas", type=int, default=8)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
cache = f"/tmp/bp_c3_traces_{args.queries}.pkl"
if os.path.exists(cache):
    traces = pickle.load(open(cache, "rb"))
else:
    from tests.util import OracleBackend

    traces = {}
    for qid in range(args.queries):
        ob, infl, st, en, wmin, wmax = scenes.config_c3_query(qid)
        be = OracleBackend(ob, infl, list(wmax), list(wmin))
        pl = SetSequencePlanner(ob, infl, list(wmax), list(wmin), backend=be, rng=np.random.default_rng(qid))
        gen, tr = pl.plan_gen(st.copy(), en.copy(), r0, r0), []
        try:
            req = next(gen)
            while True:
                try:
                    ans = be.execute(req)
                except (RuntimeError, ValueError) as e:
                    ans = e
                tr.append(ans)
                req = gen.throw(ans) if isinstance(ans, Exception) else gen.send(ans)
        except (StopIteration, RuntimeError, ValueError):
            pass
        traces[qid] = tr
        print(f"recorded query {qid}: {len(tr)} requests", flush=True)
    pickle.dump(traces, open(cache, "wb"))
queries, seeds, tr = [], [], []
for rep in range(args.replicas):
    for qid in sorted(traces):
        ob, infl, st, en, wmin, wmax = scenes.config_c3_query(qid)
        queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
        seeds.append(qid)
        tr.append(traces[qid])


class Replay:
    def __init__(self):
        self.pos, self.calls = [0] * len(queries), 0

    def execute(self, pending):
        out = {}
        for q in pending:
            out[q] = tr[q][self.pos[q]]
            self.pos[q] += 1
        return out


def run():
    return plan_batch(queries, 0.01, list(wmax), list(wmin), rng_seeds=seeds, executor=Replay())


run()
t0 = time.process_time()
res, stats = run()
dt = time.process_time() - t0
print(f"{len(queries)} queries, host loop only: {dt:.3f} s CPU = {dt / len(queries) * 1e3:.2f} ms per query, "
      f"{stats['rounds']} rounds, {sum(not isinstance(r, Exception) for r in res)} planned")
if args.profile:
    pr = cProfile.Profile()
    pr.enable()
    run()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
