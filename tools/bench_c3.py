"""C3 (BASELINE configs[2]): independent planning queries over 200-obstacle scenes, planned in lock step
with batched kernel calls.  Under torchrun the queries are sharded over the ranks (no communication).

  python tools/bench_c3.py [--queries 512]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c3.py --queries 512
(--queries is per GPU)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import plan_batch

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=512)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
ids = list(range(rank * args.queries, (rank + 1) * args.queries))
queries = []
for i in ids:
    ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
    queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
plan_batch(queries[:8], 0.01, list(wmax), list(wmin), rng_seeds=ids[:8])          # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
results, stats = plan_batch(queries, 0.01, list(wmax), list(wmin), rng_seeds=ids)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = tt.item()
ok = [r for r in results if not isinstance(r, Exception)]
errs = {}
for r in results:
    if isinstance(r, Exception):
        k = type(r).__name__ + ": " + str(r)[:40]
        errs[k] = errs.get(k, 0) + 1
if rank == 0:
    print(json.dumps({"config": "C3", "n_gpus": world, "queries_per_gpu": args.queries, "seconds": dt,
                      "queries_per_sec": args.queries * world / dt, "planned_rank0": len(ok), "errors_rank0": errs,
                      "rounds": stats["rounds"], "kernel_batches": stats["kernel_batches"],
                      "mean_sets_built": float(np.mean([r["graph"].number_of_nodes() for r in ok])) if ok else 0.0}))
if world > 1:
    dist.destroy_process_group()
