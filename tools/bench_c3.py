"""C3 (BASELINE configs[2]): independent planning queries over 200-obstacle scenes, planned in lock step
with batched kernel calls.  Under torchrun the queries are sharded over the ranks (no communication).

  python tools/bench_c3.py [--queries 512] [--procs 4]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c3.py --queries 512
(--queries is per GPU; --procs P drives every GPU from P host processes -- the planner loop is host Python and
one process cannot keep a B200 busy: the kernels of different processes interleave on the same device)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import scenes
from boundplanner_b200.planner import plan_batch


def build_queries(ids):
    r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
    queries = []
    for i in ids:
        ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
        queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
    return queries, list(wmin), list(wmax)


def summarise(results, stats):
    errs = {}
    for r in results:
        if isinstance(r, Exception):
            k = type(r).__name__ + ": " + str(r)[:40]
            errs[k] = errs.get(k, 0) + 1
    ok = [r for r in results if not isinstance(r, Exception)]
    return dict(planned=len(ok), errors=errs, rounds=stats["rounds"], kernel_batches=stats["kernel_batches"],
                sets=[r["graph"].number_of_nodes() for r in ok])


def _worker(conn, device, ids):
    """One host process of a GPU: builds its queries, warms up, then plans them on `go`."""
    torch.cuda.set_device(device)
    queries, wmin, wmax = build_queries(ids)
    plan_batch(queries[:8], 0.01, wmax, wmin, rng_seeds=ids[:8])
    torch.cuda.synchronize()
    conn.send("ready")
    conn.recv()
    results, stats = plan_batch(queries, 0.01, wmax, wmin, rng_seeds=ids)
    torch.cuda.synchronize()
    conn.send(summarise(results, stats))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=512)
    ap.add_argument("--procs", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    base = rank * args.queries
    if args.procs > 1:
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        per = (args.queries + args.procs - 1) // args.procs
        workers = []
        for k in range(args.procs):
            ids = list(range(base + k * per, min(base + (k + 1) * per, base + args.queries)))
            if ids:
                a, b = ctx.Pipe()
                pr = ctx.Process(target=_worker, args=(b, lr, ids), daemon=True)
                pr.start()
                workers.append((pr, a))
        for _, a in workers:
            assert a.recv() == "ready"
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _, a in workers:
            a.send("go")
        outs = [a.recv() for _, a in workers]
        dt = time.perf_counter() - t0
        for pr, _ in workers:
            pr.join()
    else:
        ids = list(range(base, base + args.queries))
        queries, wmin, wmax = build_queries(ids)
        plan_batch(queries[:8], 0.01, wmax, wmin, rng_seeds=ids[:8])          # warm-up
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        results, stats = plan_batch(queries, 0.01, wmax, wmin, rng_seeds=ids)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        outs = [summarise(results, stats)]
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = tt.item()
    if rank == 0:
        errs = {}
        for o in outs:
            for k, v in o["errors"].items():
                errs[k] = errs.get(k, 0) + v
        sets = [x for o in outs for x in o["sets"]]
        print(json.dumps({"config": "C3", "n_gpus": world, "host_procs_per_gpu": args.procs,
                          "host_cores": os.cpu_count(), "queries_per_gpu": args.queries, "seconds": dt,
                          "queries_per_sec": args.queries * world / dt,
                          "planned_rank0": sum(o["planned"] for o in outs), "errors_rank0": errs,
                          "rounds": max(o["rounds"] for o in outs),
                          "kernel_batches": sum(o["kernel_batches"] for o in outs),
                          "mean_sets_built": float(np.mean(sets)) if sets else 0.0}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
