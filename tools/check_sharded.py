"""torchrun check (N >= 2 GPUs): the sharded CUDA-graph pipeline gives the same global adjacency as one
rank testing every pair of the all-gathered sets; prints OK per rank.  --peer: the peer-store pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from boundplanner_b200 import geometry as geo, scenes
from boundplanner_b200.pipeline import PeerSetGraphPipeline, ShardedSetGraphPipeline
PEER = "--peer" in sys.argv
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
boxes, inflate, seeds0, ws_min, ws_max = scenes.config_c2(1000, 128)
seeds = seeds0 if rank == 0 else scenes.free_points(128, boxes, inflate, np.random.default_rng(100 + rank), ws_min, ws_max)
sc = geo.Scene(boxes, inflate)
pipe = (PeerSetGraphPipeline if PEER else ShardedSetGraphPipeline)(sc, 128, ws_min, ws_max, tol=0.01)
if PEER:
    print(f"rank {rank}: peer pipeline, CUDA graph captured: {pipe._graph is not None}", getattr(pipe, "capture_error", ""))
pipe.seeds_dev.copy_(torch.as_tensor(seeds).cuda())
for _ in range(3):
    pipe.run_device()
torch.cuda.synchronize()
bits = pipe.adjacency_bits().cpu().numpy()
full = geo.pair_feasible(pipe.Ag, pipe.bg, pipe.mg, 0.01).cpu().numpy()
own = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
lo = rank * 128
ok = np.array_equal(bits, full) and np.array_equal(pipe.Ag[lo:lo + 128].cpu().numpy(), own.A.cpu().numpy())
print(f"rank {rank}: sharded adjacency == single-rank adjacency: {ok}; edges {int(np.unpackbits(bits.view(np.uint8)).sum())}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
