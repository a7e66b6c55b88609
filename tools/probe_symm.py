"""Probe (N >= 2 GPUs): torch symmetric memory on this box -- peer buffer pointers, a peer store, barrier."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm_mem.empty((world, 1024), dtype=torch.float64, device=f"cuda:{lr}")
t.zero_()
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok: ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast ptr", hex(hdl.multicast_ptr), flush=True)
hdl.barrier(channel=0)
for peer in range(world):
    buf = hdl.get_buffer(peer, (world, 1024), torch.float64)
    buf[rank].fill_(float(rank + 1))                    # store into the peer's memory
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = all(bool((t[r] == r + 1).all().item()) for r in range(world))
print(rank, "peer stores visible:", ok, flush=True)
# latency of barrier and of a small NCCL all_gather for comparison
x = torch.zeros((256, 49, 4), dtype=torch.float64, device="cuda")
out = torch.zeros((world * 256, 49, 4), dtype=torch.float64, device="cuda")
for name, fn in (("symm barrier", lambda: hdl.barrier(channel=0)), ("nccl all_gather 400KB", lambda: dist.all_gather_into_tensor(out, x))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(name, a.elapsed_time(b) / 50 * 1e3, "us", flush=True)
dist.barrier()
dist.destroy_process_group()
