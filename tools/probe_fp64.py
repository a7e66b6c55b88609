"""GPU microprobe: FP64 DFMA latency / throughput on this part."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from boundplanner_b200 import _lib
lib = _lib.load()
out = torch.empty(148*32*1024, dtype=torch.float64, device='cuda')
def run(chains, blocks, threads, iters):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        lib.bp_probe_fp64(chains, blocks, threads, iters, ctypes.c_void_p(out.data_ptr()), st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    lib.bp_probe_fp64(chains, blocks, threads, iters, ctypes.c_void_p(out.data_ptr()), st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    n = blocks*threads*chains*iters
    return ms, n*2/ms/1e9
it = 100000
ms, _ = run(1, 1, 32, it); print(f"1 warp, 1 chain: {ms*1e6/it:.1f} ns per dependent DFMA = {ms*1e-3*1.965e9/it:.1f} cycles")
ms, _ = run(8, 1, 32, it); print(f"1 warp, 8 chains: {ms*1e6/it/8:.2f} ns per DFMA = {ms*1e-3*1.965e9/it/8:.2f} cycles/instr")
for thr in (128, 256, 1024):
    ms, tf = run(8, 148*2, thr, 20000); print(f"full chip {thr} thr x 296 blocks, 8 chains: {tf:.2f} TFLOP/s")
