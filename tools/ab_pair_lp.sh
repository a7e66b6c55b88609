# pair stage after a change of the LP: golden graphs, the stage on one rank's block of the 8-GPU graph and on C4
python -m pytest tests/test_gpu_golden_graph.py tests/test_gpu_graph_fk.py tests/test_gpu_scale.py tests/test_planner_native.py -m gpu -x -q 2>&1 | tail -3
python tools/prof_pairs_block.py 8 2>&1 | grep -v "^$" | head -4
python tools/prof_pairs_block.py 1 2>&1 | grep -v "^$" | head -2
python - <<PY
import os,sys
sys.path.insert(0,".")
import torch
from boundplanner_b200 import geometry as geo, scenes
from boundplanner_b200.distributed import balanced_row_blocks
for S, world in ((256,1),(2048,8)):
    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(n_seeds=S)
    sc = geo.Scene(boxes, inflate)
    aabb = torch.empty((S, 6), dtype=torch.float64, device="cuda")
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True, aabb=aabb)
    r0,r1 = balanced_row_blocks(S,world)[0]
    bufs = geo.alloc_pair_buffers(S, r1-r0)
    for _ in range(3): st = geo.pair_feasible_stages(out.A, out.b, out.m, 0.01, r0, r1, out=bufs, aabb=aabb)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01, r0, r1, out=bufs, aabb=aabb)
    print("S",S,"rank0 block", {k: round(v,4) for k,v in st.items()}, "edges", int(geo.unpack_adjacency(bits, S, r0).sum()))
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
sc = geo.Scene(boxes, inflate); S=2048
aabb = torch.empty((S, 6), dtype=torch.float64, device="cuda")
out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True, aabb=aabb)
bufs = geo.alloc_pair_buffers(S, S)
for _ in range(3): st = geo.pair_feasible_stages(out.A, out.b, out.m, 0.01, 0, S, out=bufs, aabb=aabb)
bits = geo.pair_feasible(out.A, out.b, out.m, 0.01, 0, S, out=bufs, aabb=aabb)
print("C4", {k: round(v,4) for k,v in st.items()}, "edges", int(geo.unpack_adjacency(bits, S).sum()))
PY
