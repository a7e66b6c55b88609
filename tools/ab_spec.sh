for SP in 0 1; do
echo "== BPGEO_SPEC=$SP"
BPGEO_SPEC=$SP timeout 600 python -m pytest tests/test_gpu_golden_graph.py tests/test_gpu_sets.py -x -q -k "not sweep" 2>&1 | tail -2
BPGEO_SPEC=$SP timeout 300 python - <<PY
import torch, time, numpy as np
from boundplanner_b200 import geometry as geo, scenes
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate); sd=torch.as_tensor(seeds).cuda(); out=geo.alloc_set_batch(256)
print("C2 256 seeds", round(t(lambda: geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True, out=out)),4))
s8 = scenes.free_points(2048, boxes, inflate, np.random.default_rng(7), ws_min, ws_max); sd8=torch.as_tensor(s8).cuda(); out8=geo.alloc_set_batch(2048)
print("C2 scene 2048 seeds", round(t(lambda: geo.build_sets_point(sc, sd8, ws_min, ws_max, fixed_mid=True, optimize=True, out=out8)),4))
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c4()
sc4 = geo.Scene(boxes, inflate); sd4=torch.as_tensor(seeds).cuda()
print("C4 2048 seeds", round(t(lambda: geo.build_sets_point(sc4, sd4, ws_min, ws_max, fixed_mid=True, optimize=True, out=out8)),4))
PY
done
