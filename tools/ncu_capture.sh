set -x
cd $GRAFT_REPO_ROOT
# launch list of two bench steps (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-plan-latency --no-cpu-baseline --no-extras --no-cuda-graph > gpurun_out/r02_launches_bench.log 2>&1
# full capture of the dominant kernel (C2 set build) and of the pair LP / FK kernels
cat > /tmp/drv.py <<'PY'
import numpy as np, torch
from boundplanner_b200 import geometry as geo, scenes
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
sc = geo.Scene(boxes, inflate)
sd = torch.as_tensor(seeds).cuda()
for _ in range(3):
    out = geo.build_sets_point(sc, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
    bits = geo.pair_feasible(out.A, out.b, out.m, 0.01)
q = torch.rand((1 << 20, 7), dtype=torch.float64, device="cuda")
for _ in range(3):
    geo.fk_iiwa14(q)
seeds8 = scenes.free_points(2048, boxes, inflate, np.random.default_rng(7), ws_min, ws_max)
sd8 = torch.as_tensor(seeds8).cuda()
for _ in range(2):
    geo.build_sets_point(sc, sd8, ws_min, ws_max, fixed_mid=True, optimize=True)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 2 -c 1 -o gpurun_out/r02_prof_iris python /tmp/drv.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_iris_fused -s 4 -c 1 -o gpurun_out/r02_prof_iris_sat python /tmp/drv.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_pair_lp|k_pair_filter|k_fk" -s 6 -c 3 -o gpurun_out/r02_prof_pair_fk python /tmp/drv.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
