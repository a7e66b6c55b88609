"""Small workloads for compute-sanitizer (memcheck / racecheck) on the GPU box: the fused set build on boxes and on
polytopes (cached and re-solving form), the segment sets, the pair stage, a short native planner run, and (argument `spec`) the SPEC instantiation of the fused set build.
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.spatial.transform import Rotation as R
from boundplanner_b200 import geometry as geo, scenes, planner_native as pn

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "spec":
    os.environ["BPGEO_SPEC"] = "1"          # read once, at the library's first set build
boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2(300, 6)
sc = geo.Scene(boxes, inflate)
if which in ("all", "sets"):
    aabb = torch.empty((6, 6), dtype=torch.float64, device="cuda")
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True, aabb=aabb)
    geo.pair_feasible(out.A, out.b, out.m, 0.01, aabb=aabb)
    geo.build_sets_line(sc, seeds, seeds + 0.03, ws_min, ws_max, compute_ellipsoid=True)
    geo.reduce_ineqs(out.A, out.b, out.m)
    print("sets ok", out.status.cpu().numpy())
if which == "spec":
    # SPEC instantiation of k_iris_fused (speculative free-centre solve on a second warp, shared abort flag):
    # the same seeds must give the same sets as the plain kernel's golden rows, checked by the GPU tests; here
    # the point is memcheck / racecheck over the two solver warps
    out = geo.build_sets_point(sc, seeds, ws_min, ws_max, fixed_mid=True, optimize=True)
    print("spec ok", out.status.cpu().numpy(), out.m.cpu().numpy())
if which in ("all", "poly"):
    rng = np.random.default_rng(1)
    obs_sets, obs_points = scenes.random_polytope_scene(60, rng, 0.04, 0.16)
    sd = scenes.polytope_free_points(3, obs_sets, 0.03, rng)
    sp = geo.PolytopeScene(obs_sets, obs_points)
    o1 = geo.build_sets_point(sp, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
    geo.build_sets_line(sp, sd, sd + 0.03, ws_min, ws_max, compute_ellipsoid=True)
    # re-solving form (> 3072 obstacles): the same 60 plus far-away copies
    from oracle.obstacles import obstacle_reps
    far = np.tile(np.array([[40.0, 40.0, 40.0, 40.1, 40.1, 40.1]]), (3100, 1)) + np.arange(3100)[:, None] * 0.2
    os_far, pts_far, _ = obstacle_reps(far, 0.0)
    big = geo.PolytopeScene(list(obs_sets) + list(os_far), list(obs_points) + list(pts_far))
    o2 = geo.build_sets_point(big, sd, ws_min, ws_max, fixed_mid=True, optimize=True)
    l2 = geo.build_sets_line(big, sd, sd + 0.03, ws_min, ws_max, compute_ellipsoid=True)
    print("poly ok", torch.equal(o1.A, o2.A), o2.status.cpu().numpy())
if which in ("all", "plan"):
    r0 = R.from_euler("XYZ", [0, 90, 0], degrees=True).as_matrix()
    queries = []
    for i in (0, 7):
        ob, infl, st, en, wmin, wmax = scenes.config_c3_query(i)
        queries.append(dict(obstacles=ob, start=st, end=en, r0=r0, r1=r0))
    res, stats = pn.plan_batch_native(queries, infl, list(wmax), list(wmin), rng_seeds=[0, 7])
    print("plan ok", stats["rounds"], [r if isinstance(r, Exception) else r["set_ids"] for r in res])
torch.cuda.synchronize()
