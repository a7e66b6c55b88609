"""boundplanner_b200 -- B200-native geometry core for Thieso/BoundPlanner.

Hand-written sm_100a CUDA kernels (csrc/) behind a C ABI (include/bpgeo.h),
exposed through the reference's own Python surface:

* ``ConvexSetFinder``   -- bound_planner/BoundPlanner/ConvexSetFinder.py
* ``set_intersection``  -- BoundPlanner.set_intersection (BoundPlanner.py:774-787)
* ``RobotModel``        -- numeric FK of bound_planner/RobotModel/RobotModel.py
* ``reduce_ineqs``      -- bound_planner/utils/util_functions.py:82-88
* ``compute_polytope_vertices`` -- bound_planner/utils/util_functions.py:66-79
* ``geometry``          -- the batched device API underneath

Importing the package does not touch the GPU; the first call does, and fails
loudly if libbpgeo.so or a CUDA device is missing (there is no CPU fallback).
"""
__all__ = ["ConvexSetFinder", "RobotModel", "set_intersection", "adjacency", "reduce_ineqs", "normalize_set_size",
           "compute_polytope_vertices", "geometry", "scenes"]


def __getattr__(name):
    if name == "ConvexSetFinder":
        from .convex_set_finder import ConvexSetFinder
        return ConvexSetFinder
    if name == "RobotModel":
        from .robot_model import RobotModel
        return RobotModel
    if name in ("set_intersection", "adjacency"):
        from . import set_graph
        return getattr(set_graph, name)
    if name in ("reduce_ineqs", "normalize_set_size", "compute_polytope_vertices", "obstacle_points_sets"):
        from . import utils
        return getattr(utils, name)
    if name in ("geometry", "scenes", "distributed", "set_graph"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
