"""Shared helpers for the parity tests."""
import numpy as np

from oracle.convex_set_finder import ConvexSetFinder as OracleFinder
from oracle.obstacles import obstacle_reps

# north_star: halfspace coefficients agree within 1e-6 relative (fp64)
RTOL = 1e-6


def oracle_finder(boxes, inflate, ws_min, ws_max, max_rows=None):
    obs_sets, pts, _ = obstacle_reps(boxes, inflate)
    return OracleFinder(obs_sets, pts, ws_max, ws_min, max_rows=max_rows)


def assert_rows_close(A, b, Ao, bo, tag=""):
    assert A.shape == Ao.shape, f"{tag}: row count {A.shape[0]} vs oracle {Ao.shape[0]}"
    scale = max(1.0, np.abs(bo).max())
    assert np.abs(A - Ao).max() <= RTOL, f"{tag}: normals differ {np.abs(A - Ao).max()}"
    assert np.abs(b - bo).max() <= RTOL * scale, f"{tag}: offsets differ {np.abs(b - bo).max()}"
