import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def host_harness():
    """TEST-ONLY host build of the thread-serial device primitives (g++)."""
    import ctypes
    import subprocess

    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    so = os.path.join(build, "libbp_host_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cpp")
    deps = [src] + [os.path.join(ROOT, "boundplanner_b200", "csrc", f)
                    for f in ("bp_math.cuh", "bp_mvie.cuh", "bp_lp.cuh", "bp_fk.cuh", "bp_mvie_fixed_r.cuh",
                              "bp_mvie_pd.cuh", "bp_planner.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hh_mvie.restype = ctypes.c_int
    lib.hh_pair_lp.restype = ctypes.c_int
    lib.hh_min_eig.restype = ctypes.c_double
    return lib
