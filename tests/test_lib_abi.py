"""CPU: the C-ABI library builds, loads and exports every symbol include/bpgeo.h declares
(no compute calls -- there is no GPU here), and the product never imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "bpgeo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp\w+|bpgeo\w+)\s*\(", text)))


def test_build_and_exports():
    import __graft_entry__ as ge

    ge.build()
    from boundplanner_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bpgeo.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.load().bpgeo_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_and_oracle_isolation():
    """Product code must not import oracle/ and must fail loudly without a GPU."""
    import pytest
    import torch

    pkg = os.path.join(ROOT, "boundplanner_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
    if not torch.cuda.is_available():
        from boundplanner_b200 import _lib, geometry

        with pytest.raises(_lib.BpGeoError):
            geometry.Scene([[0, 0, 0, 1, 1, 1]], 0.0)


def test_sm100a_sass_present():
    """The shipped library carries sm_100a code for every kernel."""
    import subprocess

    from boundplanner_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
