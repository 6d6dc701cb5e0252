"""CPU: the oracle reproduces every committed golden vector (which came from the reference's own compiled
code, see oracle/gen_golden.py), and the C-ABI library loads and exports every symbol of the header."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("name", ["icm_m8_gauss", "icm_m7_uniform", "icm_m16_gauss"])
def test_oracle_icm_golden(name):
    g = gold(name)
    r = orc.encode_icm(g["X"], g["C"], g["B"], int(g["ilsiter"]), int(g["icmiter"]), int(g["npert"]),
                       bool(g["randord"]), seed=int(g["seed"]), g0=int(g["g0"]), snap_iters=g["snap_iters"])
    assert np.array_equal(r["B"], g["B_out"])
    assert np.array_equal(r["cost"].view(np.uint32), g["cost"].view(np.uint32))
    assert np.array_equal(r["stats"], g["stats"])
    assert np.array_equal(r["B_snap"], g["B_snap"])
    assert np.allclose(r["objs"], g["objs"], rtol=1e-6)


@pytest.mark.parametrize("name", ["scan_lsq_m8", "scan_lsq_m7_ties", "scan_cq_m8", "scan_pq_m8", "scan_pq_m16_ties"])
def test_oracle_scan_golden(name):
    g = gold(name)
    d, i = orc.linscan(int(g["kind"]), g["B"], g["Xq"], g["cb"], int(g["k"]), g.get("nrm"))
    assert np.array_equal(d.view(np.uint32), g["dists"].view(np.uint32))
    assert np.array_equal(i, g["idx"])


def test_oracle_pq_encode_golden():
    g = gold("pq_encode_m8")
    assert np.array_equal(orc.quantize_pq(g["X"], g["Cpq"], int(g["m"])), g["B_out"])


def test_cabi_exports_every_declared_symbol():
    """Every function declared in include/rayuela_b200.h is exported by the built library and bound by the
    Python layer (no compute call: there is no GPU here)."""
    hdr = open(os.path.join(ROOT, "include", "rayuela_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = "\n".join(l for l in hdr.splitlines() if not l.lstrip().startswith("#"))
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", hdr))
    declared -= {"defined"}
    assert {"rayuela_encode_icm", "linscan_aqd_query_extra_byte", "condition", "rayuela_index_search"} <= declared
    from rayuela_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "library does not export " + name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (tests/bench/smoke are the only users)."""
    for path in glob.glob(os.path.join(ROOT, "rayuela.jl_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
            src = open(path, errors="replace").read()
            assert "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, path


def test_no_cpu_fallback_without_library(monkeypatch):
    from rayuela_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librayuela_b200.so")
    with pytest.raises(_lib.RayuelaError):
        _lib.lib()
