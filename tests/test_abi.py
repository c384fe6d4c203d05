"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sim3opt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s3o_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sim3opt_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sim3opt_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype"


def test_no_cpu_fallback_without_device():
    import sim3opt_b200 as s3
    from sim3opt_b200 import _lib
    lib = _lib.load()
    if lib.s3o_device_count() > 0:
        return
    try:
        s3.Problem(s3.KIND_SIM3)
    except s3.S3OError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("s3o_create must fail without a CUDA device")


def test_host_structure_matches_oracle(kitti_k1, kitti_k118, sphere_small):
    from sim3opt_b200 import api
    from conftest import make_oracle
    for g, nb in ((kitti_k1, 1540), (kitti_k118, 1657), (sphere_small, None)):
        cp, ri, h = api.host_structure(len(g["est"]), g["fixed"], g["v0"], g["v1"])
        o = make_oracle(g)
        cp2, ri2 = o.build_structure()
        assert np.array_equal(cp, cp2) and np.array_equal(ri, ri2)       # bit-exact block-CCS
        assert np.array_equal(h, o.hessian_index())
        if nb:
            assert len(ri) == nb


def test_host_structure_edge_cases():
    from sim3opt_b200 import api
    # empty graph
    cp, ri, h = api.host_structure(0, None, [], [])
    assert list(cp) == [0] and len(ri) == 0
    # all vertices fixed -> no free blocks
    cp, ri, h = api.host_structure(3, [1, 1, 1], [0, 1], [1, 2])
    assert list(cp) == [0] and len(ri) == 0 and list(h) == [-1, -1, -1]
    # duplicate edges share one block; an edge to a fixed vertex adds no off-diagonal block
    cp, ri, h = api.host_structure(4, [1, 0, 0, 0], [0, 1, 2, 1, 3], [1, 2, 1, 3, 1])
    assert list(h) == [-1, 0, 1, 2]
    assert list(cp) == [0, 1, 3, 5] and list(ri) == [0, 0, 1, 0, 2]
    # invalid vertex index is rejected
    try:
        api.host_structure(2, None, [0], [5])
    except api.S3OError:
        pass
    else:
        raise AssertionError("invalid edge accepted")


def test_header_is_plain_c_and_integration_snippets_type_check(tmp_path):
    """include/sim3opt_b200.h through a C compiler (the boundary is extern "C", plain pointers and sizes) together
    with the call sequences INTEGRATION.md shows."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "abi_snippets.c")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-c", src, "-o", str(tmp_path / "abi_snippets.o")], check=True)
