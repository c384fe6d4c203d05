"""Device-side structure build (csrc/structure_dev.cu: stable radix sort + scans) against the host twin
(csrc/structure.cpp) and the oracle: the g2o-order block-CCS, the Hessian indices and -- through a linearisation --
every internal index array (edge permutation, block ranges, incidence lists, column view) must be identical."""
import os

import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def _build(g, where, **kw):
    os.environ["S3O_STRUCTURE"] = where
    try:
        p = make_gpu(g, jac=1, **kw)
        colptr, rowidx = p.build_structure()
        hidx = p.hessian_index()
        H, b = p.linearize()
        return p, colptr, rowidx, hidx, H, b
    finally:
        del os.environ["S3O_STRUCTURE"]


@pytest.fixture(params=["kitti_k1", "kitti_k118", "sphere_small", "manhattan_small"])
def graph(request):
    return request.getfixturevalue(request.param)


def test_device_structure_bit_exact(graph):
    _, cp_h, ri_h, hx_h, H_h, b_h = _build(graph, "host")
    _, cp_d, ri_d, hx_d, H_d, b_d = _build(graph, "device")
    assert np.array_equal(cp_h, cp_d) and np.array_equal(ri_h, ri_d) and np.array_equal(hx_h, hx_d)
    # same edge order, same block ranges, same incidence lists => the assembled system is the same bits
    assert np.array_equal(H_h, H_d) and np.array_equal(b_h, b_d)
    cpu = make_oracle(graph)
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(cp_d, cp_c) and np.array_equal(ri_d, ri_c)


def test_device_structure_ragged_graph():
    """Duplicate edges (multi-edge blocks), a hub vertex, fixed vertices in the middle, edges between fixed vertices,
    one-free-end edges, an isolated free vertex."""
    rng = np.random.default_rng(3)
    n = 400
    est = np.tile(np.array([0, 0, 0, 1, 0, 0, 0, 1.0]), (n, 1))
    est[:, 4:7] = rng.standard_normal((n, 3))
    fixed = np.zeros(n, np.uint8)
    fixed[[0, 17, 18, 200]] = 1
    v0 = list(range(1, n - 1)) + [5] * 60 + [17, 30, 30, 30, 250]
    v1 = list(range(0, n - 2)) + list(range(100, 160)) + [18, 31, 31, 29, 17]
    v0, v1 = np.array(v0, np.int32), np.array(v1, np.int32)          # vertex n-1 stays isolated
    meas = np.tile(np.array([0, 0, 0, 1, 0.1, 0, 0, 1.0]), (len(v0), 1))
    g = dict(est=est, fixed=fixed, v0=v0, v1=v1, meas=meas)
    _, cp_h, ri_h, hx_h, H_h, b_h = _build(g, "host")
    _, cp_d, ri_d, hx_d, H_d, b_d = _build(g, "device")
    assert np.array_equal(cp_h, cp_d) and np.array_equal(ri_h, ri_d) and np.array_equal(hx_h, hx_d)
    assert np.array_equal(H_h, H_d) and np.array_equal(b_h, b_d)
    cpu = make_oracle(g)
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(cp_d, cp_c) and np.array_equal(ri_d, ri_c)


def test_device_structure_large_and_fast():
    """100k poses: identical to the host twin, and the device build is not the slow part any more."""
    import time
    from sim3opt_b200 import synth
    g = synth.sphere(100, 1000, seed=42)
    t0 = time.perf_counter()
    _, cp_h, ri_h, hx_h, H_h, b_h = _build(g, "host")
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, cp_d, ri_d, hx_d, H_d, b_d = _build(g, "device")
    t_dev = time.perf_counter() - t0
    assert np.array_equal(cp_h, cp_d) and np.array_equal(ri_h, ri_d) and np.array_equal(hx_h, hx_d)
    assert np.array_equal(H_h, H_d) and np.array_equal(b_h, b_d)
    print(f"structure + linearize: host {t_host:.3f} s, device {t_dev:.3f} s")
