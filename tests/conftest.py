import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KITTI_DIR = os.path.join(ROOT, "tests", "golden", "kitti00")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The CPU oracle and the C-ABI library must exist before any test imports them."""
    import __graft_entry__ as entry
    entry.build_oracle()
    if not os.path.exists(os.path.join(ROOT, "sim3opt_b200", "lib", "libsim3opt_b200.so")):
        entry.build()


@pytest.fixture(scope="session")
def kitti_k1():
    from oracle import kitti_io
    return kitti_io.build_kitti_sim3_graph(KITTI_DIR, use_one_constraint=True)


@pytest.fixture(scope="session")
def kitti_k118():
    from oracle import kitti_io
    return kitti_io.build_kitti_sim3_graph(KITTI_DIR, use_one_constraint=False)


@pytest.fixture(scope="session")
def sphere_small():
    from sim3opt_b200 import synth
    return synth.sphere(n_laps=12, poses_per_lap=40, seed=7)


@pytest.fixture(scope="session")
def manhattan_small():
    from sim3opt_b200 import synth
    return synth.manhattan3d(500, seed=5)


def make_oracle(g, kind=None, jac=None, robust=None):
    from oracle import oracle as orc
    p = orc.Problem(orc.KIND_SIM3 if kind is None else kind)
    p.set_vertices(g["est"], g["fixed"], g.get("aux"))
    p.set_edges(g["v0"], g["v1"], g["meas"], g.get("info"))
    if jac is not None:
        p.set_jacobian_mode(jac)
    if robust is not None:
        p.set_robust(*robust)
    return p


def make_gpu(g, kind=None, jac=None, robust=None, math_mode=None):
    import sim3opt_b200 as s3
    p = s3.Problem(s3.KIND_SIM3 if kind is None else kind)
    p.set_vertices(g["est"], g["fixed"], g.get("aux"))
    p.set_edges(g["v0"], g["v1"], g["meas"], g.get("info"))
    if jac is not None:
        p.set_jacobian_mode(jac)
    if robust is not None:
        p.set_robust(*robust)
    if math_mode is not None:
        p.set_math_mode(math_mode)
    return p
