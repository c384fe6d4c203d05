"""CPU: the DEVICE math header (sim3opt_b200/csrc/sim3_math.cuh -- what linearize_kernel / chi2_kernel compile), built
for the host with g++, against the oracle's independent restatement of sim3_rv.h (oracle/lie.c): exp / ln in both math
modes, the edge error, the inverse left Jacobian (scalar-coefficient series on the device side, 7x7 matrix series in the
oracle) and the analytic edge Jacobians.  Gives the kernels' arithmetic a gate in the `-m "not gpu"` suite."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from conftest import ROOT

SHIM = os.path.join(ROOT, "tests", "shim", "device_math_shim.cpp")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    so = str(tmp_path_factory.mktemp("dm") / "libdm.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-D__noinline__=", "-I" + CUDA_INC, SHIM, "-o", so], check=True)
    L = C.CDLL(so)
    P = C.POINTER(C.c_double)
    L.dm_exp.argtypes = [P, C.c_int, P]
    L.dm_log.argtypes = [P, C.c_int, P]
    L.dm_edge_error.argtypes = [P, P, P, C.c_int, P]
    L.dm_edge_jacobians.argtypes = [P, P, P, P]
    L.dm_jl_inv.argtypes = [P, P]
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _rand_tangent(rng, ang, trans, sig):
    return np.concatenate([rng.normal(0, ang, 3), rng.normal(0, trans, 3), rng.normal(0, sig, 1)])


@pytest.mark.parametrize("mode", [orc.MATH_REFERENCE, orc.MATH_CORRECTED])
def test_exp_log_and_edge_error_match_the_oracle(dm, mode):
    rng = np.random.default_rng(11)
    orc.set_math_mode(mode)
    try:
        for k in range(300):
            small = k % 3 == 0       # the eps = 1e-5 branches of sim3_rv.h:143-181 / :261-303 as well
            v = _rand_tangent(rng, 1e-6 if small else 0.8, 2.0, 1e-6 if k % 6 == 0 else 0.4)
            S, Sd = orc.sim3_exp(v), np.zeros(8)
            dm.dm_exp(_p(v), int(mode == orc.MATH_CORRECTED), _p(Sd))
            assert np.allclose(Sd, S, rtol=0, atol=1e-13)
            w, wd = orc.sim3_log(S), np.zeros(7)
            dm.dm_log(_p(S), int(mode == orc.MATH_CORRECTED), _p(wd))
            assert np.allclose(wd, w, rtol=0, atol=1e-12)
            Cm, Si, Sj = (orc.sim3_exp(_rand_tangent(rng, 0.7, 3.0, 0.3)) for _ in range(3))
            e, ed = orc.sim3_edge_error(Cm, Si, Sj), np.zeros(7)
            dm.dm_edge_error(_p(Cm), _p(Si), _p(Sj), int(mode == orc.MATH_CORRECTED), _p(ed))
            assert np.allclose(ed, e, rtol=0, atol=1e-11 * max(1.0, np.abs(e).max()))
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)


def test_inverse_left_jacobian_matches_the_oracle_series(dm):
    """The kernels run the series on the scalar coefficients of {I, Om, Om^2} / {Om^a Up Om^b}; the oracle multiplies
    7x7 matrices.  Also the defining property Jl^-1(e) Jl(e) = I through the oracle's adjoint: Jl^-1(e) e = e."""
    rng = np.random.default_rng(12)
    for scale in (1e-7, 1e-3, 0.05, 0.3, 1.0):
        for k in range(100):
            e = _rand_tangent(rng, scale, scale, scale)
            if k % 5 == 0:
                e[6] = 0
            if k % 7 == 0:
                e[:3] = 0
            if k % 11 == 0:
                e[3:6] = 0
            J, Jd = orc.sim3_jl_inv(e), np.zeros(49)
            dm.dm_jl_inv(_p(e), _p(Jd))
            Jd = Jd.reshape(7, 7)
            assert np.abs(Jd - J).max() <= 1e-12 * max(1.0, np.abs(J).max()), (scale, k)
            assert np.abs(Jd @ e - e).max() <= 1e-12 * max(1.0, np.abs(e).max())      # ad_e e = 0


def test_edge_jacobians_match_the_oracle_and_finite_differences(dm):
    rng = np.random.default_rng(13)
    orc.set_math_mode(orc.MATH_CORRECTED)
    try:
        for _ in range(60):
            Cm, Si, Sj = (orc.sim3_exp(_rand_tangent(rng, 0.5, 2.0, 0.2)) for _ in range(3))
            e = orc.sim3_edge_error(Cm, Si, Sj)
            Ji, Jj = np.zeros(49), np.zeros(49)
            dm.dm_edge_jacobians(_p(Cm), _p(e), _p(Ji), _p(Jj))
            Ai, Aj = orc.sim3_edge_jac_analytic(Cm, Si, Sj)
            assert np.abs(Ji.reshape(7, 7) - Ai).max() <= 1e-11 * max(1.0, np.abs(Ai).max())
            assert np.abs(Jj.reshape(7, 7) - Aj).max() <= 1e-11 * max(1.0, np.abs(Aj).max())
            Ni, Nj = orc.sim3_edge_jac_numeric(Cm, Si, Sj, 1e-6)
            assert np.abs(Ji.reshape(7, 7) - Ni).max() <= 1e-6 * max(1.0, np.abs(Ni).max())
            assert np.abs(Jj.reshape(7, 7) - Nj).max() <= 1e-6 * max(1.0, np.abs(Nj).max())
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)
