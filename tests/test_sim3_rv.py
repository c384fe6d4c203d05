"""RobotVision::Sim3 (include/sim3opt_b200/sim3_rv.hpp, the reference's sim3_rv.h:71-320 convention with
tangent order [upsilon, omega, sigma]) against the CPU oracle's exp / log / compose / inverse, which hold
the same math in g2o's order [omega, upsilon, sigma].  Host code only: built with plain g++."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sim3rv") / "sim3_rv_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "sim3_rv_check.cpp"), "-o", out], check=True)
    return out


def _cases():
    rng = np.random.default_rng(4)
    v = np.concatenate([rng.normal(0, 2.0, (40, 3)), rng.normal(0, 0.6, (40, 3)), rng.normal(0, 0.4, (40, 1))], axis=1)
    v[5, 3:6] = 0                      # theta < eps, sigma != 0  (the branch with the as-written B)
    v[6, 3:6] = 1e-7
    v[7, 6] = 0                        # |sigma| < eps
    v[8, 6] = 1e-7
    v[9, 3:] = 0                       # both small
    v[10, 3:6] = [3.0, 0.2, -0.4]      # large angle
    return v


@pytest.mark.parametrize("mode", ["reference", "corrected"])
def test_exp_ln_compose_match_oracle(checker, mode):
    from oracle import oracle as orc
    v = _cases()
    text = "\n".join(" ".join(repr(float(x)) for x in row) for row in v) + "\n"
    out = subprocess.run([checker] + (["c"] if mode == "corrected" else []), input=text, capture_output=True, text=True,
                         check=True).stdout
    got = np.array([[float(x) for x in line.split()] for line in out.strip().splitlines()])
    assert got.shape == (len(v), 13 + 7 + 13)
    orc.set_math_mode(orc.MATH_CORRECTED if mode == "corrected" else orc.MATH_REFERENCE)
    try:
        prev = np.array([0, 0, 0, 1, 0, 0, 0, 1.0])
        for k, row in enumerate(v):
            g2o_order = np.concatenate([row[3:6], row[0:3], row[6:7]])       # [omega, upsilon, sigma]
            S = orc.sim3_exp(g2o_order)
            R = orc.quat_to_rot(S[:4])
            assert np.abs(got[k, 0:9].reshape(3, 3) - R).max() <= 1e-12
            assert np.abs(got[k, 9:12] - S[4:7]).max() <= 1e-12 * max(1.0, np.abs(S[4:7]).max())
            assert abs(got[k, 12] - S[7]) <= 1e-14 * S[7]
            back = orc.sim3_log(S)                                           # oracle log of the oracle exp
            mine = np.concatenate([got[k, 16:19], got[k, 13:16], got[k, 19:20]])
            assert np.abs(mine - back).max() <= 1e-9 * max(1.0, np.abs(back).max())
            rel = orc.sim3_mul(S, orc.sim3_inv(prev))
            assert np.abs(got[k, 20:29].reshape(3, 3) - orc.quat_to_rot(rel[:4])).max() <= 1e-12
            assert np.abs(got[k, 29:32] - rel[4:7]).max() <= 1e-11 * max(1.0, np.abs(rel[4:7]).max())
            assert abs(got[k, 32] - rel[7]) <= 1e-13 * rel[7]
            prev = S
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)
