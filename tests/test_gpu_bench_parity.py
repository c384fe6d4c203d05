"""Parity at the benchmark's OWN settings (bench.py: PCG forcing tolerance, stop rule, preconditioner AUTO, corrected
math mode, analytic Jacobians) against the way the reference runs its optimiser: optimize(100) with exact LDL^T
solves and no stop rule (kitti_surf.cpp:674-675, LinearSolverEigen :553-557), restated by the CPU oracle.

The oracle's answer on the s10k sphere graph is the committed fixture tests/golden/s10k_oracle100.npz (about 95 s of
CPU per Jacobian mode; tools/make_s10k_golden.py).  Gates are BASELINE.json's: final chi2 1e-4 relative, per-pose
translation 1e-4 m, rotation 1e-5 rad.
"""
import os

import numpy as np
import pytest

import bench

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s10k_oracle100.npz")


class _Args:
    pcg_tol = bench.PCG_TOL
    pcg_max_iter = 20000
    stop_gain = bench.STOP_REL_GAIN
    stop_step = bench.STOP_STEP
    precond = "auto"
    seed = 42


def test_golden_fixture_is_the_oracles_fixed_point():
    """CPU: the stored estimate reproduces the stored chi2, and one more exact LM iteration does not move it."""
    from oracle import oracle as orc
    from sim3opt_b200 import synth
    z = np.load(GOLD)
    g = synth.sphere(int(z["laps"]), int(z["per"]), seed=int(z["seed"]))
    orc.set_math_mode(orc.MATH_CORRECTED)
    try:
        p = orc.Problem(orc.KIND_SIM3)
        p.set_vertices(z["analytic_est"], g["fixed"])
        p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
        p.set_jacobian_mode(orc.JAC_ANALYTIC)
        chi = p.chi2()
        assert abs(chi - float(z["analytic_chi2"])) <= 1e-12 * chi
        # numeric (g2o-faithful) and analytic runs agree in chi2 to 1e-8 although their poses differ by millimetres
        assert abs(float(z["numeric_chi2"]) - float(z["analytic_chi2"])) <= 1e-8 * chi
        # stationarity: the gradient at the stored estimate is at round-off level relative to the first iteration's
        H, b = p.linearize()
        p0 = orc.Problem(orc.KIND_SIM3)
        p0.set_vertices(g["est"], g["fixed"])
        p0.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
        p0.set_jacobian_mode(orc.JAC_ANALYTIC)
        H0, b0 = p0.linearize()
        assert np.abs(b).max() <= 1e-7 * np.abs(b0).max()
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)


@pytest.mark.gpu
def test_bench_settings_reach_the_reference_answer():
    import sim3opt_b200 as s3
    from sim3opt_b200 import synth
    rec = bench.parity_s10k(_Args, s3, synth, 0)
    assert rec["chi2_rel"] <= bench.TOL_CHI2, rec
    assert rec["max_translation_m"] <= bench.TOL_TRANS, rec
    assert rec["max_rotation_rad"] <= bench.TOL_ROT, rec
    assert rec["pass"]
    # the round-1 settings (1e-6 gain rule) stop an order of magnitude of metres short of it: keep that visible
    class Old(_Args):
        stop_gain = 1e-6
        stop_step = 0.0
        pcg_tol = 1e-3
    old = bench.parity_s10k(Old, s3, synth, 0)
    assert old["max_translation_m"] > 10 * bench.TOL_TRANS


@pytest.mark.gpu
@pytest.mark.parametrize("precond", ["multilevel", "block-jacobi"])
def test_bench_settings_parity_both_preconditioners(precond):
    import sim3opt_b200 as s3
    from sim3opt_b200 import synth

    class A(_Args):
        pass
    A.precond = precond
    rec = bench.parity_s10k(A, s3, synth, 0)
    assert rec["pass"], rec
