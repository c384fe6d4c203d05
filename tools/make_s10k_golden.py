#!/usr/bin/env python
"""Generates tests/golden/s10k_oracle100.npz: the CPU oracle run the way the reference runs its optimiser --
optimize(100), exact sparse LDL^T solves, no stop rule (kitti_surf.cpp:674-675, LinearSolverEigen :553-557) --
on the s10k synthetic sphere graph (sim3opt_b200.synth.sphere(10, 1000, seed=42), corrected math mode).

The fixture pins the answer that bench.py's own settings (multilevel PCG at the bench tolerance, gain stop rule)
must reproduce: tests/test_gpu_bench_parity.py and bench.py's `parity_check` key compare the device result with
it (chi2 1e-4 relative, 1e-4 m, 1e-5 rad: BASELINE.json north_star).  About 6 minutes per Jacobian mode on one core.

  python tools/make_s10k_golden.py            # analytic + numeric (h = 1e-9, g2o linearizeOplus)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc            # noqa: E402
from sim3opt_b200 import synth              # noqa: E402

LAPS, PER, SEED, ITERS = 10, 1000, 42, 100


def run(jac):
    g = synth.sphere(LAPS, PER, seed=SEED)
    orc.set_math_mode(orc.MATH_CORRECTED)
    orc.set_threads(os.cpu_count() or 1)
    p = orc.Problem(orc.KIND_SIM3)
    p.set_vertices(g["est"], g["fixed"])
    p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
    p.set_jacobian_mode(jac, 1e-9)
    p.build_structure()
    t0 = time.time()
    n, chi2, lam, hist = p.optimize(ITERS, 0.0)
    dt = time.time() - t0
    print(f"jac={jac}: {n} iterations in {dt:.1f} s, chi2={chi2!r}, lambda={lam:.3e}", flush=True)
    return n, chi2, lam, np.asarray(hist), p.vertices(), dt


def main():
    out = {}
    for name, jac in (("analytic", orc.JAC_ANALYTIC), ("numeric", orc.JAC_NUMERIC)):
        n, chi2, lam, hist, est, dt = run(jac)
        out[f"{name}_iterations"] = n
        out[f"{name}_chi2"] = chi2
        out[f"{name}_lambda"] = lam
        out[f"{name}_hist"] = hist
        out[f"{name}_est"] = est
        out[f"{name}_seconds"] = dt
    out["laps"], out["per"], out["seed"] = LAPS, PER, SEED
    path = os.path.join(ROOT, "tests", "golden", "s10k_oracle100.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
