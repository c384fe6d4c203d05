#!/usr/bin/env python
"""Parity of the LM result at candidate bench settings against the oracle's optimize(100) answer on s10k
(tests/golden/s10k_oracle100.npz).  Prints, per setting and per LM iteration: chi2, PCG iterations, device ms,
and the distance of the estimate from the oracle's final one (chi2 rel, max translation m, max rotation rad).

  python tools/parity_probe.py [--workload s10k|s100k|s1m] [--iters 16] [--tols 1e-3,1e-4]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sim3opt_b200 as s3                  # noqa: E402
from sim3opt_b200 import synth             # noqa: E402

WL = {"s10k": (10, 1000), "s100k": (100, 1000), "s1m": (1000, 1000)}


def pose_diff(a, b):
    dt = np.abs(a[:, 4:7] - b[:, 4:7]).max()
    dots = np.abs((a[:, :4] * b[:, :4]).sum(1) / (np.linalg.norm(a[:, :4], axis=1) * np.linalg.norm(b[:, :4], axis=1)))
    dr = (2 * np.arccos(np.clip(dots, -1, 1))).max()
    ds = np.abs(a[:, 7] - b[:, 7]).max()
    return dt, dr, ds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="s10k")
    ap.add_argument("--iters", type=int, default=16)
    ap.add_argument("--tols", default="1e-3")
    ap.add_argument("--max-pcg", type=int, default=20000)
    ap.add_argument("--forcing", type=int, default=0)
    args = ap.parse_args()
    laps, per = WL[args.workload]
    g = synth.sphere(laps, per, seed=42)
    gold = None
    if args.workload == "s10k":
        z = np.load(os.path.join(ROOT, "tests", "golden", "s10k_oracle100.npz"))
        gold = (float(z["analytic_chi2"]), z["analytic_est"])
    for tol in [float(t) for t in args.tols.split(",")]:
        p = s3.Problem(s3.KIND_SIM3)
        p.set_math_mode(s3.MATH_CORRECTED)
        p.set_pcg(tol, args.max_pcg)
        if args.forcing and hasattr(p, "set_forcing"):
            p.set_forcing(args.forcing)
        p.set_vertices(g["est"], g["fixed"])
        p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])
        t0 = time.perf_counter()
        p.build_structure()
        print(f"== {args.workload} tol {tol:g}: setup {time.perf_counter() - t0:.2f} s", flush=True)
        p.set_lm_resume(True)
        prev = None
        tsum = 0.0
        for it in range(args.iters):
            t0 = time.perf_counter()
            n, chi2, lam, hist = p.optimize(1, 0.0)
            dt = time.perf_counter() - t0
            tsum += dt
            st = p.stats()
            line = (f"it {it:2d} chi2 {chi2:.9f} lambda {lam:.3e} trials {int(hist[0][2])} pcg {int(hist[0][4])} "
                    f"ms {dt * 1e3:8.1f} (lin {st['ms_linearize']:.1f} solve {st['ms_solve']:.1f}) cum {tsum:.3f}s "
                    f"step {st['last_step_inf']:.2e} est {st['est_distance']:.2e}")
            if prev is not None:
                line += f" gain {(prev - chi2) / chi2:.2e}"
            if gold is not None:
                dtm, dr, ds = pose_diff(p.vertices(), gold[1])
                line += f" | vs oracle100: chi2 rel {(chi2 - gold[0]) / gold[0]:+.2e} trans {dtm:.2e} m rot {dr:.2e} rad scale {ds:.2e}"
            print(line, flush=True)
            prev = chi2
        st = p.stats()
        print("unconverged PCG solves:", st["pcg_unconverged"], "| coarse operators rebuilt", st["multilevel_rebuilds"],
              "kept", st["multilevel_reuses"], flush=True)


if __name__ == "__main__":
    main()
