"""Times the BA path on a synthetic Ladybug-size problem: PCG vs the exact Schur solve."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sim3opt_b200 as s3
from sim3opt_b200 import synth

C, P, K = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 500000, 10)
t0 = time.time(); g = synth.ba_loop(C, P, K, seed=42); print(f"gen {time.time()-t0:.1f}s obs {len(g['uv'])}")
for solver in (s3.LINSOLVER_PCG, s3.LINSOLVER_DIRECT, s3.LINSOLVER_AUTO):
    p = s3.BAProblem()
    p.set(g["cams"], g["points"], g["obs_cam"], g["obs_pt"], g["uv"], g["focal"], g["cx"], g["cy"])
    p.set_robust(s3.ROBUST_HUBER, 2.5)
    p.set_pcg(1e-8, 5000)
    p.set_linear_solver(solver)
    t0 = time.time(); p.build_structure(); t_s = time.time() - t0
    p.snapshot_estimates()
    p.optimize(2, 0.0)
    p.restore_estimates()
    t0 = time.time(); n, chi2, lam, hist = p.optimize(10, 0.0); dt = time.time() - t0
    st = p.stats()
    print(f"solver {solver}: structure {t_s:.2f}s iters {n} chi2 {chi2:.6f} wall {dt:.3f}s ({1e3*dt/max(n,1):.1f} ms/it) pcg {int(hist[:,4].sum())} "
          f"direct_levels {st['direct_levels']} direct_blocks {st['direct_blocks']} schur blocks {p.nb}", flush=True)
