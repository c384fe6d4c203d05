"""Times the BA path on a synthetic Ladybug-size problem."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sim3opt_b200 as s3
from sim3opt_b200 import synth

C, P, K = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 500000, 10)
t0 = time.time(); g = synth.ba_loop(C, P, K, seed=42); print(f"gen {time.time()-t0:.1f}s obs {len(g['uv'])}")
p = s3.BAProblem()
p.set(g["cams"], g["points"], g["obs_cam"], g["obs_pt"], g["uv"], g["focal"], g["cx"], g["cy"])
p.set_robust(s3.ROBUST_HUBER, 2.5)
p.set_pcg(1e-8, 5000)
t0 = time.time(); p.build_structure(); print(f"structure {time.time()-t0:.1f}s ncf {p.ncf} npf {p.npf} schur blocks {p.nb} contributions {p.ncon}")
print("chi2_0", p.chi2())
t0 = time.time(); n, chi2, lam, hist = p.optimize(10, 1e-6); dt = time.time() - t0
print(f"iters {n} chi2 {chi2} wall {dt:.3f}s"); print(hist)
print(p.stats())
