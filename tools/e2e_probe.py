"""Times every C-ABI call of bench.py's e2e loop separately (diagnostic, not a bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sim3opt_b200 as s3
from sim3opt_b200 import synth

laps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = synth.sphere(laps, 1000, seed=42)
nv = len(g["est"])
p = s3.Problem(s3.KIND_SIM3)
p.set_math_mode(s3.MATH_CORRECTED)
p.set_pcg(1e-3, 2000)
p.set_vertices(g["est"], g["fixed"]); p.set_edges(g["v0"], g["v1"], g["meas"], g["info"]); p.build_structure()
p.snapshot_estimates()
p.set_lm_resume(2)
est = torch.empty((nv, 8), dtype=torch.float64).pin_memory().numpy()
est0 = torch.empty((nv, 8), dtype=torch.float64).pin_memory().numpy()
p.vertices(out=est0)
def T(f, *a):
    t = time.perf_counter(); r = f(*a); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3
for rep in range(2):
    for k in range(5):
        if k == 0:
            _, t_r = T(p.restore_estimates); _, t_s = T(p.set_estimates, est0)
        else:
            t_r = 0; _, t_s = T(p.set_estimates, est)
        (n, chi2, lam, h), t_o = T(p.optimize, 1, 0.0)
        st = p.stats()
        _, t_g = T(p.vertices, est)
        print(f"rep {rep} step {k}: restore {t_r:.1f} set {t_s:.1f} optimize {t_o:.1f} (device ms_total {st['ms_total']:.1f} lin {st['ms_linearize']:.1f} solve {st['ms_solve']:.1f} upd {st['ms_update']:.1f}) get {t_g:.1f} pcg {int(h[0][4])}", flush=True)
