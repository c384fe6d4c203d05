import sys, os
sys.path.insert(0, "/root/repo")
import bench
import sim3opt_b200 as s3
from sim3opt_b200 import synth
class A:
    pcg_tol = bench.PCG_TOL; pcg_max_iter = 20000; stop_gain = bench.STOP_REL_GAIN; stop_step = bench.STOP_STEP; precond = "auto"; seed = 42
for pc in ("block-jacobi", "multilevel", "auto"):
    A.precond = pc
    rec = bench.parity_s10k(A, s3, synth, 0)
    print(pc, {k: rec[k] for k in ("lm_iterations", "pcg_iterations", "chi2_rel", "max_translation_m", "max_rotation_rad", "pass")})
