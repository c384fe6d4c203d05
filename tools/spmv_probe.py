"""Times the SpMV kernel variants on a synthetic sphere (env S3O_SPMV_VERSION / S3O_SPMV4_CFG picks the kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sim3opt_b200 as s3
from sim3opt_b200 import synth

laps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = synth.sphere(laps, 1000, seed=42)
p = s3.Problem(s3.KIND_SIM3)
p.set_math_mode(s3.MATH_CORRECTED)
p.set_preconditioner(s3.PRECOND_BLOCK_JACOBI)
p.set_vertices(g["est"], g["fixed"]); p.set_edges(g["v0"], g["v1"], g["meas"], g["info"]); p.build_structure()
p.linearize_only()
lam = 1e-5 * p.max_diag()
# correctness of the product against a reference version is covered by pytest; here: timing + checksum
rng = np.random.default_rng(0)
x = rng.normal(size=p.num_free * 7)
y = p.hessian_multiply(lam, x)
print("version", os.environ.get("S3O_SPMV_VERSION", "default"), "cfg", os.environ.get("S3O_SPMV4_CFG", "-"),
      "checksum %.12e" % float(np.dot(y, x)), "abs %.12e" % float(np.abs(y).sum()))
p.set_pcg(1e-30, 400)
p.reset_stats()
t = time.time(); rc, xs, it, rel = p.solve(lam); wall = time.time() - t
st = p.stats()
print("pcg iters", it, "wall %.3f s  => %.3f ms/iter;" % (wall, 1e3 * wall / max(it, 1)),
      "spmv sampled avg %.4f ms over %d" % (st["ms_spmv_sampled"] / max(st["n_spmv_sampled"], 1), st["n_spmv_sampled"]))
