"""Quick device probe: build a synthetic sphere graph, run LM, print phase timings (not a bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sim3opt_b200 as s3
from sim3opt_b200 import synth

laps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
t = time.time(); g = (synth.manhattan3d(laps * per) if os.environ.get("GRAPH") == "manhattan" else synth.sphere(laps, per)); print("gen %.2fs N=%d E=%d" % (time.time() - t, len(g["est"]), len(g["v0"])))
p = s3.Problem(s3.KIND_SIM3)
p.set_math_mode(s3.MATH_CORRECTED)
t = time.time(); p.set_vertices(g["est"], g["fixed"]); p.set_edges(g["v0"], g["v1"], g["meas"], g["info"]); print("upload %.3fs" % (time.time() - t))
t = time.time(); p.build_structure(); print("structure %.3fs" % (time.time() - t), p.num_free, p.num_blocks)
p.set_pcg(tol, 5000)
p.set_preconditioner(int(os.environ.get("PRECOND", "0")))
print("chi2_0", p.chi2())
t = time.time(); n, chi2, lam, hist = p.optimize(iters); wall = time.time() - t
st = p.stats()
print("iters", n, "chi2", chi2, "wall %.3fs" % wall)
print(hist)
print({k: v for k, v in st.items()})
