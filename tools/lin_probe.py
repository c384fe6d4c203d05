#!/usr/bin/env python
"""Times linearize + assemble alone on one workload (diagnostic): median of 10 s3o_linearize calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sim3opt_b200 as s3
from sim3opt_b200 import synth
wl = {"s10k": (10, 1000), "s100k": (100, 1000), "s1m": (1000, 1000)}[sys.argv[1] if len(sys.argv) > 1 else "s1m"]
g = synth.sphere(*wl, seed=42)
p = s3.Problem(s3.KIND_SIM3); p.set_math_mode(s3.MATH_CORRECTED)
p.set_vertices(g["est"], g["fixed"]); p.set_edges(g["v0"], g["v1"], g["meas"], g["info"]); p.build_structure()
p.linearize_only()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); p.linearize_only(); ts.append(time.perf_counter() - t0)
print(f"linearize+assemble median {np.median(ts)*1e3:.3f} ms  min {min(ts)*1e3:.3f} ms")
