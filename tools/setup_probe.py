#!/usr/bin/env python
"""Times the set-up calls (upload, structure, hierarchy) of one workload separately (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sim3opt_b200 as s3
from sim3opt_b200 import synth
wl = {"s10k": (10, 1000), "s100k": (100, 1000), "s1m": (1000, 1000)}[sys.argv[1] if len(sys.argv) > 1 else "s1m"]
t0 = time.perf_counter(); g = synth.sphere(*wl, seed=42); print(f"generate {time.perf_counter()-t0:.2f} s")
p = s3.Problem(s3.KIND_SIM3); p.set_math_mode(s3.MATH_CORRECTED)
for name, f in (("set_vertices", lambda: p.set_vertices(g["est"], g["fixed"])),
                ("set_edges", lambda: p.set_edges(g["v0"], g["v1"], g["meas"], g["info"])),
                ("build_structure", lambda: p.build_structure()),
                ("first optimize(1)", lambda: p.optimize(1)), ("second optimize(1)", lambda: p.optimize(1))):
    t0 = time.perf_counter(); f(); print(f"{name}: {time.perf_counter()-t0:.3f} s", flush=True)
