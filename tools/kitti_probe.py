#!/usr/bin/env python
"""KITTI-00 K118 direct optimisation through the Python mirror (for profiling the sparse block Cholesky)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sim3opt_b200 as s3
from oracle import kitti_io
g = kitti_io.build_kitti_sim3_graph(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kitti00"), len(sys.argv) > 1 and sys.argv[1] == "k1")
p = s3.Problem(s3.KIND_SIM3)
p.set_vertices(g["est"], g["fixed"]); p.set_edges(g["v0"], g["v1"], g["meas"])
p.build_structure()
t0 = time.perf_counter(); n, chi2, lam, hist = p.optimize(20); dt = time.perf_counter() - t0
st = p.stats()
print(f"iterations {n} chi2 {chi2:.9f} wall {dt*1e3:.1f} ms direct_solves {st['direct_solves']} rounds {st['direct_levels']} launches {st['kernel_launches']}")
