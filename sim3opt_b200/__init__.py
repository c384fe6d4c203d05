"""sim3opt_b200 -- B200-native Sim3/SE3 nonlinear least-squares back-end (host-side Python mirror).

The product is the C-ABI shared library sim3opt_b200/lib/libsim3opt_b200.so (include/sim3opt_b200.h);
this package only binds it for tests and bench.py.
"""
from .api import (Problem, BAProblem, LinearSolver, S3OError, KIND_SIM3, KIND_SCALE_TRANS, KIND_SCALE, KIND_BA, JAC_NUMERIC, JAC_ANALYTIC, MATH_REFERENCE, MATH_CORRECTED, PRECOND_AUTO, PRECOND_BLOCK_JACOBI, PRECOND_MULTILEVEL, LINSOLVER_AUTO, LINSOLVER_PCG, LINSOLVER_DIRECT, SCALE_MODEL_DIFFERENCE, SCALE_MODEL_LOGRATIO, host_structure, host_partition, host_multilevel, host_direct_plan, align_similarity, comm_unique_id,
                  ROBUST_NONE, ROBUST_HUBER, ROBUST_PTAM_TUKEY, ROBUST_PTAM_CAUCHY, ROBUST_PTAM_HUBER,
                  ROBUST_PTAM_LS)

__all__ = ["Problem", "BAProblem", "LinearSolver", "S3OError"]
