"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): partitioned solve == single-GPU solve."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precond,graph", [(1, "sphere"), (2, "sphere"), (2, "manhattan")])
def test_partitioned_solve_matches_single_gpu(precond, graph):
    """block-Jacobi and multilevel (replicated coarse levels) PCG; regular and irregular (many cut loop edges) graphs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), "30", "80", "5", str(precond), graph]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ESTIMATES_IDENTICAL_ACROSS_RANKS True" in out.stdout
    assert "UPDATE_PARTITIONED" in out.stdout and "UPDATE_PARTITIONED max diff" in out.stdout and "FAIL" not in out.stdout
    assert "ESTIMATE_SLICES PASS" in out.stdout        # s3o_set_estimates_slice / s3o_get_vertices_slice
    # ghost columns are read over NVLink inside the SpMV (CUDA IPC mappings); a box without peer access falls back to
    # NCCL send/recv on all ranks, which is correct but not what this test is meant to exercise
    assert "P2P_HALO " in out.stdout
    if "P2P_HALO 1" not in out.stdout:
        import warnings
        warnings.warn("peer-to-peer halo not available on this box: the NCCL send/recv fallback was tested instead")
    if precond == 2:
        assert "MULTILEVEL_LEVELS 0" not in out.stdout


def test_partitioned_kcycle_large_graph():
    """30k poses: the deep-hierarchy settings (K-cycle on every level, cooperative kernel, large dense coarsest
    level) in the partitioned solve, replicated coarse levels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), "100", "300", "4", "2", "sphere", "1e-6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ESTIMATES_IDENTICAL_ACROSS_RANKS True" in out.stdout
