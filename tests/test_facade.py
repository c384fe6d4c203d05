"""The C++ host side: g2o-spelled facade (include/sim3opt_b200/g2o_facade.hpp) + the KITTI pipelines
(examples/kitti_pgo.cpp, mirroring kitti_surf.cpp:542-709 and :713-1086).

CPU: the example builds against the C-ABI library and its loaders reproduce the structural known
answers (1540 / 1657 upper blocks).  GPU: the pipelines run end to end and agree with the same graph
driven through the Python mirror of the ABI and with the CPU oracle.
"""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import KITTI_DIR, ROOT, make_gpu, make_oracle

BIN = os.path.join(ROOT, "examples", "bin", "kitti_pgo")
BA_BIN = os.path.join(ROOT, "examples", "bin", "ba_demo")


@pytest.fixture(scope="module")
def kitti_pgo():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def run(binary, *args):
    r = subprocess.run([binary, *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_dry_run_structure(kitti_pgo):
    out = run(kitti_pgo, "dry-run", KITTI_DIR)
    assert "vertices 771 edges 771 free 770 blocks 1540" in out
    out = run(kitti_pgo, "dry-run", KITTI_DIR, "--all-loops")
    assert "vertices 771 edges 888 free 770 blocks 1657" in out


def test_fails_loudly_without_gpu(kitti_pgo, tmp_path):
    from sim3opt_b200 import _lib
    if _lib.load().s3o_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([kitti_pgo, "direct", KITTI_DIR, str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def read_result(path):
    rows = [l.split() for l in open(path) if not l.startswith("%")]
    return np.array(rows, float)


@pytest.mark.gpu
@pytest.mark.parametrize("all_loops", [False, True])
def test_direct_pipeline_matches_python_path(kitti_pgo, tmp_path, kitti_k1, kitti_k118, all_loops):
    g = kitti_k118 if all_loops else kitti_k1
    out_file = str(tmp_path / "direct.txt")
    args = ["direct", KITTI_DIR, out_file, "--iters", "4", "--precision", "17"] + (["--all-loops"] if all_loops else [])
    out = run(kitti_pgo, *args)
    m = re.search(r"direct: iterations (\d+) free (\d+) blocks (\d+) chi2_first (\S+) chi2_final (\S+)", out)
    assert m, out
    assert int(m.group(2)) == 770 and int(m.group(3)) == (1657 if all_loops else 1540)
    gpu = make_gpu(g, jac=1)
    gpu.set_pcg(1e-13, 100000)
    n, chi2, lam, hist = gpu.optimize(4)
    # same library, same graph (the C++ loaders and the Python loaders agree to round-off):
    # iteration 0 is stable, later iterations amplify the 1e-16 input differences
    assert abs(float(m.group(4)) - hist[0, 0]) <= 1e-6 * hist[0, 0]
    res = read_result(out_file)
    assert res.shape == (771, 9)
    assert np.array_equal(res[:, 0].astype(int), g["frame_ids"])
    assert np.allclose(np.linalg.norm(res[:, 5:9], axis=1), 1.0, atol=1e-12)
    assert res[0, 1] == 1.0                       # the fixed first key frame keeps scale 1


@pytest.mark.gpu
def test_direct_first_iteration_against_oracle(kitti_pgo, tmp_path, kitti_k1):
    from oracle import oracle as orc
    out_file = str(tmp_path / "direct1.txt")
    out = run(kitti_pgo, "direct", KITTI_DIR, out_file, "--iters", "1", "--precision", "17")
    chi2 = float(re.search(r"chi2_final (\S+)", out).group(1))
    cpu = make_oracle(kitti_k1, jac=orc.JAC_ANALYTIC)
    n, chi2_c, lam_c, hist_c = cpu.optimize(1)
    assert abs(chi2 - chi2_c) <= 1e-4 * chi2_c
    # poses written as (s_w2i, t_i_in_w, q_i2w) = components of S_iw^-1
    res = read_result(out_file)
    est = cpu.vertices()
    inv = np.array([orc.sim3_inv(s) for s in est])
    assert np.abs(res[:, 1] - est[:, 7]).max() <= 1e-4
    assert np.abs(res[:, 2:5] - inv[:, 4:7]).max() <= 1e-3 * max(1.0, np.abs(inv[:, 4:7]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("stages", [2, 3])
def test_stepwise_pipeline(kitti_pgo, tmp_path, stages):
    out_file = str(tmp_path / "stepwise.txt")
    out = run(kitti_pgo, "stepwise", KITTI_DIR, out_file, "--stages", str(stages), "--iters", "10")
    assert "scale_dlt: inverse iterations" in out
    m = re.search(r"scale_trans: iterations (\d+) free 770 blocks 1540 chi2_first (\S+) chi2_final (\S+)", out)
    assert m, out
    assert float(m.group(3)) <= float(m.group(2))
    if stages == 3:
        m3 = re.search(r"sim3_optim: iterations (\d+) free 770 blocks 1540 chi2_first (\S+) chi2_final (\S+)", out)
        assert m3, out
        assert float(m3.group(3)) <= float(m3.group(2))
    res = read_result(out_file)
    assert res.shape == (771, 9) and np.isfinite(res).all()


@pytest.mark.gpu
def test_scale_null_vector_matches_dense_svd(kitti_k1, kitti_k118):
    """s3o_smallest_eigenvector against numpy's SVD of the reference's dense constraint matrix
    (kitti_surf.cpp:894-915): rows x[k-1]-x[k] and s_loop x[id1]-x[id2], last right singular vector / v[0]."""
    import sim3opt_b200 as s3
    for g in (kitti_k1, kitti_k118):
        n = len(g["est"])
        v0, v1, s = g["v0"], g["v1"], g["meas"][:, 7]
        A = np.zeros((len(v0), n))
        for r, (i, j, m) in enumerate(zip(v0, v1, s)):
            A[r, i] = m
            A[r, j] = -1.0
        _, sv, Vt = np.linalg.svd(A)
        ref = Vt[-1] / Vt[-1][0]
        p = s3.Problem(s3.KIND_SCALE)
        p.set_vertices(np.ones((n, 1)))
        p.set_edges(v0, v1, s.reshape(-1, 1))
        p.set_pcg(1e-13, 100000)
        x, lmin, lmax, its = p.smallest_eigenvector(60, 1e-13)
        x = x / x[0]
        assert np.abs(x - ref).max() <= 1e-6 * np.abs(ref).max()
        assert abs(np.sqrt(max(lmin, 0)) - sv[-1]) <= 1e-6 * sv[0]
        assert abs(np.sqrt(lmax) - sv[0]) <= 0.05 * sv[0]


@pytest.mark.gpu
def test_ba_demo_matches_oracle(kitti_pgo, tmp_path):
    """examples/ba_demo.cpp (bal_example.cpp:44-243 through the facade) on a synthetic BAL file."""
    from oracle import oracle as orc
    from sim3opt_b200 import synth
    g = synth.ba_loop(30, 900, 6, seed=21)
    bal = str(tmp_path / "problem.txt")
    synth.write_bal(bal, g)
    out_file = str(tmp_path / "cams.txt")
    out = run(BA_BIN, "-i", "8", "-o", out_file, "-v", bal)
    cpu = orc.BAProblem()
    cpu.set(g["cams"], g["points"], g["obs_cam"], g["obs_pt"], g["uv"], g["focal"], g["cx"], g["cy"])
    cpu.set_robust(1, 2.5)
    cpu.build_structure()
    chi0 = cpu.chi2()
    assert abs(float(re.search(r"initial chi2 (\S+)", out).group(1)) - chi0) <= 1e-9 * chi0
    assert f"free cameras 30 schur blocks {cpu.nb}" in out
    n, chi2, lam, hist = cpu.optimize(8)
    m = re.search(r"iterations (\d+) chi2_final (\S+)", out)
    assert int(m.group(1)) == n and abs(float(m.group(2)) - chi2) <= 1e-5 * chi2
    res = read_result(out_file)                      # id, t_c_in_w, q_c2w (xyzw)
    cams = cpu.cameras()
    for k in (0, 13, 29):
        R = orc.quat_to_rot(cams[k, :4])
        assert np.abs(res[k, 1:4] - (-R.T @ cams[k, 4:7])).max() <= 1e-4


@pytest.mark.gpu
def test_align_mode_reports_rmse_against_ground_truth(kitti_pgo, tmp_path, kitti_k1):
    """kitti_pgo align (kitti_surf.cpp:1381-1452): result file of the direct pipeline -> Umeyama alignment to the
    KITTI-00 ground truth; numbers agree with the oracle's alignment of the same estimates."""
    from oracle import oracle as orc, kitti_io
    out_file = str(tmp_path / "direct.txt")
    run(kitti_pgo, "direct", KITTI_DIR, out_file, "--iters", "3", "--precision", "17")
    out = run(kitti_pgo, "align", out_file, os.path.join(KITTI_DIR, "00.txt"))
    m = re.search(r"RMSE and Max deviation (\S+) (\S+)", out)
    m2 = re.search(r"total distance (\S+) ratio of rmse and max error (\S+) (\S+)", out)
    res = read_result(out_file)
    gt = kitti_io.load_kitti_gt_positions(os.path.join(KITTI_DIR, "00.txt"))
    S, rmse, mx = orc.umeyama(res[:, 2:5], gt[res[:, 0].astype(int)])
    assert abs(float(m.group(1)) - rmse) <= 1e-6 * rmse and abs(float(m.group(2)) - mx) <= 1e-6 * mx
    total = np.linalg.norm(np.diff(gt, axis=0), axis=1).sum()
    assert abs(float(m2.group(1)) - total) <= 1e-6 * total
    assert abs(float(m2.group(2)) - rmse / total) <= 1e-6 * rmse / total


def oracle_stepwise(g, iters, stages):
    """testStepwiseSim3Optimization (kitti_surf.cpp:713-1086) on the CPU oracle: dense SVD null vector of the scale
    constraints (:891-915), scale-trans LM (:1021-1022), Sim3 LM from its result (:1044-1045); exact LDL^T, analytic
    Jacobians (what the facade runs by default)."""
    from oracle import kitti_io, oracle as orc
    n = len(g["est"])
    A = np.zeros((len(g["v0"]), n))
    for r, (i, j, m) in enumerate(zip(g["v0"], g["v1"], g["meas"][:, 7])):
        A[r, i] = m
        A[r, j] = -1.0
    Vt = np.linalg.svd(A)[2]
    st = kitti_io.to_scale_trans_graph(g)
    st["est"] = st["est"].copy()
    st["est"][:, 0] = Vt[-1] / Vt[-1][0]
    p = make_oracle(st, kind=orc.KIND_SCALE_TRANS, jac=orc.JAC_ANALYTIC)
    _, chi_st, _, hist_st = p.optimize(iters)
    v = p.vertices()
    est = np.concatenate([g["est"][:, :4], v[:, 1:4], v[:, 0:1]], axis=1)
    chi_s3 = None
    if stages == 3:
        g3 = dict(g)
        g3["est"] = est
        p3 = make_oracle(g3, jac=orc.JAC_ANALYTIC)
        _, chi_s3, _, _ = p3.optimize(iters)
        est = p3.vertices()
    return est, hist_st[0, 0], chi_st, chi_s3


@pytest.mark.gpu
@pytest.mark.parametrize("stages", [2, 3])
def test_stepwise_pipeline_against_oracle_pipeline(kitti_pgo, tmp_path, kitti_k1, stages):
    """Pipeline-level gate: the whole stepwise pipeline through the facade and the device library
    against the same pipeline on the CPU oracle -- every stage's chi2 and the poses that are written out."""
    from oracle import oracle as orc
    iters = 20
    out_file = str(tmp_path / "stepwise_gate.txt")
    out = run(kitti_pgo, "stepwise", KITTI_DIR, out_file, "--stages", str(stages), "--iters", str(iters), "--precision", "17")
    est, chi_st0, chi_st, chi_s3 = oracle_stepwise(kitti_k1, iters, stages)
    m = re.search(r"scale_trans: iterations (\d+) free 770 blocks 1540 chi2_first (\S+) chi2_final (\S+)", out)
    assert m, out
    print("scale_trans chi2: gpu", m.group(2), m.group(3), "oracle", chi_st0, chi_st)
    assert abs(float(m.group(3)) - chi_st) <= 1e-4 * chi_st
    if stages == 3:
        m3 = re.search(r"sim3_optim: iterations (\d+) free 770 blocks 1540 chi2_first (\S+) chi2_final (\S+)", out)
        assert m3, out
        print("sim3 chi2: gpu", m3.group(3), "oracle", chi_s3)
        assert abs(float(m3.group(3)) - chi_s3) <= 1e-4 * chi_s3
    res = read_result(out_file)
    inv = np.array([orc.sim3_inv(s) for s in est])           # written: (s_w2i, t_i_in_w, q_i2w) = components of S_iw^-1
    ds = np.abs(res[:, 1] - est[:, 7]).max()
    dt = np.abs(res[:, 2:5] - inv[:, 4:7]).max()
    dq = np.minimum(np.abs(res[:, 5:9] - inv[:, 0:4]).max(axis=1), np.abs(res[:, 5:9] + inv[:, 0:4]).max(axis=1)).max()
    print(f"stages {stages}: max scale diff {ds:.3e}  translation {dt:.3e} m (path extent {np.abs(inv[:, 4:7]).max():.1f})  quaternion {dq:.3e}")
    assert ds <= 1e-4 and dt <= 1e-4 and dq <= 1e-5
