"""Trajectory alignment and RMSE (kitti_surf.cpp:1091-1161, :1432-1452): the oracle's
Umeyama restatement on known answers, and the device reduction against it (KITTI-00 ground truth included)."""
import os

import numpy as np
import pytest

from conftest import KITTI_DIR


def _random_similarity(rng):
    from scipy.spatial.transform import Rotation
    R = Rotation.from_rotvec(rng.normal(0, 1.0, 3)).as_matrix()
    return rng.uniform(0.2, 5.0), R, rng.normal(0, 20, 3)


def test_oracle_umeyama_recovers_a_known_similarity():
    from oracle import oracle as orc
    rng = np.random.default_rng(2)
    q = rng.normal(0, 30, (500, 3))
    c, R, t = _random_similarity(rng)
    train = c * q @ R.T + t
    S, rmse, mx = orc.umeyama(q, train)
    assert np.abs(S[:3, :3] - c * R).max() <= 1e-12 * c and np.abs(S[:3, 3] - t).max() <= 1e-10
    assert rmse <= 1e-11 and mx <= 1e-10
    # with noise the residual is the noise level, and a reflection is never returned
    S2, rmse2, _ = orc.umeyama(q, train + rng.normal(0, 0.1, train.shape))
    assert 0.15 <= rmse2 <= 0.2 and np.linalg.det(S2[:3, :3]) > 0
    # only-scale variant of the reference: extent ratio of x and z
    S3, _, _ = orc.umeyama(q, 2.5 * q, only_scale=True)
    assert np.allclose(S3[:3, :3], 2.5 * np.eye(3)) and np.allclose(S3[:3, 3], 0)


def test_kitti_vo_trajectory_against_ground_truth(kitti_k1):
    """Known answer on the fixtures: the monocular VO key-frame trajectory (scale drift, no loop closure yet)
    aligned to the KITTI-00 ground truth."""
    from oracle import oracle as orc, kitti_io
    gt = kitti_io.load_kitti_gt_positions(os.path.join(KITTI_DIR, "00.txt"))
    assert gt.shape == (4541, 3)
    ids = kitti_k1["frame_ids"]
    pos = kitti_io.camera_positions(kitti_k1["est"])
    S, rmse, mx = orc.umeyama(pos, gt[ids])
    total = np.linalg.norm(np.diff(gt, axis=0), axis=1).sum()
    assert 3700 < total < 3750                      # KITTI-00 path length in metres
    c = np.cbrt(np.linalg.det(S[:3, :3]))
    assert abs(c - 2.6459) < 1e-3 and abs(rmse - 130.2362) < 1e-3 and abs(mx - 264.0688) < 1e-3    # [DERIVED] known answer
    # aligning the ground truth onto itself is exact
    S0, r0, m0 = orc.umeyama(gt[ids], gt[ids])
    assert np.abs(S0 - np.eye(4)).max() <= 1e-9 and r0 <= 1e-9


@pytest.mark.gpu
def test_device_alignment_matches_oracle(kitti_k1):
    import sim3opt_b200 as s3
    from oracle import oracle as orc, kitti_io
    rng = np.random.default_rng(5)
    for n in (3, 50, 771, 200000):
        q = rng.normal(0, 30, (n, 3)) + np.array([100.0, -50.0, 7.0])
        c, R, t = _random_similarity(rng)
        train = c * q @ R.T + t + rng.normal(0, 0.05, (n, 3))
        S_g, rmse_g, mx_g = s3.align_similarity(q, train)
        S_c, rmse_c, mx_c = orc.umeyama(q, train)
        assert np.abs(S_g - S_c).max() <= 1e-9 * max(1.0, np.abs(S_c).max())
        assert abs(rmse_g - rmse_c) <= 1e-9 * max(rmse_c, 1e-3) and abs(mx_g - mx_c) <= 1e-9 * max(mx_c, 1e-3)
        S_g2, rmse_g2, _ = s3.align_similarity(q, train)
        assert np.array_equal(S_g, S_g2) and rmse_g == rmse_g2          # deterministic reductions
        So_g, ro_g, mo_g = s3.align_similarity(q, train, only_scale=True)
        So_c, ro_c, mo_c = orc.umeyama(q, train, only_scale=True)
        assert np.abs(So_g - So_c).max() <= 1e-12 * np.abs(So_c).max() and abs(ro_g - ro_c) <= 1e-9 * ro_c
    gt = kitti_io.load_kitti_gt_positions(os.path.join(KITTI_DIR, "00.txt"))[kitti_k1["frame_ids"]]
    pos = kitti_io.camera_positions(kitti_k1["est"])
    S_g, rmse_g, mx_g = s3.align_similarity(pos, gt)
    S_c, rmse_c, mx_c = orc.umeyama(pos, gt)
    assert np.abs(S_g - S_c).max() <= 1e-9 * np.abs(S_c).max() and abs(rmse_g - rmse_c) <= 1e-9 * rmse_c
