"""Oracle-side restatement of the reference's KITTI-00 loaders and graph builders.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Follows
  LoadKFIndices        kitti_surf.cpp:232-254
  LoadKFPoses          kitti_surf.cpp:255-292   (2 header lines; T_w2c = (roteu2ro(rpy), xyz)^-1)
  LoadLoopConstraints  kitti_surf.cpp:145-205   (5 header lines; 4-line groups, ids from line 1,
                                                 Sim3 from line 4: count, scale, rpy, t)
  graph construction   kitti_surf.cpp:575-670   (vertices in kf order, vertex 0 fixed, loop edges
                                                 first then odometry edges, v0=i, v1=j, Omega=I)
  scale+trans graph    kitti_surf.cpp:767-886, toScaleTrans :533-539
"""
import os

import numpy as np

from . import oracle as orc


def load_kf_indices(cc_file):
    with open(cc_file) as f:
        return [int(tok) for tok in f.read().split()]


def load_kf_poses(pose_file, frame_ids):
    """Returns T_w2c per keyframe as (R 3x3, t 3) lists."""
    Rs, ts = [], []
    it = 0
    with open(pose_file) as f:
        lines = f.read().splitlines()[2:]
    for line in lines:
        if not line.strip():
            continue
        parts = [p.strip() for p in line.split(",")]
        fid = int(parts[0])
        if it < len(frame_ids) and fid == frame_ids[it]:
            rpy = np.array([float(parts[2]), float(parts[3]), float(parts[4])])
            xyz = np.array([float(parts[5]), float(parts[6]), float(parts[7])])
            Rc2w = orc.roteu2ro(rpy)
            Rw2c = Rc2w.T
            Rs.append(Rw2c)
            ts.append(-Rw2c @ xyz)
            it += 1
    assert it == len(frame_ids)
    return Rs, ts


def load_loop_constraints(loop_file):
    """Returns [(frame_id1, frame_id2, sim3 state[8])]."""
    with open(loop_file) as f:
        lines = f.read().splitlines()[5:]
    out = []
    k = 0
    while k + 3 < len(lines):
        first = lines[k].split()
        if len(first) < 8:
            break
        id1, id2 = int(first[0]), int(first[1])
        fourth = lines[k + 3].split()
        sf2s = float(fourth[1])
        rpy = np.array([float(x) for x in fourth[2:5]])
        t = np.array([float(x) for x in fourth[5:8]])
        assert rpy[0] != 0 and rpy[1] != 0 and rpy[2] != 0
        R = orc.roteu2ro(rpy)
        S = np.concatenate([orc.rot_to_quat(R), t, [sf2s]])
        out.append((id1, id2, S))
        k += 4
    return out


def build_kitti_sim3_graph(data_dir, use_one_constraint=True):
    """The graph of testDirectSim3Optimization (kitti_surf.cpp:560-670).

    Returns dict(est[n,8], fixed[n], v0[e], v1[e], meas[e,8], frame_ids)."""
    frame_ids = load_kf_indices(os.path.join(data_dir, "cc.txt"))
    Rs, ts = load_kf_poses(os.path.join(data_dir, "framePoses.txt"), frame_ids)
    loops = load_loop_constraints(os.path.join(data_dir, "loopConstraints.txt"))
    if use_one_constraint:
        loops = loops[:1]
    f2k = {fid: k for k, fid in enumerate(frame_ids)}
    n = len(frame_ids)
    est = np.zeros((n, 8))
    for k in range(n):
        est[k, :4] = orc.rot_to_quat(Rs[k])
        est[k, 4:7] = ts[k]
        est[k, 7] = 1.0
    fixed = np.zeros(n, np.uint8)
    fixed[0] = 1
    v0, v1, meas = [], [], []
    for id1, id2, S in loops:
        v0.append(f2k[id1])
        v1.append(f2k[id2])
        meas.append(S)
    for i in range(1, n):
        j = i - 1
        Swi = orc.sim3_inv(est[i])
        Sji = orc.sim3_mul(est[j], Swi)
        v0.append(i)
        v1.append(j)
        meas.append(Sji)
    return dict(est=est, fixed=fixed, v0=np.array(v0, np.int32), v1=np.array(v1, np.int32),
                meas=np.array(meas), frame_ids=np.array(frame_ids, np.int32))


def to_scale_trans_graph(g):
    """4-DoF twin of a Sim3 graph (kitti_surf.cpp:780-793, :834-839, :872-877)."""
    est = np.concatenate([g["est"][:, 7:8], g["est"][:, 4:7]], axis=1)
    aux = g["est"][:, :4].copy()
    meas = np.concatenate([g["meas"][:, 7:8], g["meas"][:, 4:7]], axis=1)
    return dict(est=est, aux=aux, fixed=g["fixed"], v0=g["v0"], v1=g["v1"], meas=meas)


def load_kitti_gt_positions(pose_file):
    """readKITTIPoseFile (kitti_surf.cpp:1166-1190): every line holds the 12 numbers of the 3x4 T_c2w in
    row-major order; returns the camera positions [n,3] (column 3)."""
    M = np.loadtxt(pose_file).reshape(-1, 3, 4)
    return M[:, :, 3].copy()


def camera_positions(est):
    """Camera centres in the world of Sim3 estimates S_iw = (q, t, s): -R^T t / s (what the result writer
    stores as t_i_in_w, kitti_surf.cpp:678-703)."""
    est = np.asarray(est, float).reshape(-1, 8)
    out = np.zeros((len(est), 3))
    for k, S in enumerate(est):
        out[k] = orc.sim3_inv(S)[4:7]
    return out
