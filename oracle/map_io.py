"""Checker-side restatement of the reference's map re-projection and BAL export (SURVEY.md 8f row N2;
drawPTAMPoints.cpp:33-84 LoadComboKeyFrame, :285-456 figureKITTIBA with lineFormat 1, :218-283 SaveBALFile).
Test infrastructure only."""
import glob
import os
import struct

import numpy as np
from scipy.spatial.transform import Rotation

_REC = np.dtype([("id", "<u4"), ("p", "<f8", 3), ("cos", "<f8"), ("px", "<f8", 2)])


def load_combo_keyframe(path):
    b = open(path, "rb").read()
    kf_id, name_len = struct.unpack_from("<ii", b, 0)
    o = 8 + name_len + 16 + 40
    R = np.array(struct.unpack_from("<9d", b, o)).reshape(3, 3)
    t = np.array(struct.unpack_from("<3d", b, o + 72))
    o += 72 + 24 + 1
    n, = struct.unpack_from("<i", b, o)
    rec = np.frombuffer(b, dtype=_REC, count=n, offset=o + 4)
    return kf_id, R, t, rec


def reproject_map(keyframe_dir, trans_file):
    """Returns dict(R[c,3,3], t[c,3], points[p,3], obs_cam, obs_pt, uv, image_ids) with S221 = identity."""
    files = sorted(glob.glob(os.path.join(keyframe_dir, "KeyFrame*.bin")))
    oldR, oldt, ids, obs = [], [], [], []
    latest = {}
    for k, f in enumerate(files):
        kf_id, R, t, rec = load_combo_keyframe(f)
        ids.append(kf_id); oldR.append(R); oldt.append(t)
        for r in rec:
            latest[int(r["id"])] = np.array(r["p"])
            obs.append((k, int(r["id"]), float(r["px"][0]), float(r["px"][1])))
    order = sorted(latest)
    compact = {pid: n for n, pid in enumerate(order)}
    pts_old = np.array([latest[pid] for pid in order])
    rows = np.loadtxt(trans_file, comments="%").reshape(-1, 9)
    assert len(rows) == len(files)
    newR, newt, news = [], [], []
    for row in rows:
        s = row[1]
        Rw2c = Rotation.from_quat(row[5:9]).as_matrix().T
        newR.append(Rw2c); newt.append(-s * Rw2c @ row[2:5]); news.append(s)
    pts = pts_old.copy()
    for k, pid, u, v in obs:
        c = compact[pid]
        rel = oldR[k] @ pts_old[c] + oldt[k]
        pts[c] = newR[k].T @ (rel - newt[k]) / news[k]         # Sim3(R, t, s)^-1 applied to rel
    return dict(R=np.array(newR), t=np.array([newt[k] / news[k] for k in range(len(files))]), points=pts,
                obs_cam=np.array([o[0] for o in obs], np.int32), obs_pt=np.array([compact[o[1]] for o in obs], np.int32),
                uv=np.array([[o[2], o[3]] for o in obs]), image_ids=np.array(ids, np.int32))


def read_bal(path):
    with open(path) as f:
        nc, npts, nobs = (int(x) for x in f.readline().split())
        obs = np.array([f.readline().split() for _ in range(nobs)], float)
        rest = np.array(f.read().split(), float)
    cams = rest[:9 * nc].reshape(nc, 9)
    pts = rest[9 * nc:].reshape(npts, 3)
    return obs, cams, pts
