"""Synthetic Sim3 pose graphs.

Host-side input generation only (numpy); nothing here is on the optimisation path.

sphere(n_laps, poses_per_lap): ground-truth camera centres on a radius-100 m sphere spiral, camera
z-axis along the direction of travel, ground-truth scale s_k = exp(0.3 sin(2 pi k / N)).  Edges per
vertex k: (k,k+1), (k,k+2), (k,k+L), (k,k+L+1), (k,k+2L) with L = poses_per_lap where in range
(about 5N edges).  Edge convention is the reference's (kitti_surf.cpp:624-670): vertex(0) = i,
vertex(1) = j, measurement C ~ S_j S_i^-1, so that e = log(C S_i S_j^-1) is pure noise at the ground
truth.  Measurement noise n ~ N(0, diag(0.01, 0.05, 0.01)^2) on (omega, upsilon, sigma), information
Omega = diag(1/sigma^2) stored as a full 7x7; initial guess = exp(N(0, (0.05, 0.5, 0.05)^2)) * GT;
vertex 0 fixed.  Seeded with numpy's PCG64 so the oracle and the GPU path see identical arrays.
"""
import numpy as np
from scipy.spatial.transform import Rotation


def _skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]; S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2];  S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]; S[..., 2, 1] = v[..., 0]
    return S


def sim3_exp(v):
    """Vectorised exact Sim3 exponential; v[...,7] = [omega, upsilon, sigma] -> (Rotation, t, s)."""
    v = np.asarray(v, float).reshape(-1, 7)
    om, up, sg = v[:, :3], v[:, 3:6], v[:, 6]
    th = np.linalg.norm(om, axis=1)
    s = np.exp(sg)
    small_t = th < 1e-6
    small_s = np.abs(sg) < 1e-6
    th_ = np.where(small_t, 1.0, th)
    sg_ = np.where(small_s, 1.0, sg)
    C = np.where(small_s, 1.0 + 0.5 * sg, (s - 1) / sg_)
    a, b = s * np.sin(th_), s * np.cos(th_)
    c = th_ ** 2 + sg ** 2
    A_gen = (a * sg + (1 - b) * th_) / (th_ * np.where(c == 0, 1.0, c))
    B_gen = (C - ((b - 1) * sg + a * th_) / np.where(c == 0, 1.0, c)) / th_ ** 2
    A_smallt = np.where(small_s, 0.5, ((sg - 1) * s + 1) / sg_ ** 2)
    B_smallt = np.where(small_s, 1.0 / 6.0, ((0.5 * sg ** 2 - sg + 1) * s - 1) / sg_ ** 3)
    A = np.where(small_t, A_smallt, A_gen)
    B = np.where(small_t, B_smallt, B_gen)
    Om = _skew(om)
    W = A[:, None, None] * Om + B[:, None, None] * (Om @ Om) + C[:, None, None] * np.eye(3)
    t = np.einsum("nij,nj->ni", W, up)
    return Rotation.from_rotvec(om), t, s


def _mul(Ra, ta, sa, Rb, tb, sb):
    return Ra * Rb, sa[:, None] * Ra.apply(tb) + ta, sa * sb


def _inv(R, t, s):
    Ri = R.inv()
    return Ri, -Ri.apply(t) / s[:, None], 1.0 / s


def _pack(R, t, s):
    q = R.as_quat()  # x y z w
    q = np.where(q[:, 3:4] < 0, -q, q)
    return np.concatenate([q, t, s[:, None]], axis=1)


def sphere(n_laps=100, poses_per_lap=1000, seed=42, radius=100.0,
           meas_sigma=(0.01, 0.05, 0.01), init_sigma=(0.05, 0.5, 0.05), full_info=True):
    """Returns dict(est, gt, fixed, v0, v1, meas, info) -- see the module docstring."""
    N, L = n_laps * poses_per_lap, poses_per_lap
    rng = np.random.default_rng(seed)
    k = np.arange(N)
    az = 2 * np.pi * k / L
    pol = np.pi * (0.1 + 0.8 * (k + 0.5) / N)
    c = radius * np.stack([np.sin(pol) * np.cos(az), np.sin(pol) * np.sin(az), np.cos(pol)], axis=1)
    # tangent of the spiral (dominant azimuthal motion) -> camera z; radial -> camera x
    dpol = np.pi * 0.8 / N
    daz = 2 * np.pi / L
    tang = radius * np.stack([
        np.cos(pol) * np.cos(az) * dpol - np.sin(pol) * np.sin(az) * daz,
        np.cos(pol) * np.sin(az) * dpol + np.sin(pol) * np.cos(az) * daz,
        -np.sin(pol) * dpol], axis=1)
    z = tang / np.linalg.norm(tang, axis=1, keepdims=True)
    x = c / np.linalg.norm(c, axis=1, keepdims=True)
    x = x - (x * z).sum(1, keepdims=True) * z
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    Rc2w = np.stack([x, y, z], axis=2)
    Rw2c = Rotation.from_matrix(np.transpose(Rc2w, (0, 2, 1)))
    s_gt = np.exp(0.3 * np.sin(2 * np.pi * k / N))
    t_gt = -s_gt[:, None] * Rw2c.apply(c)

    offs = [1, 2, L, L + 1, 2 * L]
    i_idx = np.concatenate([k[: N - o] for o in offs if o < N])
    j_idx = np.concatenate([k[: N - o] + o for o in offs if o < N])
    order = np.lexsort((j_idx, i_idx))
    i_idx, j_idx = i_idx[order], j_idx[order]
    E = len(i_idx)

    Ri, ti, si = Rw2c[i_idx], t_gt[i_idx], s_gt[i_idx]
    Rj, tj, sj = Rw2c[j_idx], t_gt[j_idx], s_gt[j_idx]
    Rii, tii, sii = _inv(Ri, ti, si)
    Rji, tji, sji = _mul(Rj, tj, sj, Rii, tii, sii)
    sig = np.array([meas_sigma[0]] * 3 + [meas_sigma[1]] * 3 + [meas_sigma[2]])
    noise = rng.standard_normal((E, 7)) * sig
    Rn, tn, sn = sim3_exp(noise)
    Rm, tm, sm = _mul(Rn, tn, sn, Rji, tji, sji)
    meas = _pack(Rm, tm, sm)

    isig = np.array([init_sigma[0]] * 3 + [init_sigma[1]] * 3 + [init_sigma[2]])
    pert = rng.standard_normal((N, 7)) * isig
    pert[0] = 0
    Rp, tp, sp = sim3_exp(pert)
    Re, te, se = _mul(Rp, tp, sp, Rw2c, t_gt, s_gt)
    est = _pack(Re, te, se)
    gt = _pack(Rw2c, t_gt, s_gt)
    fixed = np.zeros(N, np.uint8)
    fixed[0] = 1
    info = None
    if full_info:
        info = np.zeros((E, 7, 7))
        info[:, np.arange(7), np.arange(7)] = 1.0 / sig ** 2
    return dict(est=est, gt=gt, fixed=fixed, v0=i_idx.astype(np.int32), v1=j_idx.astype(np.int32),
                meas=meas, info=info)


def manhattan3d(n_poses=10000, seed=42, box=None, loop_radius=2.0, max_loops_per_pose=3, min_gap=10,
                meas_sigma=(0.01, 0.05, 0.01), init_sigma=(0.05, 0.3, 0.05)):
    """Manhattan-3D variant of config 3: a unit-step random walk on a 3-D grid confined to
    a box (reflecting walls), camera z-axis along the last step, ground-truth scale exp(0.2 sin(2 pi k/N));
    odometry edges (k, k+1) plus loop edges (i, j), j >= i + min_gap, between poses at most `loop_radius`
    cells apart (at most `max_loops_per_pose` per pose, nearest in time first).  Same edge convention,
    noise model and return dict as sphere(); the graph is irregular (ragged block rows)."""
    N = int(n_poses)
    rng = np.random.default_rng(seed)
    if box is None:
        box = max(4, int(round((N / 4.0) ** (1.0 / 3.0))))
    dirs = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]])
    pos = np.zeros((N, 3), np.int64)
    step = np.zeros((N, 3), np.int64)
    step[0] = dirs[0]
    choice = rng.integers(0, 6, N)
    for k in range(1, N):
        d = dirs[choice[k]]
        q = pos[k - 1] + d
        bad = (q < 0) | (q >= box)
        d = np.where(bad, -d, d)
        pos[k] = pos[k - 1] + d
        step[k] = d
    # camera frame: z along the step, x any perpendicular axis
    z = step.astype(float)
    x = np.where(np.abs(z[:, [0]]) > 0.5, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    y = np.cross(z, x)
    Rc2w = np.stack([x, y, z], axis=2)
    Rw2c = Rotation.from_matrix(np.transpose(Rc2w, (0, 2, 1)))
    k = np.arange(N)
    s_gt = np.exp(0.2 * np.sin(2 * np.pi * k / N))
    c = pos.astype(float)
    t_gt = -s_gt[:, None] * Rw2c.apply(c)
    # loop edges through a cell hash
    cells = {}
    for idx in range(N):
        cells.setdefault(tuple(pos[idx]), []).append(idx)
    r = int(np.ceil(loop_radius))
    offs = [(a, b, cc) for a in range(-r, r + 1) for b in range(-r, r + 1) for cc in range(-r, r + 1)
            if a * a + b * b + cc * cc <= loop_radius * loop_radius]
    li, lj = [], []
    for j in range(N):
        cand = []
        pj = pos[j]
        for o in offs:
            for i in cells.get((pj[0] + o[0], pj[1] + o[1], pj[2] + o[2]), ()):
                if i <= j - min_gap:
                    cand.append(i)
        cand.sort(reverse=True)
        for i in cand[:max_loops_per_pose]:
            li.append(i); lj.append(j)
    i_idx = np.concatenate([k[:-1], np.array(li, np.int64)])
    j_idx = np.concatenate([k[1:], np.array(lj, np.int64)])
    order = np.lexsort((j_idx, i_idx))
    i_idx, j_idx = i_idx[order], j_idx[order]
    E = len(i_idx)
    Ri, ti, si = Rw2c[i_idx], t_gt[i_idx], s_gt[i_idx]
    Rj, tj, sj = Rw2c[j_idx], t_gt[j_idx], s_gt[j_idx]
    Rii, tii, sii = _inv(Ri, ti, si)
    Rji, tji, sji = _mul(Rj, tj, sj, Rii, tii, sii)
    sig = np.array([meas_sigma[0]] * 3 + [meas_sigma[1]] * 3 + [meas_sigma[2]])
    Rn, tn, sn = sim3_exp(rng.standard_normal((E, 7)) * sig)
    Rm, tm, sm = _mul(Rn, tn, sn, Rji, tji, sji)
    meas = _pack(Rm, tm, sm)
    isig = np.array([init_sigma[0]] * 3 + [init_sigma[1]] * 3 + [init_sigma[2]])
    pert = rng.standard_normal((N, 7)) * isig
    pert[0] = 0
    Rp, tp, sp = sim3_exp(pert)
    Re, te, se = _mul(Rp, tp, sp, Rw2c, t_gt, s_gt)
    fixed = np.zeros(N, np.uint8)
    fixed[0] = 1
    info = np.zeros((E, 7, 7))
    info[:, np.arange(7), np.arange(7)] = 1.0 / sig ** 2
    return dict(est=_pack(Re, te, se), gt=_pack(Rw2c, t_gt, s_gt), fixed=fixed, v0=i_idx.astype(np.int32),
                v1=j_idx.astype(np.int32), meas=meas, info=info)


def ba_loop(n_cams=1000, n_points=500000, obs_per_point=10, seed=42, loop_length=200.0, focal=718.856,
            cx=607.1928, cy=185.2157, pixel_sigma=1.0, cam_pert=(0.01, 0.1), point_pert=0.2):
    """Synthetic bundle adjustment of SURVEY.md 8(d) config 5 (bal_example.cpp:87-88 intrinsics).

    Cameras ride a horizontal circle of circumference loop_length looking along the direction of
    travel; a point is created in front of a randomly chosen "home" camera (depth 5-50 m, lateral
    +-0.35*depth, vertical +-0.12*depth, i.e. inside the image) and is observed by that camera and
    the obs_per_point-1 cameras behind it on the path that still see it in front (z > 1 m).
    Returns dict(cams[n,7] SE3Quat world->camera [q xyzw, t], points[m,3], cams_gt, points_gt,
    obs_cam, obs_pt, uv, focal, cx, cy).  Observations are sorted by (point, camera) -- the order the
    reference's file format lists them in is arbitrary (bal_example.cpp:132-166)."""
    rng = np.random.default_rng(seed)
    C = n_cams
    r = loop_length / (2 * np.pi)
    ang = 2 * np.pi * np.arange(C) / C
    centre = np.stack([r * np.cos(ang), r * np.sin(ang), np.zeros(C)], axis=1)
    zc = np.stack([-np.sin(ang), np.cos(ang), np.zeros(C)], axis=1)           # direction of travel
    yc = np.tile(np.array([0.0, 0.0, -1.0]), (C, 1))                          # image y points down
    xc = np.cross(yc, zc)
    Rc2w = np.stack([xc, yc, zc], axis=2)
    Rw2c = Rotation.from_matrix(np.transpose(Rc2w, (0, 2, 1)))
    t_gt = -Rw2c.apply(centre)
    home = rng.integers(0, C, n_points)
    depth = rng.uniform(5.0, 50.0, n_points)
    lat = rng.uniform(-0.35, 0.35, n_points) * depth
    ver = rng.uniform(-0.12, 0.12, n_points) * depth
    Xc = np.stack([lat, ver, depth], axis=1)
    pts_gt = np.einsum("nij,nj->ni", Rc2w[home], Xc) + centre[home]
    obs_cam, obs_pt, uv = [], [], []
    for back in range(obs_per_point):
        cam = (home - back) % C
        Xk = Rw2c[cam].apply(pts_gt) + t_gt[cam]
        ok = Xk[:, 2] > 1.0
        u = focal * Xk[:, 0] / Xk[:, 2] + cx
        v = focal * Xk[:, 1] / Xk[:, 2] + cy
        idx = np.nonzero(ok)[0]
        obs_cam.append(cam[idx]); obs_pt.append(idx); uv.append(np.stack([u[idx], v[idx]], axis=1))
    obs_cam = np.concatenate(obs_cam); obs_pt = np.concatenate(obs_pt); uv = np.concatenate(uv)
    order = np.lexsort((obs_cam, obs_pt))
    obs_cam, obs_pt, uv = obs_cam[order], obs_pt[order], uv[order]
    uv = uv + rng.standard_normal(uv.shape) * pixel_sigma
    # perturbed initial guess: cameras exp([dw, dt]) * T_gt, points + noise
    dw = rng.standard_normal((C, 3)) * cam_pert[0]
    dt = rng.standard_normal((C, 3)) * cam_pert[1]
    Rp = Rotation.from_rotvec(dw)
    R_est = Rp * Rw2c
    t_est = Rp.apply(t_gt) + dt
    q = R_est.as_quat(); q = np.where(q[:, 3:4] < 0, -q, q)
    qg = Rw2c.as_quat(); qg = np.where(qg[:, 3:4] < 0, -qg, qg)
    cams = np.concatenate([q, t_est], axis=1)
    cams_gt = np.concatenate([qg, t_gt], axis=1)
    pts = pts_gt + rng.standard_normal(pts_gt.shape) * point_pert
    return dict(cams=cams, points=pts, cams_gt=cams_gt, points_gt=pts_gt, obs_cam=obs_cam.astype(np.int32),
                obs_pt=obs_pt.astype(np.int32), uv=uv, focal=focal, cx=cx, cy=cy)


def write_bal(path, g):
    """Writes a ba_loop() problem in the text layout ba_demo reads (bal_example.cpp:104-194): header,
    observations "cam point u v", 9 numbers per camera (angle-axis, t, f, k1, k2), points; %.17g."""
    rot = Rotation.from_quat(g["cams"][:, :4]).as_rotvec()
    with open(path, "w") as f:
        f.write(f"{len(g['cams'])} {len(g['points'])} {len(g['uv'])}\n")
        for c, p, (u, v) in zip(g["obs_cam"], g["obs_pt"], g["uv"]):
            f.write(f"{c} {p} {u:.17g} {v:.17g}\n")
        for aa, cam in zip(rot, g["cams"]):
            for x in list(aa) + list(cam[4:7]) + [g["focal"], 0.0, 0.0]:
                f.write(f"{x:.17g}\n")
        for pt in g["points"]:
            for x in pt:
                f.write(f"{x:.17g}\n")
