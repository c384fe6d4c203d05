"""CPU tests of the bundle-adjustment oracle (oracle/ba.c, restating bal_example.cpp:44-243 + g2o's
EdgeProjectXYZ2UV / VertexSE3Expmap / Schur solve).  The reference pins nothing here (no tests, and
its own call sites are broken), so the oracle is pinned by self-consistency."""
import numpy as np
import pytest

from oracle import oracle as orc
from sim3opt_b200 import synth

HUBER = 1


@pytest.fixture(scope="module")
def ba_small():
    return synth.ba_loop(24, 300, 5, seed=11)


def make(g, robust=True, cam_fixed=None, pt_fixed=None):
    p = orc.BAProblem()
    p.set(g["cams"], g["points"], g["obs_cam"], g["obs_pt"], g["uv"], g["focal"], g["cx"], g["cy"],
          cam_fixed=cam_fixed, pt_fixed=pt_fixed)
    if robust:
        p.set_robust(HUBER, 2.5)
    return p


def project(g, cam, pt):
    R = orc.quat_to_rot(cam[:4])
    x = R @ pt + cam[4:7]
    return np.array([g["focal"] * x[0] / x[2] + g["cx"], g["focal"] * x[1] / x[2] + g["cy"]])


def test_error_is_observation_minus_projection(ba_small):
    g = ba_small
    p = make(g)
    p.build_structure()
    e = p.edge_errors()
    for k in (0, 7, len(e) - 1):
        ref = g["uv"][k] - project(g, g["cams"][g["obs_cam"][k]], g["points"][g["obs_pt"][k]])
        assert np.abs(e[k] - ref).max() <= 1e-9


def test_analytic_jacobians_match_central_differences(ba_small):
    """g2o's analytic linearizeOplus against differences through the vertices' own oplus."""
    g = ba_small
    p = make(g)
    p.build_structure()
    h = 1e-6
    for k in (0, 13, 101):
        Jp, Jc = p.edge_jacobians(k)
        cam, pt, z = g["cams"][g["obs_cam"][k]], g["points"][g["obs_pt"][k]], g["uv"][k]
        for c in range(3):
            d = np.zeros(3); d[c] = h
            num = ((z - project(g, cam, pt + d)) - (z - project(g, cam, pt - d))) / (2 * h)
            assert np.abs(num - Jp[:, c]).max() <= 1e-5 * max(1.0, np.abs(Jp).max())
        for c in range(6):
            d = np.zeros(6); d[c] = h
            cp = orc.se3_mul(orc.se3_exp(d), cam)
            cm = orc.se3_mul(orc.se3_exp(-d), cam)
            num = ((z - project(g, cp, pt)) - (z - project(g, cm, pt))) / (2 * h)
            assert np.abs(num - Jc[:, c]).max() <= 1e-5 * max(1.0, np.abs(Jc).max())


def dense_system(p, g, lam):
    Hpp, Hll, Hpl, b = p.linearize()
    ncf, npf = p.ncf, p.npf
    n = 6 * ncf + 3 * npf
    A = np.zeros((n, n))
    for c in range(ncf):
        A[6 * c:6 * c + 6, 6 * c:6 * c + 6] = Hpp[c]
    for l in range(npf):
        o = 6 * ncf + 3 * l
        A[o:o + 3, o:o + 3] = Hll[l]
    for k in range(p.no):          # nothing fixed in these tests: Hessian index == vertex index
        c, l = g["obs_cam"][k], g["obs_pt"][k]
        o = 6 * ncf + 3 * l
        A[6 * c:6 * c + 6, o:o + 3] += Hpl[k]
        A[o:o + 3, 6 * c:6 * c + 6] += Hpl[k].T
    return A + lam * np.eye(n), b


def test_schur_solve_equals_full_dense_solve(ba_small):
    g = ba_small
    p = make(g)
    colptr, rowidx = p.build_structure()
    assert p.ncf == 24 and p.npf == 300
    # H_schur pattern: every camera pair sharing a point, upper triangle, rows ascending per column
    pairs = set()
    for l in range(300):
        cams = sorted(set(g["obs_cam"][g["obs_pt"] == l]))
        for a in cams:
            for b in cams:
                if a <= b:
                    pairs.add((a, b))
    got = {(rowidx[k], c) for c in range(24) for k in range(colptr[c], colptr[c + 1])}
    assert got == pairs
    for c in range(24):
        assert np.all(np.diff(rowidx[colptr[c]:colptr[c + 1]]) > 0)
    lam = 1e-5 * 1.0
    A, b = dense_system(p, g, lam)
    assert np.abs(A - A.T).max() <= 1e-9 * np.abs(A).max()
    x_ref = np.linalg.solve(A, b)
    rc, x = p.solve(lam)
    assert rc == 0
    assert np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)
    assert np.abs(x - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    # explicit Schur complement against the dense formula
    rc, S, bs = p.schur(lam)
    App, Apl, All = A[:144, :144], A[:144, 144:], A[144:, 144:]
    Sd = App - Apl @ np.linalg.solve(All, Apl.T)
    for c in range(24):
        for k in range(colptr[c], colptr[c + 1]):
            r = rowidx[k]
            assert np.abs(S[k] - Sd[6 * r:6 * r + 6, 6 * c:6 * c + 6]).max() <= 1e-8 * np.abs(Sd).max()
    assert np.abs(bs - (b[:144] - Apl @ np.linalg.solve(All, b[144:]))).max() <= 1e-8 * np.abs(b).max()


def test_lm_converges_to_noise_floor(ba_small):
    g = ba_small
    p = make(g)
    p.build_structure()
    chi0 = p.chi2()
    n, chi2, lam, hist = p.optimize(15, 1e-6)
    dof = 2 * len(g["uv"]) - 6 * 24 - 3 * 300
    assert chi2 < 0.05 * chi0
    assert 0.7 * dof <= chi2 <= 1.3 * dof            # sigma = 1 px noise, Huber barely active at the optimum
    assert np.all(np.diff(hist[:, 0]) <= 0)
    # gauge is free (bal_example.cpp:110-117 fixes no camera): compare reprojection, not poses
    assert np.abs(p.edge_errors()).max() < 6.0


def test_fixed_vertices_are_excluded(ba_small):
    g = ba_small
    cf = np.zeros(24, np.uint8); cf[0] = 1
    pf = np.zeros(300, np.uint8); pf[:5] = 1
    p = make(g, cam_fixed=cf, pt_fixed=pf)
    p.build_structure()
    assert p.ncf == 23 and p.npf == 295
    Hpp, Hll, Hpl, b = p.linearize()
    touches_fixed = (g["obs_cam"] == 0) | (g["obs_pt"] < 5)
    assert np.all(Hpl[touches_fixed] == 0) and np.any(Hpl[~touches_fixed] != 0)
    cams0, pts0 = p.cameras().copy(), p.points().copy()
    p.optimize(3)
    assert np.array_equal(p.cameras()[0], cams0[0]) and np.array_equal(p.points()[:5], pts0[:5])
    assert not np.array_equal(p.cameras()[1], cams0[1])


def test_ba_lm_fixed_point_against_scipy_least_squares():
    """Independent pin of the BA oracle's LM (Jacobians, Schur complement, LDL^T, LM rules): scipy's trust-region
    least squares with finite-difference Jacobians, over the same residuals on a chart around the initial guess
    (cameras: exp(delta) * cam, points: additive), reaches the same stationary point."""
    from scipy.optimize import least_squares
    g = synth.ba_loop(6, 40, 4, seed=3)
    nc, npt = len(g["cams"]), len(g["points"])
    cam_fixed = np.zeros(nc, np.uint8)
    cam_fixed[:2] = 1                                   # gauge: two cameras pin the similarity
    p = make(g, robust=False, cam_fixed=cam_fixed)
    p.build_structure()
    n, chi2, _, _ = p.optimize(60)
    cams_lm, pts_lm = p.cameras(), p.points()
    free_c = np.flatnonzero(cam_fixed == 0)
    q = make(g, robust=False, cam_fixed=cam_fixed)

    def state(x):
        cams = g["cams"].copy()
        for k, c in enumerate(free_c):
            cams[c] = orc.se3_mul(orc.se3_exp(x[6 * k:6 * k + 6]), g["cams"][c])
        pts = g["points"] + x[6 * len(free_c):].reshape(npt, 3)
        return cams, pts

    def residuals(x):
        cams, pts = state(x)
        q.set(cams, pts, g["obs_cam"], g["obs_pt"], g["uv"], g["focal"], g["cx"], g["cy"], cam_fixed=cam_fixed)
        q.build_structure()
        return q.edge_errors().ravel()

    sol = least_squares(residuals, np.zeros(6 * len(free_c) + 3 * npt), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-10,
                        x_scale="jac", max_nfev=300)
    chi2_sp = float((sol.fun ** 2).sum())
    cams_sp, pts_sp = state(sol.x)
    assert abs(chi2 - chi2_sp) <= 1e-8 * chi2_sp, (chi2, chi2_sp, n)
    sgn = np.sign((cams_lm[:, :4] * cams_sp[:, :4]).sum(1))[:, None]
    assert np.abs(cams_lm[:, :4] - sgn * cams_sp[:, :4]).max() <= 1e-5
    assert np.abs(cams_lm[:, 4:] - cams_sp[:, 4:]).max() <= 1e-4
    assert np.abs(pts_lm - pts_sp).max() <= 1e-3
