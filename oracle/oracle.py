"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by sim3opt_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

KIND_SIM3, KIND_SCALE_TRANS, KIND_SCALE, KIND_BA = 0, 1, 2, 3
JAC_NUMERIC, JAC_ANALYTIC = 0, 1
ROBUST_NONE, ROBUST_HUBER, ROBUST_PTAM_TUKEY, ROBUST_PTAM_CAUCHY, ROBUST_PTAM_HUBER, ROBUST_PTAM_LS = range(6)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("lie.c", "ldlt.c", "lm.c", "ba.c", "oracle.h", "ldlt.h", "Makefile")]
    if not force and os.path.exists(_LIB_PATH):
        t = os.path.getmtime(_LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in srcs):
            return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_vertices.argtypes = [C.c_void_p, C.c_int, _dp, _up, _dp]
    L.orc_set_edges.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp]
    L.orc_set_robust.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.orc_set_scale_model.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_jacobian_mode.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.orc_set_lm.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    L.orc_build_structure.argtypes = [C.c_void_p]
    L.orc_num_free.argtypes = [C.c_void_p]
    L.orc_num_blocks.argtypes = [C.c_void_p]
    L.orc_dim.argtypes = [C.c_void_p]
    L.orc_get_structure.argtypes = [C.c_void_p, _ip, _ip]
    L.orc_get_hessian_index.argtypes = [C.c_void_p, _ip]
    L.orc_chi2.restype = C.c_double
    L.orc_chi2.argtypes = [C.c_void_p]
    L.orc_edge_errors.argtypes = [C.c_void_p, _dp]
    L.orc_linearize.argtypes = [C.c_void_p]
    L.orc_get_H.argtypes = [C.c_void_p, _dp]
    L.orc_get_b.argtypes = [C.c_void_p, _dp]
    L.orc_max_diag.restype = C.c_double
    L.orc_max_diag.argtypes = [C.c_void_p]
    L.orc_solve.argtypes = [C.c_void_p, C.c_double, _dp]
    L.orc_update.argtypes = [C.c_void_p, _dp]
    L.orc_get_vertices.argtypes = [C.c_void_p, _dp]
    L.orc_optimize.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp]
    L.orc_get_timing.argtypes = [C.c_void_p, _dp]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_get_threads.restype = C.c_int
    for name, n_in in (("orc_sim3_exp", 1), ("orc_sim3_log", 1), ("orc_sim3_inv", 1), ("orc_sim3_adjoint", 1),
                       ("orc_sim3_ad", 1), ("orc_sim3_jl_inv", 1), ("orc_quat_to_rot", 1), ("orc_rot_to_quat", 1),
                       ("orc_roteu2ro", 1), ("orc_se3_exp", 1), ("orc_sim3_mul", 2), ("orc_se3_mul", 2)):
        getattr(L, name).argtypes = [_dp] * (n_in + 1)
    L.orc_sim3_edge_error.argtypes = [_dp] * 4
    L.orc_sim3_edge_jac_numeric.argtypes = [_dp, _dp, _dp, C.c_double, _dp, _dp]
    L.orc_sim3_edge_jac_analytic.argtypes = [_dp] * 5
    L.orc_robustify.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
    # bundle adjustment (ba.c)
    vp = C.c_void_p
    L.orc_ba_create.restype = vp
    L.orc_ba_destroy.argtypes = [vp]
    L.orc_ba_set.argtypes = [vp, C.c_int, _dp, _up, C.c_int, _dp, _up, C.c_int, _ip, _ip, _dp, _dp,
                             C.c_double, C.c_double, C.c_double]
    L.orc_ba_set_robust.argtypes = [vp, C.c_int, C.c_double]
    L.orc_ba_set_lm.argtypes = [vp, C.c_double, C.c_double, C.c_int]
    for name in ("orc_ba_build_structure", "orc_ba_num_free_cameras", "orc_ba_num_free_points", "orc_ba_num_blocks",
                 "orc_ba_linearize"):
        getattr(L, name).argtypes = [vp]
    L.orc_ba_get_structure.argtypes = [vp, _ip, _ip]
    L.orc_ba_chi2.restype = C.c_double
    L.orc_ba_chi2.argtypes = [vp]
    L.orc_ba_edge_errors.argtypes = [vp, _dp]
    L.orc_ba_edge_jacobians.argtypes = [vp, C.c_int, _dp, _dp]
    L.orc_ba_get_system.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.orc_ba_max_diag.restype = C.c_double
    L.orc_ba_max_diag.argtypes = [vp]
    L.orc_ba_schur.argtypes = [vp, C.c_double, _dp, _dp]
    L.orc_ba_solve.argtypes = [vp, C.c_double, _dp]
    L.orc_ba_update.argtypes = [vp, _dp]
    L.orc_ba_get_cameras.argtypes = [vp, _dp]
    L.orc_ba_get_points.argtypes = [vp, _dp]
    L.orc_ba_optimize.argtypes = [vp, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp]
    L.orc_ptam_find_sigma_squared.restype = C.c_double
    L.orc_ptam_find_sigma_squared.argtypes = [C.c_int, _dp, C.c_int]


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _call1(name, x, n_out):
    x = _f64(x)
    out = np.zeros(n_out)
    getattr(lib(), name)(_d(x), _d(out))
    return out


def sim3_exp(v): return _call1("orc_sim3_exp", v, 8)
def sim3_log(S): return _call1("orc_sim3_log", S, 7)
def sim3_inv(S): return _call1("orc_sim3_inv", S, 8)
def sim3_adjoint(S): return _call1("orc_sim3_adjoint", S, 49).reshape(7, 7)
def sim3_ad(e): return _call1("orc_sim3_ad", e, 49).reshape(7, 7)
def sim3_jl_inv(e): return _call1("orc_sim3_jl_inv", e, 49).reshape(7, 7)
def quat_to_rot(q): return _call1("orc_quat_to_rot", q, 9).reshape(3, 3)
def rot_to_quat(R): return _call1("orc_rot_to_quat", _f64(R).reshape(-1), 4)
def roteu2ro(eul): return _call1("orc_roteu2ro", eul, 9).reshape(3, 3)
def se3_exp(v): return _call1("orc_se3_exp", v, 7)


def sim3_mul(A, B):
    A, B = _f64(A), _f64(B)
    out = np.zeros(8)
    lib().orc_sim3_mul(_d(A), _d(B), _d(out))
    return out


def se3_mul(A, B):
    A, B = _f64(A), _f64(B)
    out = np.zeros(7)
    lib().orc_se3_mul(_d(A), _d(B), _d(out))
    return out


def sim3_edge_error(Cm, Si, Sj):
    Cm, Si, Sj = _f64(Cm), _f64(Si), _f64(Sj)
    e = np.zeros(7)
    lib().orc_sim3_edge_error(_d(Cm), _d(Si), _d(Sj), _d(e))
    return e


def sim3_edge_jac_numeric(Cm, Si, Sj, h=1e-9):
    Cm, Si, Sj = _f64(Cm), _f64(Si), _f64(Sj)
    Ji, Jj = np.zeros(49), np.zeros(49)
    lib().orc_sim3_edge_jac_numeric(_d(Cm), _d(Si), _d(Sj), h, _d(Ji), _d(Jj))
    return Ji.reshape(7, 7), Jj.reshape(7, 7)


def sim3_edge_jac_analytic(Cm, Si, Sj):
    Cm, Si, Sj = _f64(Cm), _f64(Si), _f64(Sj)
    Ji, Jj = np.zeros(49), np.zeros(49)
    lib().orc_sim3_edge_jac_analytic(_d(Cm), _d(Si), _d(Sj), _d(Ji), _d(Jj))
    return Ji.reshape(7, 7), Jj.reshape(7, 7)


MATH_REFERENCE, MATH_CORRECTED = 0, 1


def set_math_mode(mode):
    lib().orc_set_math_mode(int(mode))


def set_threads(n):
    """Worker threads of the per-edge loops (1 = the reference's single-threaded behaviour)."""
    lib().orc_set_threads(int(n))


def get_threads():
    return int(lib().orc_get_threads())


def robustify(kind, param, e2):
    rho = np.zeros(3)
    lib().orc_robustify(kind, param, e2, _d(rho))
    return rho


def ptam_find_sigma_squared(kind, err_sq):
    a = _f64(err_sq).copy()
    return lib().orc_ptam_find_sigma_squared(kind, _d(a), len(a))


class Problem:
    """Pose-graph problem of one kind (SIM3 d=7, SCALE_TRANS d=4, SCALE d=1)."""

    def __init__(self, kind=KIND_SIM3):
        self.L = lib()
        self.kind = kind
        self.h = self.L.orc_create(kind)
        if not self.h:
            raise ValueError("unsupported kind")
        self.d = self.L.orc_dim(self.h)
        self.est_dim = {KIND_SIM3: 8, KIND_SCALE_TRANS: 4, KIND_SCALE: 1}[kind]
        self.nv = self.ne = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_vertices(self, est, fixed=None, aux=None):
        est = _f64(est).reshape(-1, self.est_dim)
        self.nv = est.shape[0]
        fx = np.zeros(self.nv, np.uint8) if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        auxp = None
        if aux is not None:
            aux = _f64(aux).reshape(self.nv, 4)
            auxp = _d(aux)
        rc = self.L.orc_set_vertices(self.h, self.nv, _d(est), fx.ctypes.data_as(_up), auxp)
        if rc != 0:
            raise ValueError("orc_set_vertices failed")

    def set_edges(self, v0, v1, meas, info=None):
        v0 = np.ascontiguousarray(v0, np.int32)
        v1 = np.ascontiguousarray(v1, np.int32)
        meas = _f64(meas).reshape(-1, self.est_dim)
        self.ne = len(v0)
        infop = None
        if info is not None:
            info = _f64(info).reshape(self.ne, self.d, self.d)
            infop = _d(info)
        rc = self.L.orc_set_edges(self.h, self.ne, v0.ctypes.data_as(_ip), v1.ctypes.data_as(_ip), _d(meas), infop)
        if rc != 0:
            raise ValueError("orc_set_edges: bad vertex index")

    def set_robust(self, kind, param): self.L.orc_set_robust(self.h, kind, float(param))
    def set_scale_model(self, model): self.L.orc_set_scale_model(self.h, int(model))
    def set_jacobian_mode(self, mode, h=0.0): self.L.orc_set_jacobian_mode(self.h, mode, float(h))
    def set_lm(self, tau=0.0, lambda_init=0.0, max_trials=0): self.L.orc_set_lm(self.h, tau, lambda_init, max_trials)

    def build_structure(self):
        nb = self.L.orc_build_structure(self.h)
        nf = self.L.orc_num_free(self.h)
        colptr = np.zeros(nf + 1, np.int32)
        rowidx = np.zeros(max(nb, 1), np.int32)
        self.L.orc_get_structure(self.h, colptr.ctypes.data_as(_ip), rowidx.ctypes.data_as(_ip))
        return colptr, rowidx[:nb]

    @property
    def num_free(self): return self.L.orc_num_free(self.h)
    @property
    def num_blocks(self): return self.L.orc_num_blocks(self.h)

    def hessian_index(self):
        h = np.zeros(self.nv, np.int32)
        self.L.orc_get_hessian_index(self.h, h.ctypes.data_as(_ip))
        return h

    def chi2(self): return self.L.orc_chi2(self.h)

    def edge_errors(self):
        e = np.zeros((self.ne, self.d))
        self.L.orc_edge_errors(self.h, _d(e))
        return e

    def linearize(self):
        self.L.orc_linearize(self.h)
        H = np.zeros((self.num_blocks, self.d, self.d))
        b = np.zeros(self.num_free * self.d)
        self.L.orc_get_H(self.h, _d(H))
        self.L.orc_get_b(self.h, _d(b))
        return H, b

    def max_diag(self): return self.L.orc_max_diag(self.h)

    def solve(self, lam):
        x = np.zeros(self.num_free * self.d)
        rc = self.L.orc_solve(self.h, float(lam), _d(x))
        return rc, x

    def update(self, x):
        x = _f64(x)
        self.L.orc_update(self.h, _d(x))

    def vertices(self):
        est = np.zeros((self.nv, self.est_dim))
        self.L.orc_get_vertices(self.h, _d(est))
        return est

    def optimize(self, max_iter, stop_rel_gain=0.0):
        hist = np.zeros((max(max_iter, 1), 4))
        chi2 = C.c_double(0)
        lam = C.c_double(0)
        n = self.L.orc_optimize(self.h, max_iter, stop_rel_gain, _d(hist), max_iter, C.byref(chi2), C.byref(lam))
        return n, chi2.value, lam.value, hist[:max(n, 0)]

    def timing(self):
        t = np.zeros(4)
        self.L.orc_get_timing(self.h, _d(t))
        return t


class BAProblem:
    """Bundle-adjustment oracle (ba.c): cameras SE3Quat [q xyzw, t], points xyz, observations
    (camera, point, u, v); Huber etc. through set_robust; LM with Schur complement + sparse LDLT."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_ba_create())
        self.nc = self.np_ = self.no = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_ba_destroy(self.h)
            self.h = None

    def set(self, cams, pts, obs_cam, obs_pt, uv, focal, cx, cy, cam_fixed=None, pt_fixed=None, info=None):
        cams = _f64(cams).reshape(-1, 7)
        pts = _f64(pts).reshape(-1, 3)
        oc = np.ascontiguousarray(obs_cam, np.int32)
        op = np.ascontiguousarray(obs_pt, np.int32)
        uv = _f64(uv).reshape(-1, 2)
        self.nc, self.np_, self.no = len(cams), len(pts), len(oc)
        cf = np.zeros(self.nc, np.uint8) if cam_fixed is None else np.ascontiguousarray(cam_fixed, np.uint8)
        pf = np.zeros(self.np_, np.uint8) if pt_fixed is None else np.ascontiguousarray(pt_fixed, np.uint8)
        infop = None
        if info is not None:
            info = _f64(info).reshape(self.no, 3)
            infop = _d(info)
        rc = self.L.orc_ba_set(self.h, self.nc, _d(cams), cf.ctypes.data_as(_up), self.np_, _d(pts),
                               pf.ctypes.data_as(_up), self.no, oc.ctypes.data_as(_ip), op.ctypes.data_as(_ip), _d(uv),
                               infop, float(focal), float(cx), float(cy))
        if rc != 0:
            raise ValueError("orc_ba_set: observation index out of range")

    def set_robust(self, kind, param): self.L.orc_ba_set_robust(self.h, kind, float(param))
    def set_lm(self, tau=0.0, lambda_init=0.0, max_trials=0): self.L.orc_ba_set_lm(self.h, tau, lambda_init, max_trials)

    def build_structure(self):
        nb = self.L.orc_ba_build_structure(self.h)
        self.ncf = self.L.orc_ba_num_free_cameras(self.h)
        self.npf = self.L.orc_ba_num_free_points(self.h)
        colptr = np.zeros(self.ncf + 1, np.int32)
        rowidx = np.zeros(max(nb, 1), np.int32)
        self.L.orc_ba_get_structure(self.h, colptr.ctypes.data_as(_ip), rowidx.ctypes.data_as(_ip))
        self.nb = nb
        return colptr, rowidx[:nb]

    def chi2(self): return self.L.orc_ba_chi2(self.h)

    def edge_errors(self):
        e = np.zeros((self.no, 2))
        self.L.orc_ba_edge_errors(self.h, _d(e))
        return e

    def edge_jacobians(self, k):
        Jp, Jc = np.zeros((2, 3)), np.zeros((2, 6))
        self.L.orc_ba_edge_jacobians(self.h, int(k), _d(Jp), _d(Jc))
        return Jp, Jc

    def linearize(self):
        self.L.orc_ba_linearize(self.h)
        Hpp, Hll = np.zeros((self.ncf, 6, 6)), np.zeros((self.npf, 3, 3))
        Hpl, b = np.zeros((self.no, 6, 3)), np.zeros(6 * self.ncf + 3 * self.npf)
        self.L.orc_ba_get_system(self.h, _d(Hpp), _d(Hll), _d(Hpl), _d(b))
        return Hpp, Hll, Hpl, b

    def max_diag(self): return self.L.orc_ba_max_diag(self.h)

    def schur(self, lam):
        S, bs = np.zeros((self.nb, 6, 6)), np.zeros(6 * self.ncf)
        rc = self.L.orc_ba_schur(self.h, float(lam), _d(S), _d(bs))
        return rc, S, bs

    def solve(self, lam):
        x = np.zeros(6 * self.ncf + 3 * self.npf)
        rc = self.L.orc_ba_solve(self.h, float(lam), _d(x))
        return rc, x

    def update(self, x):
        x = _f64(x)
        self.L.orc_ba_update(self.h, _d(x))

    def cameras(self):
        c = np.zeros((self.nc, 7))
        self.L.orc_ba_get_cameras(self.h, _d(c))
        return c

    def points(self):
        p = np.zeros((self.np_, 3))
        self.L.orc_ba_get_points(self.h, _d(p))
        return p

    def optimize(self, max_iter, stop_rel_gain=0.0):
        hist = np.zeros((max(max_iter, 1), 4))
        chi2, lam = C.c_double(0), C.c_double(0)
        n = self.L.orc_ba_optimize(self.h, max_iter, stop_rel_gain, _d(hist), max_iter, C.byref(chi2), C.byref(lam))
        return n, chi2.value, lam.value, hist[:max(n, 0)]


# ---- trajectory alignment (evaluation modes of the reference) --------------------------------------
def umeyama(query, train, only_scale=False):
    """estimateSimilarityTransform + the RMSE loop (kitti_surf.cpp:1091-1161, :1432-1452).

    Eigen::umeyama(query, train, with_scaling=True) [EXT Eigen; S. Umeyama, "Least-squares estimation of
    transformation parameters between two point patterns", PAMI 13(4), 1991, eq. 40-43]: with the means mq,
    mt, the query variance vq = mean |q - mq|^2 and Sigma = mean (t - mt)(q - mq)^T = U D V^T,
    S = diag(1, 1, sign(det U det V)),  R = U S V^T,  c = trace(D S) / vq,  t = mt - c R mq.
    only_scale: the reference's extent-ratio variant (:1104-1136).  Returns (S221 4x4, rmse, max_dev)."""
    q = np.asarray(query, float).reshape(-1, 3)
    t = np.asarray(train, float).reshape(-1, 3)
    S = np.eye(4)
    if only_scale:
        ratio = (t.max(0) - t.min(0)) / (q.max(0) - q.min(0))
        S[:3, :3] *= 0.5 * (ratio[0] + ratio[2])
    else:
        mq, mt = q.mean(0), t.mean(0)
        vq = ((q - mq) ** 2).sum(1).mean()
        Sigma = (t - mt).T @ (q - mq) / len(q)
        U, D, Vt = np.linalg.svd(Sigma)
        sgn = np.ones(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            sgn[2] = -1
        R = U @ np.diag(sgn) @ Vt
        c = (D * sgn).sum() / vq
        S[:3, :3] = c * R
        S[:3, 3] = mt - c * R @ mq
    dev = t - (q @ S[:3, :3].T + S[:3, 3])
    d = np.sqrt((dev ** 2).sum(1))
    return S, float(np.sqrt((d ** 2).mean())), float(d.max())
