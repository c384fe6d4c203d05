"""Preconditioner study (CPU, scipy): block-Jacobi PCG vs an aggregation multigrid whose prolongation
blocks are Sim3 adjoints Ad(S_i S_root^-1) (the gauge near-null space of a pose graph in the
left-multiplicative tangent).  Research tool only -- not on any product or test path.

usage: python tools/precond_study.py [kitti1|kitti118|sphere LAPS PER] [lm_iters_before]
"""
import sys
import os
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, kitti_io  # noqa: E402
from sim3opt_b200 import synth  # noqa: E402


def quat_to_R(q):
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def skew(t):
    S = np.zeros((len(t), 3, 3))
    S[:, 0, 1] = -t[:, 2]; S[:, 0, 2] = t[:, 1]
    S[:, 1, 0] = t[:, 2]; S[:, 1, 2] = -t[:, 0]
    S[:, 2, 0] = -t[:, 1]; S[:, 2, 1] = t[:, 0]
    return S


def adjoint(R, t, s):
    """Ad_S for S=(R,t,s), tangent [omega, upsilon, sigma]."""
    n = len(t)
    A = np.zeros((n, 7, 7))
    A[:, 0:3, 0:3] = R
    A[:, 3:6, 0:3] = skew(t) @ R
    A[:, 3:6, 3:6] = s[:, None, None] * R
    A[:, 3:6, 6] = -t
    A[:, 6, 6] = 1
    return A


def rel_pose(Ri, ti, si, Rr, tr, sr):
    """S_i S_r^-1"""
    Rinv = np.transpose(Rr, (0, 2, 1))
    tinv = -np.einsum("nij,nj->ni", Rinv, tr) / sr[:, None]
    sinv = 1.0 / sr
    R = Ri @ Rinv
    t = si[:, None] * np.einsum("nij,nj->ni", Ri, tinv) + ti
    return R, t, si * sinv


def aggregate(A_pat, n, max_size=None):
    """Greedy neighbourhood aggregation on the block graph (standard AMG pass 1 + pass 2)."""
    indptr, indices = A_pat.indptr, A_pat.indices
    agg = -np.ones(n, np.int64)
    roots = []
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = indices[indptr[i]:indptr[i + 1]]
        if np.all(agg[nb] < 0):
            a = len(roots)
            roots.append(i)
            agg[i] = a
            agg[nb] = a
    for i in range(n):
        if agg[i] < 0:
            nb = indices[indptr[i]:indptr[i + 1]]
            cand = agg[nb]
            cand = cand[cand >= 0]
            if len(cand):
                agg[i] = cand[0]
            else:
                agg[i] = len(roots)
                roots.append(i)
    return agg, np.array(roots)


def block_diag_inv(A, n, d, M=None, lam=0.0):
    D = np.empty((n, d, d))
    Ab = A.tobsr((d, d))
    Ab.sort_indices()
    for i in range(n):
        for k in range(Ab.indptr[i], Ab.indptr[i + 1]):
            if Ab.indices[k] == i:
                D[i] = Ab.data[k]
    return np.linalg.inv(D)


class Level:
    pass


def build_hierarchy(A, R, t, s, d=7, min_size=40, max_levels=10, verbose=True):
    levels = []
    while True:
        n = A.shape[0] // d
        L = Level()
        L.A = A.tocsr()
        L.n = n
        L.Dinv = block_diag_inv(A, n, d)
        levels.append(L)
        if n <= min_size or len(levels) >= max_levels:
            L.dense = np.linalg.inv(A.toarray())
            break
        pat = A.tobsr((d, d))
        pat = sp.csr_matrix((np.ones(len(pat.indices)), pat.indices, pat.indptr), shape=(n, n))
        agg, roots = aggregate(pat, n)
        nc = len(roots)
        if os.environ.get("DOUBLE") and (os.environ["DOUBLE"] == "all" or (os.environ["DOUBLE"] == "1" and len(levels) == 1)
                                         or (os.environ["DOUBLE"] == "coarse" and len(levels) > 1)):
            # aggressive coarsening: aggregate the aggregate graph once more, keep the first-pass root of the coarse root
            Pa = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
            pat2 = (Pa.T @ pat @ Pa).tocsr(); pat2.data[:] = 1
            agg2, roots2 = aggregate(pat2, nc)
            agg = agg2[agg]; roots = roots[roots2]; nc = len(roots)
        Rr, tr, sr = R[roots][agg], t[roots][agg], s[roots][agg]
        Rrel, trel, srel = rel_pose(R, t, s, Rr, tr, sr)
        Pb = adjoint(Rrel, trel, srel)
        P = sp.bsr_matrix((Pb, agg, np.arange(n + 1)), shape=(n * d, nc * d)).tocsr()
        L.P = P
        A = (P.T @ A @ P).tocsr()
        R, t, s = R[roots], t[roots], s[roots]
        if verbose:
            print(f"  level {len(levels) - 1}: n={n} -> {nc}, nnz blocks/row {pat.nnz / n:.1f}")
    return levels


def bjac(L, r, d=7):
    return np.einsum("nij,nj->ni", L.Dinv, r.reshape(-1, d)).reshape(-1)


def vcycle(levels, l, r, omega=0.7, sweeps=1):
    L = levels[l]
    if l == len(levels) - 1:
        return L.dense @ r
    x = omega * bjac(L, r)
    for _ in range(sweeps - 1):
        x += omega * bjac(L, r - L.A @ x)
    rc = L.P.T @ (r - L.A @ x)
    x += L.P @ vcycle(levels, l + 1, rc, omega, sweeps)
    for _ in range(sweeps):
        x += omega * bjac(L, r - L.A @ x)
    return x


def acycle(levels, l, r, omega=0.7, sweeps=1, additive_levels=1):
    """additive at the first `additive_levels` levels (no extra fine-level products), V-cycle below"""
    L = levels[l]
    if l == len(levels) - 1:
        return L.dense @ r
    if l >= additive_levels:
        return vcycle(levels, l, r, omega, sweeps)
    return omega * bjac(L, r) + L.P @ acycle(levels, l + 1, L.P.T @ r, omega, sweeps, additive_levels)


def pcg(A, b, M, tol, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    r0 = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        if np.linalg.norm(r) <= tol * r0:
            return x, it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "sphere"
    if which.startswith("kitti"):
        g = kitti_io.build_kitti_sim3_graph(os.path.join(ROOT, "tests", "golden", "kitti00"), which == "kitti1")
        pre = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        orc.set_math_mode(orc.MATH_REFERENCE)
    else:
        laps, per = int(sys.argv[2]), int(sys.argv[3])
        pre = int(sys.argv[4]) if len(sys.argv) > 4 else 1
        g = synth.sphere(laps, per, seed=42)
        orc.set_math_mode(orc.MATH_CORRECTED)
    p = orc.Problem(orc.KIND_SIM3)
    p.set_vertices(g["est"], g["fixed"])
    p.set_edges(g["v0"], g["v1"], g["meas"], g.get("info"))
    p.set_jacobian_mode(orc.JAC_ANALYTIC)
    colptr, rowidx = p.build_structure()
    lam = None
    if pre > 0:
        n, chi2, lam, hist = p.optimize(pre)
        print("after", n, "LM iterations: chi2", chi2, "lambda", lam)
    H, b = p.linearize()
    if lam is None:
        lam = 1e-5 * p.max_diag()
    if os.environ.get("LAM"):
        lam = float(os.environ["LAM"])      # study a later-iteration damping without running the LM there
    print("lambda", lam, "max diag", p.max_diag())
    nf, d = len(colptr) - 1, 7
    cols = np.repeat(np.arange(nf), np.diff(colptr))
    up = sp.bsr_matrix((np.ascontiguousarray(H.transpose(0, 2, 1)), rowidx, colptr), shape=(nf * d, nf * d))
    # blocks are stored by column (CCS): bsr with indptr=colptr gives the transpose of the upper part
    U = up.T.tocsr()  # upper triangle incl. diagonal
    Dg = sp.block_diag([H[colptr[c + 1] - 1] for c in range(nf)], format="csr") if False else None
    A = U + U.T
    # remove the doubled diagonal blocks
    diag_blocks = np.array([H[colptr[c + 1] - 1] for c in range(nf)])
    assert np.all(rowidx[colptr[1:] - 1] == np.arange(nf))
    Dm = sp.bsr_matrix((diag_blocks, np.arange(nf), np.arange(nf + 1)), shape=(nf * d, nf * d)).tocsr()
    A = (A - Dm + lam * sp.identity(nf * d)).tocsr()
    est = p.vertices()
    free = np.where(np.asarray(g["fixed"]) == 0)[0]
    S = est[free]
    R, t, s = quat_to_R(S[:, 0:4]), S[:, 4:7], S[:, 7]

    tols = (1e-3,) if os.environ.get("FAST") else (1e-3, 1e-8)
    for tol in tols:
        L0 = Level(); L0.Dinv = block_diag_inv(A, nf, d)
        t0 = time.time()
        x, it = pcg(A, b, lambda r: bjac(L0, r), tol)
        print(f"block-Jacobi PCG tol {tol:g}: {it} iterations ({time.time() - t0:.1f}s)")
    t0 = time.time()
    levels = build_hierarchy(A, R, t, s)
    print(f"hierarchy: {len(levels)} levels ({time.time() - t0:.1f}s)")
    for omega in (0.6, 0.8, 1.0):
        for nadd in ((1, 2, 99) if os.environ.get("FAST") else (1, 99)):
            for tol in tols:
                x, it = pcg(A, b, lambda r: acycle(levels, 0, r, omega, 1, nadd), tol)
                print(f"additive x{nadd} (omega={omega}) PCG tol {tol:g}: {it} iterations")
    for omega in (() if os.environ.get("FAST") else (0.6, 0.8)):
        for sweeps in (1, 2):
            for tol in (1e-3, 1e-8):
                x, it = pcg(A, b, lambda r: vcycle(levels, 0, r, omega, sweeps), tol)
                print(f"AMG(V, omega={omega}, sweeps={sweeps}) PCG tol {tol:g}: {it} iterations")


if __name__ == "__main__":
    main()
