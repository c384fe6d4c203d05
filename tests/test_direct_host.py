"""CPU tests of the DIRECT linear solver's symbolic analysis (sim3opt_b200/csrc/direct_host.cpp): the plan is
executed here in numpy exactly the way direct.cu executes it (scatter, per round: gather lists -> pivot
Cholesky -> column scaling; triangular solves by rounds) and checked against a dense solve."""
import numpy as np
import pytest

import sim3opt_b200 as s3


def bsr_upper(nf, hidx, v0, v1):
    """BSR-upper block list in the library's order: rows ascending, diagonal first, then columns ascending."""
    rows = [set() for _ in range(nf)]
    for a, b in zip(v0, v1):
        ha, hb = hidx[a], hidx[b]
        if ha >= 0 and hb >= 0:
            rows[min(ha, hb)].add(max(ha, hb))
    blocks = []
    for r in range(nf):
        blocks.append((r, r))
        blocks += [(r, c) for c in sorted(rows[r])]
    return blocks


def emulate(plan, blocks, Hb, lam, b, d):
    n, cptr, brow, src = plan["n"], plan["cptr"], plan["brow"], plan["src"]
    nL = len(brow)
    bcol = np.zeros(nL, int)
    for j in range(n):
        bcol[cptr[j]:cptr[j + 1]] = j
    L = np.zeros((nL, d, d))
    for t in range(nL):
        if src[t] >= 0:
            M = Hb[src[t] >> 1]
            L[t] = M.T if (src[t] & 1) else M
        if brow[t] == bcol[t]:
            L[t] += lam * np.eye(d)
    Linv = np.zeros((n, d, d))
    lp = plan["lev_ptr"]
    done = np.zeros(nL, bool)
    for lev in range(plan["rounds"]):
        c0, c1 = lp[lev], lp[lev + 1]
        for t in range(cptr[c0], cptr[c1]):
            for q in range(plan["upd_ptr"][t], plan["upd_ptr"][t + 1]):
                a, bb = plan["upd_a"][q], plan["upd_b"][q]
                assert done[a] and done[bb], "update reads a block of the same or a later round"
                L[t] -= L[a] @ L[bb].T
        for k in range(c0, c1):
            C = np.linalg.cholesky(L[cptr[k]])
            L[cptr[k]] = C
            Linv[k] = np.linalg.inv(C)
        for t in range(cptr[c0], cptr[c1]):
            if brow[t] != bcol[t]:
                L[t] = L[t] @ Linv[bcol[t]].T
            done[t] = True
    y = np.zeros((n, d))
    for k in range(n):          # elimination order is a valid sequential order
        acc = b[plan["perm"][k]].copy()
        for j in range(k):
            for t in range(cptr[j] + 1, cptr[j + 1]):
                if brow[t] == k:
                    acc -= L[t] @ y[j]
        y[k] = Linv[k] @ acc
    x = np.zeros((n, d))
    for k in range(n - 1, -1, -1):
        acc = y[k].copy()
        for t in range(cptr[k] + 1, cptr[k + 1]):
            acc -= L[t].T @ y[brow[t]]
        y[k] = Linv[k].T @ acc
        x[plan["perm"][k]] = y[k]
    return x


def check_graph(g, d=3, seed=0):
    nv = len(g["est"])
    fixed = np.asarray(g["fixed"], np.uint8)
    plan = s3.host_direct_plan(nv, fixed, g["v0"], g["v1"])
    hidx = -np.ones(nv, int)
    hidx[fixed == 0] = np.arange(int((fixed == 0).sum()))
    nf = plan["n"]
    blocks = bsr_upper(nf, hidx, g["v0"], g["v1"])
    # rounds are independent sets; perm is a permutation; every block of the pattern is sourced exactly once
    assert sorted(plan["perm"]) == list(range(nf))
    srcs = plan["src"][plan["src"] >= 0] >> 1
    assert sorted(srcs) == list(range(len(blocks)))
    rng = np.random.default_rng(seed)
    Hb = rng.standard_normal((len(blocks), d, d))
    A = np.zeros((nf * d, nf * d))
    for k, (r, c) in enumerate(blocks):
        if r == c:
            Hb[k] = Hb[k] @ Hb[k].T + 4 * len(blocks) / nf * d * np.eye(d)      # diagonally dominant: SPD
            A[r * d:(r + 1) * d, r * d:(r + 1) * d] = Hb[k]
        else:
            A[r * d:(r + 1) * d, c * d:(c + 1) * d] = Hb[k]
            A[c * d:(c + 1) * d, r * d:(r + 1) * d] = Hb[k].T
    lam = 0.37
    b = rng.standard_normal((nf, d))
    x = emulate(plan, blocks, Hb, lam, b, d)
    ref = np.linalg.solve(A + lam * np.eye(nf * d), b.reshape(-1)).reshape(nf, d)
    assert np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max()
    return plan


def test_kitti_plans_are_shallow(kitti_k1, kitti_k118):
    # chain-dominated graphs: the multiple-minimum-degree rounds are a cyclic reduction
    p1 = check_graph(kitti_k1)
    assert p1["rounds"] <= 16 and p1["n_pairs"] < 5000
    p118 = check_graph(kitti_k118)
    assert p118["rounds"] <= 40 and p118["n_pairs"] < 10000


def test_mesh_plans_solve(sphere_small, manhattan_small):
    check_graph(sphere_small, d=2)
    check_graph(manhattan_small, d=2)


def test_plan_limits_and_edge_cases():
    # a graph with every vertex fixed but one, no free-free edge
    g = {"est": np.zeros((3, 8)), "fixed": np.array([1, 0, 1], np.uint8), "v0": np.array([0, 1], np.int32), "v1": np.array([1, 2], np.int32)}
    plan = s3.host_direct_plan(3, g["fixed"], g["v0"], g["v1"])
    assert plan["n"] == 1 and plan["rounds"] == 1 and plan["n_pairs"] == 0 and list(plan["src"]) == [0]
    # clique of 12: every round eliminates one vertex; the pair budget is enforced
    n = 12
    v0, v1 = np.array([(a, b) for a in range(n) for b in range(a + 1, n)], np.int32).T
    plan = s3.host_direct_plan(n, None, v0, v1)
    assert plan["rounds"] == n and plan["n_pairs"] == sum(c * (c + 1) // 2 for c in range(n))
    with pytest.raises(s3.S3OError):
        s3.host_direct_plan(n, None, v0, v1, max_pairs=10)
