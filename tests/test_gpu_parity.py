"""GPU parity tests (run on the B200 box with -m gpu): every call goes through the C ABI.

Lock-step protocol of SURVEY.md section 8(d): identical state and lambda into the oracle and the GPU
path, then compare chi2 (rel <= 1e-12), block structure (bit-exact), every H block and b
(rel <= 1e-10 against the same-Jacobian-mode oracle), the damped step by backward error
|(H + lambda I) x - b| / |b| <= 1e-10, and the post-retraction states (<= 1e-12).
"""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def _orc():
    from oracle import oracle as orc
    return orc


def graphs(request):
    return request.getfixturevalue(request.param)


@pytest.fixture(params=["kitti_k1", "kitti_k118", "sphere_small", "manhattan_small"])
def graph(request):
    return request.getfixturevalue(request.param)


def test_structure_bit_exact(graph):
    gpu, cpu = make_gpu(graph), make_oracle(graph)
    cp_g, ri_g = gpu.build_structure()
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(cp_g, cp_c) and np.array_equal(ri_g, ri_c)
    assert np.array_equal(gpu.hessian_index(), cpu.hessian_index())


def test_chi2_and_edge_errors(graph):
    gpu, cpu = make_gpu(graph), make_oracle(graph)
    c_g, c_c = gpu.chi2(), cpu.chi2()
    assert abs(c_g - c_c) <= 1e-12 * c_c
    e_g, e_c = gpu.edge_errors(), cpu.edge_errors()
    assert np.abs(e_g - e_c).max() <= 1e-11 * max(1.0, np.abs(e_c).max())
    # determinism: the reduction is bitwise reproducible
    assert gpu.chi2() == c_g


def test_k1_known_answers_on_gpu(kitti_k1):
    gpu = make_gpu(kitti_k1)
    colptr, rowidx = gpu.build_structure()
    assert len(rowidx) == 1540
    assert abs(gpu.chi2() - 169.9259622426238) <= 1e-9
    ref = np.array([0.010350516, 0.013424595, 0.004786277, 11.481364008, -0.514528686, 5.919704247, 1.672212412])
    assert np.allclose(gpu.edge_errors()[0], ref, atol=5e-9)


@pytest.mark.parametrize("jac", [0, 1])
def test_linearize_lockstep(graph, jac):
    gpu, cpu = make_gpu(graph, jac=jac), make_oracle(graph, jac=jac)
    Hg, bg = gpu.linearize()
    Hc, bc = cpu.linearize()
    # numeric mode differentiates a round-off-limited function: 1e-9 steps amplify 1e-16 noise to 1e-7
    tol = 1e-10 if jac == 1 else 2e-5
    assert np.abs(Hg - Hc).max() <= tol * np.abs(Hc).max()
    assert np.abs(bg - bc).max() <= tol * np.abs(bc).max()
    assert abs(gpu.max_diag() - cpu.max_diag()) <= tol * cpu.max_diag()
    # bitwise reproducible assembly (no floating-point atomics)
    Hg2, bg2 = gpu.linearize()
    assert np.array_equal(Hg, Hg2) and np.array_equal(bg, bg2)


def dense_from_blocks(colptr, rowidx, H, d):
    nf = len(colptr) - 1
    A = np.zeros((nf * d, nf * d))
    for c in range(nf):
        for k in range(colptr[c], colptr[c + 1]):
            r = rowidx[k]
            A[r * d:(r + 1) * d, c * d:(c + 1) * d] = H[k]
            if r != c:
                A[c * d:(c + 1) * d, r * d:(r + 1) * d] = H[k].T
    return A


def test_spmv_against_dense(graph):
    gpu = make_gpu(graph, jac=1)
    colptr, rowidx = gpu.build_structure()
    H, b = gpu.linearize()
    A = dense_from_blocks(colptr, rowidx, H, 7)
    rng = np.random.default_rng(1)
    x = rng.normal(size=A.shape[0])
    lam = 0.37
    y = gpu.hessian_multiply(lam, x)
    ref = A @ x + lam * x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()


def test_solve_backward_error_and_update(graph):
    orc = _orc()
    gpu, cpu = make_gpu(graph, jac=1), make_oracle(graph, jac=orc.JAC_ANALYTIC)
    colptr, rowidx = gpu.build_structure()
    H, b = gpu.linearize()
    cpu.linearize()
    lam = 1e-5 * cpu.max_diag()
    gpu.set_pcg(1e-11, 50000)
    rc, x, iters, rel = gpu.solve(lam)
    assert rc == 0 and rel <= 1e-11
    A = dense_from_blocks(colptr, rowidx, H, 7) + lam * np.eye(len(b))
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)
    # same step into both retractions
    rc_c, x_c = cpu.solve(lam)
    assert rc_c == 0
    gpu.update(x_c)
    cpu.update(x_c)
    assert np.abs(gpu.vertices() - cpu.vertices()).max() <= 1e-12 * max(1.0, np.abs(cpu.vertices()).max())
    assert abs(gpu.chi2() - cpu.chi2()) <= 1e-10 * cpu.chi2()


def test_lm_first_iteration_matches_oracle(graph):
    """Iteration 0 is stable across implementations (SURVEY.md 0.A take-away 1)."""
    orc = _orc()
    gpu, cpu = make_gpu(graph, jac=1), make_oracle(graph, jac=orc.JAC_ANALYTIC)
    # K1/K118 have cond(H + lambda I) ~ 1e6..1e9: the chi2 after the step is sensitive
    # to the forward error of the linear solve, so the PCG runs to its fp64 floor here
    gpu.set_pcg(1e-14, 100000)
    n_g, chi_g, lam_g, hist_g = gpu.optimize(1)
    n_c, chi_c, lam_c, hist_c = cpu.optimize(1)
    assert n_g == n_c == 1
    assert abs(chi_g - chi_c) <= 1e-4 * chi_c          # BASELINE chi2 tolerance
    assert abs(lam_g - lam_c) <= 1e-3 * lam_c
    assert hist_g[0, 2] == hist_c[0, 2]


def test_kitti_final_chi2_not_worse_than_oracle(kitti_k1):
    """KITTI K1 is chaotic from iteration 1 on: gate on the final chi2 only."""
    orc = _orc()
    gpu, cpu = make_gpu(kitti_k1, jac=1), make_oracle(kitti_k1, jac=orc.JAC_ANALYTIC)
    gpu.set_pcg(1e-14, 100000)
    n_g, chi_g, _, hist_g = gpu.optimize(10)
    n_c, chi_c, _, hist_c = cpu.optimize(10)
    assert abs(hist_g[0, 0] - hist_c[0, 0]) <= 1e-4 * hist_c[0, 0]
    assert abs(hist_g[1, 0] - 0.4886) <= 1e-3
    assert chi_g <= chi_c * (1 + 1e-2)


def test_sphere_end_to_end_tolerances(sphere_small):
    """BASELINE tolerances on a graph with a unique minimum: chi2 1e-4 rel, 1e-4 m, 1e-5 rad."""
    orc = _orc()
    import sim3opt_b200 as s3
    orc.set_math_mode(orc.MATH_CORRECTED)
    try:
        gpu = make_gpu(sphere_small, jac=1, math_mode=s3.MATH_CORRECTED)
        cpu = make_oracle(sphere_small, jac=orc.JAC_ANALYTIC)
        gpu.set_pcg(1e-10, 20000)
        n_g, chi_g, _, hist_g = gpu.optimize(40)
        n_c, chi_c, _, hist_c = cpu.optimize(40)
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)
    assert abs(chi_g - chi_c) <= 1e-4 * chi_c
    vg, vc = gpu.vertices(), cpu.vertices()
    assert np.abs(vg[:, 4:7] - vc[:, 4:7]).max() <= 1e-4
    # rotation difference angle from quaternions
    dots = np.abs((vg[:, :4] * vc[:, :4]).sum(1) / (np.linalg.norm(vg[:, :4], axis=1) * np.linalg.norm(vc[:, :4], axis=1)))
    assert (2 * np.arccos(np.clip(dots, -1, 1))).max() <= 1e-5
    assert np.abs(vg[:, 7] - vc[:, 7]).max() <= 1e-5


def test_manhattan_end_to_end_tolerances(manhattan_small):
    """Same gate on the irregular Manhattan-3D variant of config 3 (ragged block rows, many loop edges)."""
    orc = _orc()
    import sim3opt_b200 as s3
    orc.set_math_mode(orc.MATH_CORRECTED)
    try:
        gpu = make_gpu(manhattan_small, jac=1, math_mode=s3.MATH_CORRECTED)
        cpu = make_oracle(manhattan_small, jac=orc.JAC_ANALYTIC)
        gpu.set_pcg(1e-10, 20000)
        n_g, chi_g, _, _ = gpu.optimize(30)
        n_c, chi_c, _, _ = cpu.optimize(30)
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)
    assert abs(chi_g - chi_c) <= 1e-4 * chi_c
    vg, vc = gpu.vertices(), cpu.vertices()
    assert np.abs(vg[:, 4:7] - vc[:, 4:7]).max() <= 1e-4
    dots = np.abs((vg[:, :4] * vc[:, :4]).sum(1) / (np.linalg.norm(vg[:, :4], axis=1) * np.linalg.norm(vc[:, :4], axis=1)))
    assert (2 * np.arccos(np.clip(dots, -1, 1))).max() <= 1e-5
    assert np.abs(vg[:, 7] - vc[:, 7]).max() <= 1e-5


def test_robust_kernels_lockstep(sphere_small):
    orc = _orc()
    import sim3opt_b200 as s3
    for kind, param in ((s3.ROBUST_HUBER, 2.5), (s3.ROBUST_PTAM_TUKEY, 400.0), (s3.ROBUST_PTAM_CAUCHY, 50.0),
                        (s3.ROBUST_PTAM_HUBER, 30.0)):
        gpu = make_gpu(sphere_small, jac=1, robust=(kind, param))
        cpu = make_oracle(sphere_small, jac=orc.JAC_ANALYTIC, robust=(kind, param))
        c_g, c_c = gpu.chi2(), cpu.chi2()
        assert abs(c_g - c_c) <= 1e-11 * c_c
        Hg, bg = gpu.linearize()
        Hc, bc = cpu.linearize()
        assert np.abs(Hg - Hc).max() <= 1e-10 * np.abs(Hc).max()
        assert np.abs(bg - bc).max() <= 1e-10 * np.abs(bc).max()
    # PTAM sigma estimate
    gpu = make_gpu(sphere_small)
    cpu = make_oracle(sphere_small)
    e = cpu.edge_errors()
    chi = np.einsum("ni,nij,nj->n", e, sphere_small["info"], e)
    for kind in (s3.ROBUST_PTAM_TUKEY, s3.ROBUST_PTAM_HUBER, s3.ROBUST_PTAM_LS):
        assert np.isclose(gpu.estimate_sigma_squared(kind), orc.ptam_find_sigma_squared(kind, chi), rtol=1e-10)


def test_scale_trans_lockstep(kitti_k1):
    orc = _orc()
    import sim3opt_b200 as s3
    from oracle import kitti_io
    st = kitti_io.to_scale_trans_graph(kitti_k1)
    gpu = make_gpu(st, kind=s3.KIND_SCALE_TRANS, jac=1)
    cpu = make_oracle(st, kind=orc.KIND_SCALE_TRANS, jac=orc.JAC_ANALYTIC)
    cp_g, ri_g = gpu.build_structure()
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(cp_g, cp_c) and np.array_equal(ri_g, ri_c)
    assert abs(gpu.chi2() - cpu.chi2()) <= 1e-12 * cpu.chi2()
    Hg, bg = gpu.linearize()
    Hc, bc = cpu.linearize()
    assert np.abs(Hg - Hc).max() <= 1e-10 * np.abs(Hc).max()
    assert np.abs(bg - bc).max() <= 1e-10 * np.abs(bc).max()
    gpu.set_pcg(1e-12, 100000)
    n_g, chi_g, lam_g, _ = gpu.optimize(1)
    n_c, chi_c, lam_c, _ = cpu.optimize(1)
    assert abs(chi_g - chi_c) <= 1e-6 * max(chi_c, 1e-12)


@pytest.mark.parametrize("kind_name", ["SCALE_TRANS", "SCALE"])
def test_scale_model_logratio_lockstep(kitti_k1, kind_name):
    """Row a18: log-ratio scale error + multiplicative scale update (s3o_set_scale_model), device vs oracle:
    chi2, H, b at a perturbed estimate and three LM iterations in lock step."""
    orc = _orc()
    import sim3opt_b200 as s3
    from oracle import kitti_io
    st = kitti_io.to_scale_trans_graph(kitti_k1)
    g = dict(st)
    if kind_name == "SCALE":
        g = dict(est=st["est"][:, :1].copy(), fixed=st["fixed"], v0=st["v0"], v1=st["v1"], meas=st["meas"][:, :1].copy())
    rng = np.random.default_rng(11)
    g["est"] = g["est"].copy()
    g["est"][:, 0] *= np.exp(0.05 * rng.standard_normal(len(g["est"])))
    for jac in (0, 1):
        gpu = make_gpu(g, kind=getattr(s3, "KIND_" + kind_name), jac=jac)
        cpu = make_oracle(g, kind=getattr(orc, "KIND_" + kind_name), jac=jac)
        gpu.set_scale_model(s3.SCALE_MODEL_LOGRATIO)
        cpu.set_scale_model(1)
        assert abs(gpu.chi2() - cpu.chi2()) <= 1e-12 * cpu.chi2()
        Hg, bg = gpu.linearize()
        Hc, bc = cpu.linearize()
        # numeric mode differentiates a round-off-limited function (test_linearize_lockstep)
        tol = 1e-9 if jac == 1 else 2e-5
        assert np.abs(Hg - Hc).max() <= tol * np.abs(Hc).max()
        assert np.abs(bg - bc).max() <= tol * np.abs(bc).max()
        if jac == 0:
            continue
        gpu.set_pcg(1e-13, 100000)
        for it in range(3):
            n_g, chi_g, lam_g, _ = gpu.optimize(1)
            n_c, chi_c, lam_c, _ = cpu.optimize(1)
            assert abs(chi_g - chi_c) <= 1e-7 * max(chi_c, 1e-12), (it, chi_g, chi_c)
        assert np.abs(gpu.vertices() - cpu.vertices()).max() <= 1e-6
    # the null-vector stage is defined on the DIFFERENCE rows only
    if kind_name == "SCALE":
        with pytest.raises(s3.S3OError):
            gpu.smallest_eigenvector()
    # Sim3 problems have no scale model to switch
    with pytest.raises(s3.S3OError):
        s3.Problem(s3.KIND_SIM3).set_scale_model(1)


def test_edge_cases():
    import sim3opt_b200 as s3
    p = s3.Problem(s3.KIND_SIM3)
    I = np.array([[0, 0, 0, 1, 0, 0, 0, 1.0]])
    # invalid edge index is rejected with an error code, not a crash
    p.set_vertices(np.repeat(I, 3, 0), [1, 0, 0])
    with pytest.raises(s3.S3OError):
        p.set_edges([0], [7], I)
    # optimize with every vertex fixed reports "0 vertices to optimize"
    p.set_vertices(np.repeat(I, 2, 0), [1, 1])
    p.set_edges([0], [1], I)
    with pytest.raises(s3.S3OError):
        p.optimize(1)
    # an empty edge set has chi2 = 0
    p.set_vertices(np.repeat(I, 2, 0), [1, 0])
    p.set_edges(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 8)))
    assert p.chi2() == 0.0
    # duplicate edges between one pair accumulate into a single block
    est = np.repeat(I, 3, 0)
    est[1, 4] = 1.0
    est[2, 5] = 2.0
    p.set_vertices(est, [1, 0, 0])
    p.set_edges([1, 2, 1], [2, 1, 2], np.repeat(I, 3, 0))
    colptr, rowidx = p.build_structure()
    assert list(colptr) == [0, 1, 3] and list(rowidx) == [0, 0, 1]
    from conftest import make_oracle
    g = dict(est=est, fixed=np.array([1, 0, 0], np.uint8), v0=np.array([1, 2, 1], np.int32),
             v1=np.array([2, 1, 2], np.int32), meas=np.repeat(I, 3, 0))
    cpu = make_oracle(g, jac=1)
    Hg, bg = p.linearize()
    Hc, bc = cpu.linearize()
    assert np.abs(Hg - Hc).max() <= 1e-12 * np.abs(Hc).max()
    assert np.abs(bg - bc).max() <= 1e-12 * np.abs(bc).max()


def test_resume_and_snapshot(sphere_small):
    """optimize(3) == optimize(1) x 3 with resume; restore brings back the snapshot bit for bit."""
    import sim3opt_b200 as s3
    a = make_gpu(sphere_small, jac=1, math_mode=s3.MATH_CORRECTED)
    b = make_gpu(sphere_small, jac=1, math_mode=s3.MATH_CORRECTED)
    a.build_structure()
    a.snapshot_estimates()
    v0 = a.vertices()
    n, chi_a, lam_a, hist_a = a.optimize(3)
    b.set_lm_resume(True)
    for _ in range(3):
        n_b, chi_b, lam_b, _ = b.optimize(1)
    assert chi_a == chi_b and lam_a == lam_b
    assert np.array_equal(a.vertices(), b.vertices())
    a.restore_estimates()
    assert np.array_equal(a.vertices(), v0)
    n2, chi_a2, _, _ = a.optimize(3)
    assert chi_a2 == chi_a                              # bitwise reproducible solve


@pytest.mark.parametrize("robust", [None, (1, 2.5)])
def test_dense_information_matrices(sphere_small, robust):
    """Full (non-diagonal) SPD information matrices take the dense linearisation path; diagonal ones and NULL
    (identity) the cheaper one -- both against the oracle, with and without a robust kernel."""
    orc = _orc()
    g = dict(sphere_small)
    rng = np.random.default_rng(12)
    M = rng.normal(0, 1, (len(g["v0"]), 7, 7))
    g["info"] = np.einsum("nij,nkj->nik", M, M) + 7 * np.eye(7)
    for graph in (g, sphere_small, {k: v for k, v in sphere_small.items() if k != "info"}):
        gpu, cpu = make_gpu(graph, jac=1, robust=robust), make_oracle(graph, jac=orc.JAC_ANALYTIC, robust=robust)
        c_g, c_c = gpu.chi2(), cpu.chi2()
        assert abs(c_g - c_c) <= 1e-11 * c_c
        Hg, bg = gpu.linearize()
        Hc, bc = cpu.linearize()
        assert np.abs(Hg - Hc).max() <= 1e-10 * np.abs(Hc).max()
        assert np.abs(bg - bc).max() <= 1e-10 * np.abs(bc).max()


def test_duplicate_edges_share_a_block(sphere_small):
    """Several edges between one vertex pair (both orientations) sum into one Hessian block in edge order."""
    orc = _orc()
    g = dict(sphere_small)
    rng = np.random.default_rng(8)
    pick = rng.choice(len(g["v0"]), 25, replace=False)
    v0 = np.concatenate([g["v0"], g["v0"][pick[:15]], g["v1"][pick[15:]]]).astype(np.int32)       # last 10 reversed
    v1 = np.concatenate([g["v1"], g["v1"][pick[:15]], g["v0"][pick[15:]]]).astype(np.int32)
    meas_rev = np.array([orc.sim3_inv(m) for m in g["meas"][pick[15:]]])
    g["v0"], g["v1"] = v0, v1
    g["meas"] = np.concatenate([g["meas"], g["meas"][pick[:15]], meas_rev])
    g["info"] = np.concatenate([g["info"], g["info"][pick]])
    gpu, cpu = make_gpu(g, jac=1), make_oracle(g, jac=orc.JAC_ANALYTIC)
    cp_g, ri_g = gpu.build_structure()
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(cp_g, cp_c) and np.array_equal(ri_g, ri_c)
    Hg, bg = gpu.linearize()
    Hc, bc = cpu.linearize()
    assert np.abs(Hg - Hc).max() <= 1e-10 * np.abs(Hc).max()
    assert np.abs(bg - bc).max() <= 1e-10 * np.abs(bc).max()
    Hg2, _ = gpu.linearize()
    assert np.array_equal(Hg, Hg2)


def test_hub_vertex_rows_longer_than_a_tile():
    """A star graph: one block row holds 400 off-diagonal blocks (> one SpMV tile)."""
    orc = _orc()
    rng = np.random.default_rng(3)
    n = 402
    est = np.zeros((n, 8))
    for k in range(n):
        est[k] = orc.sim3_exp(np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 2, 3), rng.normal(0, 0.1, 1)]))
    fixed = np.zeros(n, np.uint8)
    fixed[0] = 1
    v0 = np.concatenate([[0], np.ones(n - 2, np.int32), np.arange(2, n - 1)]).astype(np.int32)
    v1 = np.concatenate([[1], np.arange(2, n), np.arange(3, n)]).astype(np.int32)
    meas = np.zeros((len(v0), 8))
    for e in range(len(v0)):
        noise = orc.sim3_exp(rng.normal(0, 0.05, 7))
        meas[e] = orc.sim3_mul(noise, orc.sim3_mul(est[v1[e]], orc.sim3_inv(est[v0[e]])))
    g = dict(est=est, fixed=fixed, v0=v0, v1=v1, meas=meas)
    gpu, cpu = make_gpu(g, jac=1), make_oracle(g, jac=orc.JAC_ANALYTIC)
    colptr, rowidx = gpu.build_structure()
    cp_c, ri_c = cpu.build_structure()
    assert np.array_equal(colptr, cp_c) and np.array_equal(rowidx, ri_c)
    Hg, bg = gpu.linearize()
    Hc, bc = cpu.linearize()
    assert np.abs(Hg - Hc).max() <= 1e-10 * np.abs(Hc).max()
    A = dense_from_blocks(colptr, rowidx, Hg, 7)
    x = rng.normal(size=A.shape[0])
    y = gpu.hessian_multiply(0.5, x)
    ref = A @ x + 0.5 * x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    gpu.set_pcg(1e-12, 5000)
    rc, xs, iters, rel = gpu.solve(1e-3)
    assert rc == 0
    Ad = A + 1e-3 * np.eye(len(bg))
    assert np.linalg.norm(Ad @ xs - bg) <= 1e-10 * np.linalg.norm(bg)


def test_stop_rules(sphere_small):
    """s3o_set_stop_rules: the step rule ends a solve at the oracle's fixed point without burning trials at the
    fp64 floor; stop_reason reports which rule fired; without rules g2o's Terminate (10 failed trials) ends it."""
    orc = _orc()
    import sim3opt_b200 as s3
    orc.set_math_mode(orc.MATH_CORRECTED)
    try:
        cpu = make_oracle(sphere_small, jac=orc.JAC_ANALYTIC)
        n_c, chi_c, _, _ = cpu.optimize(60)
        vc = cpu.vertices()
        gpu = make_gpu(sphere_small, jac=1, math_mode=s3.MATH_CORRECTED)
        gpu.set_pcg(1e-10, 20000)
        gpu.set_stop_rules(1e-6, 1e-13)
        n_g, chi_g, _, hist = gpu.optimize(60)
        st = gpu.stats()
        assert st["stop_reason"] in (2, 3) and n_g < 60
        assert hist[:, 2].max() <= 2                       # no trial storm at the end
        assert abs(chi_g - chi_c) <= 1e-9 * chi_c
        assert np.abs(gpu.vertices()[:, 4:7] - vc[:, 4:7]).max() <= 1e-5
        assert st["est_distance"] <= 1e-6 or st["stop_reason"] == 3
        free = make_gpu(sphere_small, jac=1, math_mode=s3.MATH_CORRECTED)
        free.set_pcg(1e-10, 20000)
        n_f, _, _, hist_f = free.optimize(60)
        assert free.stats()["stop_reason"] in (0, 4) and n_f >= n_g
    finally:
        orc.set_math_mode(orc.MATH_REFERENCE)


@pytest.mark.gpu
def test_estimate_slices_without_a_communicator(sphere_small):
    """s3o_*_slice on one GPU: the slice is the whole vertex range and the calls equal s3o_set_estimates / s3o_get_vertices
    (the 2-rank behaviour is checked by tools/dist_check.py under tests/test_gpu_dist.py)."""
    g = sphere_small
    p = make_gpu(g)
    assert p.estimate_slice() == (0, len(g["est"]))
    est = np.ascontiguousarray(g["est"] * 1.0)
    est[:, 4:7] += 0.5
    p.set_estimates_slice(est)
    out = np.zeros_like(est)
    p.vertices_slice(out)
    assert np.array_equal(out, est) and np.array_equal(p.vertices(), est)
    q = make_gpu(g)
    q.set_estimates(est)
    assert p.chi2() == q.chi2()
