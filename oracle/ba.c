/* ba.c -- bundle-adjustment oracle.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the reference's BA path, bal_example.cpp:44-243:
 *   cameras  g2o::VertexSE3Expmap (SE3Quat world->camera, 6-DoF, oplus = exp(delta) * T), ids 0..C-1
 *   points   g2o::VertexSBAPointXYZ (3-DoF, additive), marginalised                 bal_example.cpp:119-129
 *   edges    g2o::EdgeProjectXYZ2UV, vertex(0) = point, vertex(1) = camera,
 *            e = z - (f * (x/z, y/z) + pp),  x = T.map(p)                            bal_example.cpp:131-166
 *            analytic linearizeOplus (g2o types_six_dof_expmap.cpp)
 *   robust   RobustKernelHuber, delta 2.5                                           bal_example.cpp:149-153
 *   solver   OptimizationAlgorithmLevenberg(BlockSolver_6_3(LinearSolverEigen)):
 *            Schur complement on the points, sparse LDLT of H_schur                 bal_example.cpp:73-85
 * g2o is not vendored, so its algorithm is restated from SURVEY.md sections 3.1 / 3.3 / 8(a).
 * PARITY STATUS: unpinned by the reference (no tests, and both call sites pass a wrong argc,
 * SURVEY.md 0.5); pinned by self-consistency (numeric-vs-analytic Jacobians, Schur vs full solve).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ldlt.h"
#include "oracle.h"

struct orc_ba {
    int nc, np, no;
    double *cam, *pt;              /* nc x 7, np x 3 */
    unsigned char *cfix, *pfix;
    int *ocam, *opt;
    double *uv, *info;             /* no x 2, no x 3 (xx, xy, yy) */
    double f, cx, cy;
    int robust_kind;
    double robust_param;
    double tau, user_lambda;
    int max_trials;
    /* structure */
    int ncf, npf;
    int *chidx, *phidx;            /* Hessian index among free cameras / free points, -1 fixed */
    int nb, *colptr, *rowidx;      /* H_schur upper block-CCS */
    int *pt_ptr, *pt_obs;          /* observations per point (CSR, observation order) */
    /* system */
    double *Hpp, *Hll, *Hpl, *b;   /* ncf x 36, npf x 9, no x 18, 6 ncf + 3 npf */
    double *S, *bs;                /* nb x 36, 6 ncf */
    orc_ldlt *ldlt;
};

static void quat_rot(const double q[4], const double v[3], double out[3]) {
    double R[9];
    orc_quat_to_rot(q, R);
    for (int i = 0; i < 3; ++i) out[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}

orc_ba *orc_ba_create(void) {
    orc_ba *p = (orc_ba *)calloc(1, sizeof(orc_ba));
    p->tau = 1e-5;
    p->max_trials = 10;
    return p;
}

static void free_structure(orc_ba *p) {
    free(p->chidx); free(p->phidx); free(p->colptr); free(p->rowidx); free(p->pt_ptr); free(p->pt_obs);
    free(p->Hpp); free(p->Hll); free(p->Hpl); free(p->b); free(p->S); free(p->bs);
    if (p->ldlt) orc_ldlt_free(p->ldlt);
    p->chidx = p->phidx = p->colptr = p->rowidx = p->pt_ptr = p->pt_obs = NULL;
    p->Hpp = p->Hll = p->Hpl = p->b = p->S = p->bs = NULL;
    p->ldlt = NULL;
}

void orc_ba_destroy(orc_ba *p) {
    if (!p) return;
    free_structure(p);
    free(p->cam); free(p->pt); free(p->cfix); free(p->pfix); free(p->ocam); free(p->opt); free(p->uv); free(p->info);
    free(p);
}

static void *dup_mem(const void *src, size_t bytes) {
    void *d = malloc(bytes ? bytes : 1);
    if (src) memcpy(d, src, bytes); else memset(d, 0, bytes);
    return d;
}

int orc_ba_set(orc_ba *p, int nc, const double *cams, const unsigned char *cfix, int np, const double *pts,
               const unsigned char *pfix, int no, const int *ocam, const int *opt, const double *uv,
               const double *info, double f, double cx, double cy) {
    for (int k = 0; k < no; ++k)
        if (ocam[k] < 0 || ocam[k] >= nc || opt[k] < 0 || opt[k] >= np) return -1;
    free_structure(p);
    free(p->cam); free(p->pt); free(p->cfix); free(p->pfix); free(p->ocam); free(p->opt); free(p->uv); free(p->info);
    p->nc = nc; p->np = np; p->no = no;
    p->cam = (double *)dup_mem(cams, sizeof(double) * 7 * nc);
    p->pt = (double *)dup_mem(pts, sizeof(double) * 3 * np);
    p->cfix = (unsigned char *)dup_mem(cfix, nc);
    p->pfix = (unsigned char *)dup_mem(pfix, np);
    p->ocam = (int *)dup_mem(ocam, sizeof(int) * no);
    p->opt = (int *)dup_mem(opt, sizeof(int) * no);
    p->uv = (double *)dup_mem(uv, sizeof(double) * 2 * no);
    p->info = (double *)malloc(sizeof(double) * 3 * (no ? no : 1));
    for (int k = 0; k < no; ++k) {
        p->info[3 * k] = info ? info[3 * k] : 1.0;
        p->info[3 * k + 1] = info ? info[3 * k + 1] : 0.0;
        p->info[3 * k + 2] = info ? info[3 * k + 2] : 1.0;
    }
    p->f = f; p->cx = cx; p->cy = cy;
    return 0;
}

void orc_ba_set_robust(orc_ba *p, int kind, double param) { p->robust_kind = kind; p->robust_param = param; }
void orc_ba_set_lm(orc_ba *p, double tau, double user_lambda, int max_trials) {
    if (tau > 0) p->tau = tau;
    p->user_lambda = user_lambda;
    if (max_trials > 0) p->max_trials = max_trials;
}

static int cmp_ll(const void *a, const void *b) {
    const long long x = *(const long long *)a, y = *(const long long *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* initializeOptimization + BlockSolver::buildStructure: free cameras numbered in id order, then
 * the marginalised points; H_schur pattern = diagonal plus every pair of free cameras that
 * co-observe a free point, upper triangle, block-CCS with ascending rows. */
int orc_ba_build_structure(orc_ba *p) {
    free_structure(p);
    p->chidx = (int *)malloc(sizeof(int) * (p->nc ? p->nc : 1));
    p->phidx = (int *)malloc(sizeof(int) * (p->np ? p->np : 1));
    p->ncf = p->npf = 0;
    for (int c = 0; c < p->nc; ++c) p->chidx[c] = p->cfix[c] ? -1 : p->ncf++;
    for (int l = 0; l < p->np; ++l) p->phidx[l] = p->pfix[l] ? -1 : p->npf++;
    p->pt_ptr = (int *)calloc(p->np + 1, sizeof(int));
    p->pt_obs = (int *)malloc(sizeof(int) * (p->no ? p->no : 1));
    for (int k = 0; k < p->no; ++k) p->pt_ptr[p->opt[k] + 1]++;
    for (int l = 0; l < p->np; ++l) p->pt_ptr[l + 1] += p->pt_ptr[l];
    int *fill = (int *)malloc(sizeof(int) * (p->np ? p->np : 1));
    memcpy(fill, p->pt_ptr, sizeof(int) * p->np);
    for (int k = 0; k < p->no; ++k) p->pt_obs[fill[p->opt[k]]++] = k;
    free(fill);
    size_t cap = (size_t)p->ncf;
    for (int l = 0; l < p->np; ++l) {
        const size_t n = (size_t)(p->pt_ptr[l + 1] - p->pt_ptr[l]);
        cap += n * (n + 1) / 2;
    }
    long long *keys = (long long *)malloc(sizeof(long long) * (cap ? cap : 1));
    size_t nk = 0;
    for (int c = 0; c < p->ncf; ++c) keys[nk++] = (long long)c * p->ncf + c;     /* key = col * ncf + row */
    for (int l = 0; l < p->np; ++l) {
        if (p->phidx[l] < 0) continue;
        for (int a = p->pt_ptr[l]; a < p->pt_ptr[l + 1]; ++a)
            for (int b2 = a; b2 < p->pt_ptr[l + 1]; ++b2) {
                int c1 = p->chidx[p->ocam[p->pt_obs[a]]], c2 = p->chidx[p->ocam[p->pt_obs[b2]]];
                if (c1 < 0 || c2 < 0) continue;
                if (c1 > c2) { const int t = c1; c1 = c2; c2 = t; }
                keys[nk++] = (long long)c2 * p->ncf + c1;
            }
    }
    qsort(keys, nk, sizeof(long long), cmp_ll);
    size_t u = 0;
    for (size_t i = 0; i < nk; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) keys[u++] = keys[i];
    p->nb = (int)u;
    p->colptr = (int *)calloc(p->ncf + 1, sizeof(int));
    p->rowidx = (int *)malloc(sizeof(int) * (u ? u : 1));
    for (size_t i = 0; i < u; ++i) {
        const int col = (int)(keys[i] / (p->ncf ? p->ncf : 1)), row = (int)(keys[i] % (p->ncf ? p->ncf : 1));
        p->colptr[col + 1]++;
        p->rowidx[i] = row;
    }
    for (int c = 0; c < p->ncf; ++c) p->colptr[c + 1] += p->colptr[c];
    free(keys);
    p->Hpp = (double *)calloc((size_t)36 * (p->ncf ? p->ncf : 1), sizeof(double));
    p->Hll = (double *)calloc((size_t)9 * (p->npf ? p->npf : 1), sizeof(double));
    p->Hpl = (double *)calloc((size_t)18 * (p->no ? p->no : 1), sizeof(double));
    p->b = (double *)calloc((size_t)6 * p->ncf + 3 * p->npf + 1, sizeof(double));
    p->S = (double *)calloc((size_t)36 * (p->nb ? p->nb : 1), sizeof(double));
    p->bs = (double *)calloc((size_t)6 * p->ncf + 1, sizeof(double));
    p->ldlt = p->ncf > 0 ? orc_ldlt_analyze(p->ncf, 6, p->colptr, p->rowidx) : NULL;
    return p->nb;
}

int orc_ba_num_free_cameras(const orc_ba *p) { return p->ncf; }
int orc_ba_num_free_points(const orc_ba *p) { return p->npf; }
int orc_ba_num_blocks(const orc_ba *p) { return p->nb; }
void orc_ba_get_structure(const orc_ba *p, int *colptr, int *rowidx) {
    memcpy(colptr, p->colptr, sizeof(int) * (p->ncf + 1));
    memcpy(rowidx, p->rowidx, sizeof(int) * p->nb);
}

/* EdgeProjectXYZ2UV::computeError; also returns the camera-frame point */
static void proj_error(const orc_ba *p, int k, double e[2], double xc[3]) {
    const double *T = p->cam + 7 * p->ocam[k], *X = p->pt + 3 * p->opt[k];
    quat_rot(T, X, xc);
    xc[0] += T[4]; xc[1] += T[5]; xc[2] += T[6];
    e[0] = p->uv[2 * k] - (p->f * xc[0] / xc[2] + p->cx);
    e[1] = p->uv[2 * k + 1] - (p->f * xc[1] / xc[2] + p->cy);
}

static double edge_chi2(const orc_ba *p, int k, const double e[2]) {
    const double *o = p->info + 3 * k;
    return o[0] * e[0] * e[0] + 2 * o[1] * e[0] * e[1] + o[2] * e[1] * e[1];
}

double orc_ba_chi2(orc_ba *p) {   /* activeRobustChi2 */
    double sum = 0;
    for (int k = 0; k < p->no; ++k) {
        double e[2], xc[3], rho[3];
        proj_error(p, k, e, xc);
        const double c = edge_chi2(p, k, e);
        if (p->robust_kind != ORC_ROBUST_NONE) { orc_robustify(p->robust_kind, p->robust_param, c, rho); sum += rho[0]; }
        else sum += c;
    }
    return sum;
}

void orc_ba_edge_errors(orc_ba *p, double *err) {
    for (int k = 0; k < p->no; ++k) {
        double xc[3];
        proj_error(p, k, err + 2 * k, xc);
    }
}

/* EdgeProjectXYZ2UV::linearizeOplus: Jp 2x3 (point), Jc 2x6 (camera, [omega, upsilon]) */
void orc_ba_edge_jacobians(const orc_ba *p, int k, double Jp[6], double Jc[12]) {
    double e[2], xc[3], R[9];
    proj_error(p, k, e, xc);
    orc_quat_to_rot(p->cam + 7 * p->ocam[k], R);
    const double x = xc[0], y = xc[1], z = xc[2], z2 = z * z, f = p->f;
    const double tmp[6] = { f, 0, -x / z * f, 0, f, -y / z * f };
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c)
            Jp[r * 3 + c] = -1.0 / z * (tmp[r * 3] * R[c] + tmp[r * 3 + 1] * R[3 + c] + tmp[r * 3 + 2] * R[6 + c]);
    Jc[0] = x * y / z2 * f; Jc[1] = -(1 + (x * x / z2)) * f; Jc[2] = y / z * f;
    Jc[3] = -1.0 / z * f;   Jc[4] = 0;                        Jc[5] = x / z2 * f;
    Jc[6] = (1 + y * y / z2) * f; Jc[7] = -x * y / z2 * f;    Jc[8] = -x / z * f;
    Jc[9] = 0;              Jc[10] = -1.0 / z * f;            Jc[11] = y / z2 * f;
}

/* computeActiveErrors + BlockSolver::buildSystem (constructQuadraticForm per edge, edge order) */
void orc_ba_linearize(orc_ba *p) {
    memset(p->Hpp, 0, sizeof(double) * 36 * p->ncf);
    memset(p->Hll, 0, sizeof(double) * 9 * p->npf);
    memset(p->Hpl, 0, sizeof(double) * 18 * p->no);
    memset(p->b, 0, sizeof(double) * (6 * p->ncf + 3 * p->npf));
    double *bp = p->b, *bl = p->b + 6 * p->ncf;
    for (int k = 0; k < p->no; ++k) {
        double e[2], xc[3], Jp[6], Jc[12], rho[3] = { 0, 1, 0 };
        proj_error(p, k, e, xc);
        orc_ba_edge_jacobians(p, k, Jp, Jc);
        const double *o = p->info + 3 * k;
        if (p->robust_kind != ORC_ROBUST_NONE) orc_robustify(p->robust_kind, p->robust_param, edge_chi2(p, k, e), rho);
        const double O[4] = { rho[1] * o[0], rho[1] * o[1], rho[1] * o[1], rho[1] * o[2] };
        const double Oe[2] = { O[0] * e[0] + O[1] * e[1], O[2] * e[0] + O[3] * e[1] };
        const int ci = p->chidx[p->ocam[k]], li = p->phidx[p->opt[k]];
        double JcO[12], JpO[6];   /* J^T O' stored as [dim][2] */
        for (int a = 0; a < 6; ++a) { JcO[a * 2] = Jc[a] * O[0] + Jc[6 + a] * O[2]; JcO[a * 2 + 1] = Jc[a] * O[1] + Jc[6 + a] * O[3]; }
        for (int a = 0; a < 3; ++a) { JpO[a * 2] = Jp[a] * O[0] + Jp[3 + a] * O[2]; JpO[a * 2 + 1] = Jp[a] * O[1] + Jp[3 + a] * O[3]; }
        if (ci >= 0) {
            for (int a = 0; a < 6; ++a) {
                for (int c = 0; c < 6; ++c) p->Hpp[36 * ci + a * 6 + c] += JcO[a * 2] * Jc[c] + JcO[a * 2 + 1] * Jc[6 + c];
                bp[6 * ci + a] -= Jc[a] * Oe[0] + Jc[6 + a] * Oe[1];
            }
        }
        if (li >= 0) {
            for (int a = 0; a < 3; ++a) {
                for (int c = 0; c < 3; ++c) p->Hll[9 * li + a * 3 + c] += JpO[a * 2] * Jp[c] + JpO[a * 2 + 1] * Jp[3 + c];
                bl[3 * li + a] -= Jp[a] * Oe[0] + Jp[3 + a] * Oe[1];
            }
        }
        if (ci >= 0 && li >= 0)
            for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 3; ++c) p->Hpl[18 * k + a * 3 + c] = JcO[a * 2] * Jp[c] + JcO[a * 2 + 1] * Jp[3 + c];
    }
}

void orc_ba_get_system(const orc_ba *p, double *Hpp, double *Hll, double *Hpl, double *b) {
    if (Hpp) memcpy(Hpp, p->Hpp, sizeof(double) * 36 * p->ncf);
    if (Hll) memcpy(Hll, p->Hll, sizeof(double) * 9 * p->npf);
    if (Hpl) memcpy(Hpl, p->Hpl, sizeof(double) * 18 * p->no);
    if (b) memcpy(b, p->b, sizeof(double) * (6 * p->ncf + 3 * p->npf));
}

double orc_ba_max_diag(const orc_ba *p) {   /* computeLambdaInit: poses and landmarks */
    double m = 0;
    for (int c = 0; c < p->ncf; ++c)
        for (int a = 0; a < 6; ++a) m = fmax(m, fabs(p->Hpp[36 * c + a * 7]));
    for (int l = 0; l < p->npf; ++l)
        for (int a = 0; a < 3; ++a) m = fmax(m, fabs(p->Hll[9 * l + a * 4]));
    return m;
}

static int inv3_sym(const double A[9], double lambda, double Ainv[9]) {
    const double a = A[0] + lambda, b = A[1], c = A[2], d = A[4] + lambda, e = A[5], f = A[8] + lambda;
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(fabs(det) > 0) || !isfinite(det)) return -1;
    const double id = 1.0 / det;
    Ainv[0] = c00 * id; Ainv[1] = c01 * id; Ainv[2] = c02 * id;
    Ainv[3] = Ainv[1];  Ainv[4] = (a * f - c * c) * id; Ainv[5] = (b * c - a * e) * id;
    Ainv[6] = Ainv[2];  Ainv[7] = Ainv[5]; Ainv[8] = (a * d - b * b) * id;
    return 0;
}

static int find_block(const orc_ba *p, int row, int col) {
    int lo = p->colptr[col], hi = p->colptr[col + 1] - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (p->rowidx[mid] == row) return mid;
        if (p->rowidx[mid] < row) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* BlockSolver::solve, Schur branch: S = (Hpp + lambda I) - sum_l Hpl (Hll + lambda I)^-1 Hpl^T,
 * bs = bp - sum_l Hpl (Hll + lambda I)^-1 bl.  Blocks in the CCS order of orc_ba_get_structure. */
int orc_ba_schur(orc_ba *p, double lambda, double *S_out, double *bs_out) {
    memset(p->S, 0, sizeof(double) * 36 * p->nb);
    const double *bp = p->b, *bl = p->b + 6 * p->ncf;
    memcpy(p->bs, bp, sizeof(double) * 6 * p->ncf);
    for (int c = 0; c < p->ncf; ++c) {
        double *D = p->S + 36 * find_block(p, c, c);
        memcpy(D, p->Hpp + 36 * c, sizeof(double) * 36);
        for (int a = 0; a < 6; ++a) D[a * 7] += lambda;
    }
    for (int l = 0; l < p->np; ++l) {
        const int li = p->phidx[l];
        if (li < 0) continue;
        double Dinv[9];
        if (inv3_sym(p->Hll + 9 * li, lambda, Dinv)) return -1;
        double Dinv_b[3];
        for (int a = 0; a < 3; ++a) Dinv_b[a] = Dinv[a * 3] * bl[3 * li] + Dinv[a * 3 + 1] * bl[3 * li + 1] + Dinv[a * 3 + 2] * bl[3 * li + 2];
        for (int ia = p->pt_ptr[l]; ia < p->pt_ptr[l + 1]; ++ia) {
            const int ka = p->pt_obs[ia], ca = p->chidx[p->ocam[ka]];
            if (ca < 0) continue;
            const double *Wa = p->Hpl + 18 * ka;
            double Y[18];   /* Hpl_a Dinv */
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 3; ++c) Y[r * 3 + c] = Wa[r * 3] * Dinv[c] + Wa[r * 3 + 1] * Dinv[3 + c] + Wa[r * 3 + 2] * Dinv[6 + c];
            for (int r = 0; r < 6; ++r) p->bs[6 * ca + r] -= Wa[r * 3] * Dinv_b[0] + Wa[r * 3 + 1] * Dinv_b[1] + Wa[r * 3 + 2] * Dinv_b[2];
            for (int ib = p->pt_ptr[l]; ib < p->pt_ptr[l + 1]; ++ib) {
                const int kb = p->pt_obs[ib], cb = p->chidx[p->ocam[kb]];
                if (cb < 0 || cb < ca) continue;          /* upper triangle; ca == cb handled once per (a,b) pair below */
                if (cb == ca && ib < ia) continue;        /* two observations of one point by the same camera: count each ordered pair once... */
                const double *Wb = p->Hpl + 18 * kb;
                double *Sb = p->S + 36 * find_block(p, ca, cb);
                if (cb == ca && ib != ia) {
                    /* ...and add both Y_a W_b^T and its transpose to the diagonal block */
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 6; ++c) {
                            const double v = Y[r * 3] * Wb[c * 3] + Y[r * 3 + 1] * Wb[c * 3 + 1] + Y[r * 3 + 2] * Wb[c * 3 + 2];
                            Sb[r * 6 + c] -= v;
                            Sb[c * 6 + r] -= v;
                        }
                } else {
                    for (int r = 0; r < 6; ++r)
                        for (int c = 0; c < 6; ++c)
                            Sb[r * 6 + c] -= Y[r * 3] * Wb[c * 3] + Y[r * 3 + 1] * Wb[c * 3 + 1] + Y[r * 3 + 2] * Wb[c * 3 + 2];
                }
            }
        }
    }
    if (S_out) memcpy(S_out, p->S, sizeof(double) * 36 * p->nb);
    if (bs_out) memcpy(bs_out, p->bs, sizeof(double) * 6 * p->ncf);
    return 0;
}

/* solve the damped system; x = [x_cameras (6 ncf), x_points (3 npf)] */
int orc_ba_solve(orc_ba *p, double lambda, double *x) {
    if (orc_ba_schur(p, lambda, NULL, NULL)) return -1;
    double *xp = x, *xl = x + 6 * p->ncf;
    if (p->ncf > 0) {
        if (orc_ldlt_factor(p->ldlt, p->S, 0.0)) return -1;
        orc_ldlt_solve(p->ldlt, p->bs, xp);
    }
    const double *bl = p->b + 6 * p->ncf;
    for (int l = 0; l < p->np; ++l) {
        const int li = p->phidx[l];
        if (li < 0) continue;
        double Dinv[9], r[3] = { bl[3 * li], bl[3 * li + 1], bl[3 * li + 2] };
        if (inv3_sym(p->Hll + 9 * li, lambda, Dinv)) return -1;
        for (int ia = p->pt_ptr[l]; ia < p->pt_ptr[l + 1]; ++ia) {
            const int k = p->pt_obs[ia], c = p->chidx[p->ocam[k]];
            if (c < 0) continue;
            const double *W = p->Hpl + 18 * k;
            for (int a = 0; a < 3; ++a)
                for (int q = 0; q < 6; ++q) r[a] -= W[q * 3 + a] * xp[6 * c + q];
        }
        for (int a = 0; a < 3; ++a) xl[3 * li + a] = Dinv[a * 3] * r[0] + Dinv[a * 3 + 1] * r[1] + Dinv[a * 3 + 2] * r[2];
    }
    return 0;
}

/* oplus: cameras T <- exp(delta) * T, points p += delta */
void orc_ba_update(orc_ba *p, const double *x) {
    const double *xp = x, *xl = x + 6 * p->ncf;
    for (int c = 0; c < p->nc; ++c) {
        const int ci = p->chidx[c];
        if (ci < 0) continue;
        double U[7], T[7];
        orc_se3_exp(xp + 6 * ci, U);
        orc_se3_mul(U, p->cam + 7 * c, T);
        memcpy(p->cam + 7 * c, T, sizeof T);
    }
    for (int l = 0; l < p->np; ++l) {
        const int li = p->phidx[l];
        if (li < 0) continue;
        for (int a = 0; a < 3; ++a) p->pt[3 * l + a] += xl[3 * li + a];
    }
}

void orc_ba_get_cameras(const orc_ba *p, double *cams) { memcpy(cams, p->cam, sizeof(double) * 7 * p->nc); }
void orc_ba_get_points(const orc_ba *p, double *pts) { memcpy(pts, p->pt, sizeof(double) * 3 * p->np); }

/* OptimizationAlgorithmLevenberg::solve inside SparseOptimizer::optimize */
int orc_ba_optimize(orc_ba *p, int max_iter, double stop_rel_gain, double *hist, int hist_cap, double *final_chi2,
                    double *final_lambda) {
    if (!p->Hpp) orc_ba_build_structure(p);      /* optimize() without initializeOptimization(): build it here */
    const int n = 6 * p->ncf + 3 * p->npf;
    double *x = (double *)calloc(n + 1, sizeof(double));
    double *cam_bk = (double *)malloc(sizeof(double) * 7 * (p->nc ? p->nc : 1));
    double *pt_bk = (double *)malloc(sizeof(double) * 3 * (p->np ? p->np : 1));
    double lambda = 0, ni = 2, currentChi = 0;
    int done = 0;
    for (int it = 0; it < max_iter; ++it) {
        if (it == 0) currentChi = orc_ba_chi2(p);
        orc_ba_linearize(p);
        if (it == 0) {
            lambda = p->user_lambda > 0 ? p->user_lambda : p->tau * orc_ba_max_diag(p);
            ni = 2;
        }
        const double chi_start = currentChi;
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(cam_bk, p->cam, sizeof(double) * 7 * p->nc);
            memcpy(pt_bk, p->pt, sizeof(double) * 3 * p->np);
            const int ok = orc_ba_solve(p, lambda, x) == 0;
            double tempChi;
            if (ok) { orc_ba_update(p, x); tempChi = orc_ba_chi2(p); }
            else tempChi = 1.7976931348623157e308;
            double scale = 0;
            for (int j = 0; j < n; ++j) scale += x[j] * (lambda * x[j] + p->b[j]);
            scale += 1e-3;
            rho = (currentChi - tempChi) / scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(p->cam, cam_bk, sizeof(double) * 7 * p->nc);
                memcpy(p->pt, pt_bk, sizeof(double) * 3 * p->np);
            }
            qmax++;
        } while (rho < 0 && qmax < p->max_trials);
        done = it + 1;
        if (hist && it < hist_cap) { hist[it * 4] = currentChi; hist[it * 4 + 1] = lambda; hist[it * 4 + 2] = qmax; hist[it * 4 + 3] = rho; }
        if (qmax == p->max_trials || rho == 0) break;
        if (stop_rel_gain > 0) {
            const double gain = (chi_start - currentChi) / currentChi;
            if (gain >= 0 && gain < stop_rel_gain) break;
        }
    }
    free(x); free(cam_bk); free(pt_bk);
    if (final_chi2) *final_chi2 = currentChi;
    if (final_lambda) *final_lambda = lambda;
    return done;
}
