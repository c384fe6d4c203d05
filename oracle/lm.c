/*
 * lm.c -- graph container, block structure, linearisation and the g2o
 * Levenberg-Marquardt loop of the oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Restates, step by step:
 *   SparseOptimizer::initializeOptimization / BlockSolver::buildStructure (row a14)
 *   BaseBinaryEdge::linearizeOplus (numeric, row a11) and an analytic variant
 *   BaseBinaryEdge::constructQuadraticForm (row a12), RobustKernelHuber (row a13)
 *   BlockSolver::buildSystem / setLambda / solve / restoreDiagonal (row a16)
 *   OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale (row a15)
 *   SparseOptimizer::optimize
 * for the graphs the reference builds at kitti_surf.cpp:592-675 (7-DoF Sim3) and
 * :767-886 (4-DoF scale+translation, 1-DoF scale).
 *
 * [EXT vio_g2o] G2oEdgeScale / G2oEdgeScaleTrans sources are not available (SURVEY.md
 * row a18).  The model used here is the one the reference's own linear formulations
 * spell out: scale rows  s_ji*s_i - s_j = 0  (kitti_surf.cpp:897-906) and translation
 * rows  t_j - (s_j/s_i) R_j R_i^T t_i = t_ji  (kitti_surf.cpp:969-985); additive oplus.
 */
#include "oracle.h"
#include "ldlt.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <time.h>

#define MAXD 7

/* worker threads of the per-edge loops (orc_set_threads); 1 = the reference's behaviour (single thread,
 * CMakeLists.txt has no OpenMP flag and build.sh:49 does not enable g2o's) */
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n > 1 ? n : 1; }
int orc_get_threads(void) { return g_threads; }

/* static-schedule parallel for over [0,n) on g_threads POSIX threads (libgomp is not in this image) */
#include <pthread.h>
typedef void (*range_fn)(void *ctx, int lo, int hi);
typedef struct { range_fn fn; void *ctx; int lo, hi; } par_task;
static void *par_run(void *arg) { par_task *t = (par_task *)arg; t->fn(t->ctx, t->lo, t->hi); return 0; }
static void par_for(int n, range_fn fn, void *ctx) {
    int nt = g_threads;
    if (nt > n) nt = n;
    if (nt <= 1) { fn(ctx, 0, n); return; }
    pthread_t th[256];
    par_task task[256];
    if (nt > 256) nt = 256;
    for (int t = 0; t < nt; ++t) {
        task[t].fn = fn; task[t].ctx = ctx;
        task[t].lo = (int)((long long)n * t / nt); task[t].hi = (int)((long long)n * (t + 1) / nt);
        if (t + 1 == nt || pthread_create(&th[t], 0, par_run, &task[t]) != 0) { par_run(&task[t]); th[t] = 0; }
    }
    for (int t = 0; t + 1 < nt; ++t) if (th[t]) pthread_join(th[t], 0);
}

struct orc_problem {
    int kind, d, est_dim;
    int nv, ne;
    double *est, *aux;
    unsigned char *fixed;
    int *ev0, *ev1;
    double *meas, *info;
    int robust_kind;
    double robust_param;
    int jac_mode;
    double jac_h;
    int scale_model;     /* 0 difference / additive, 1 log-ratio / multiplicative (row a18: both readings) */
    double tau, user_lambda;
    int max_trials;
    /* structure */
    int built;
    int nfree, nblocks;
    int *hidx, *free2v;
    int *colptr, *rowidx;
    int *e_slot;
    int *diag_slot;
    double *H, *b, *x, *backup;
    orc_ldlt *ldlt;
    double timing[4];
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

orc_problem *orc_create(int kind) {
    orc_problem *p = (orc_problem *)calloc(1, sizeof(orc_problem));
    p->kind = kind;
    switch (kind) {
    case ORC_KIND_SIM3: p->d = 7; p->est_dim = 8; break;
    case ORC_KIND_SCALE_TRANS: p->d = 4; p->est_dim = 4; break;
    case ORC_KIND_SCALE: p->d = 1; p->est_dim = 1; break;
    default: free(p); return 0;
    }
    p->jac_mode = ORC_JAC_NUMERIC;
    p->jac_h = 1e-9;          /* g2o BaseBinaryEdge::linearizeOplus delta */
    p->tau = 1e-5;            /* g2o OptimizationAlgorithmLevenberg _tau */
    p->user_lambda = 0;
    p->max_trials = 10;       /* g2o _maxTrialsAfterFailure */
    return p;
}

static void free_structure(orc_problem *p) {
    free(p->hidx); free(p->free2v); free(p->colptr); free(p->rowidx); free(p->e_slot);
    free(p->diag_slot); free(p->H); free(p->b); free(p->x); free(p->backup);
    orc_ldlt_free(p->ldlt);
    p->hidx = p->free2v = p->colptr = p->rowidx = p->e_slot = p->diag_slot = 0;
    p->H = p->b = p->x = p->backup = 0;
    p->ldlt = 0;
    p->built = 0;
}

void orc_destroy(orc_problem *p) {
    if (!p) return;
    free_structure(p);
    free(p->est); free(p->aux); free(p->fixed); free(p->ev0); free(p->ev1); free(p->meas); free(p->info);
    free(p);
}

static void *dup_mem(const void *src, size_t bytes) {
    void *m = malloc(bytes > 0 ? bytes : 1);
    if (src && bytes) memcpy(m, src, bytes);
    return m;
}

int orc_set_vertices(orc_problem *p, int n, const double *est, const unsigned char *fixed, const double *aux) {
    free_structure(p);
    free(p->est); free(p->aux); free(p->fixed);
    p->nv = n;
    p->est = (double *)dup_mem(est, sizeof(double) * n * p->est_dim);
    p->fixed = (unsigned char *)calloc(n > 0 ? n : 1, 1);
    if (fixed) memcpy(p->fixed, fixed, n);
    p->aux = aux ? (double *)dup_mem(aux, sizeof(double) * n * 4) : 0;
    if (p->kind == ORC_KIND_SCALE_TRANS && !aux) return -1;
    return 0;
}

int orc_set_edges(orc_problem *p, int n, const int *v0, const int *v1, const double *meas, const double *info) {
    free_structure(p);
    free(p->ev0); free(p->ev1); free(p->meas); free(p->info);
    for (int k = 0; k < n; ++k)
        if (v0[k] < 0 || v0[k] >= p->nv || v1[k] < 0 || v1[k] >= p->nv || v0[k] == v1[k]) {
            p->ev0 = p->ev1 = 0; p->meas = p->info = 0; p->ne = 0;
            return -1;
        }
    p->ne = n;
    p->ev0 = (int *)dup_mem(v0, sizeof(int) * n);
    p->ev1 = (int *)dup_mem(v1, sizeof(int) * n);
    p->meas = (double *)dup_mem(meas, sizeof(double) * n * p->est_dim);
    p->info = info ? (double *)dup_mem(info, sizeof(double) * n * p->d * p->d) : 0;
    return 0;
}

void orc_set_scale_model(orc_problem *p, int model) { p->scale_model = model ? 1 : 0; }
void orc_set_robust(orc_problem *p, int kind, double param) { p->robust_kind = kind; p->robust_param = param; }
void orc_set_jacobian_mode(orc_problem *p, int mode, double h) { p->jac_mode = mode; if (h > 0) p->jac_h = h; }
void orc_set_lm(orc_problem *p, double tau, double user_lambda_init, int max_trials) {
    if (tau > 0) p->tau = tau;
    p->user_lambda = user_lambda_init;
    if (max_trials > 0) p->max_trials = max_trials;
}

/* ---- structure (row a14) -------------------------------------------------- */
static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

int orc_build_structure(orc_problem *p) {
    free_structure(p);
    const int nv = p->nv, ne = p->ne, d = p->d;
    p->hidx = (int *)malloc(sizeof(int) * (nv > 0 ? nv : 1));
    p->free2v = (int *)malloc(sizeof(int) * (nv > 0 ? nv : 1));
    int nf = 0;
    for (int v = 0; v < nv; ++v) {
        if (p->fixed[v]) p->hidx[v] = -1;
        else { p->hidx[v] = nf; p->free2v[nf++] = v; }
    }
    p->nfree = nf;
    /* per column c: rows {c} U {min(hi,hj) : max(hi,hj)=c} */
    int *cnt = (int *)calloc(nf + 1, sizeof(int));
    for (int c = 0; c < nf; ++c) cnt[c] = 1;
    for (int k = 0; k < ne; ++k) {
        int hi = p->hidx[p->ev0[k]], hj = p->hidx[p->ev1[k]];
        if (hi < 0 || hj < 0) continue;
        cnt[hi > hj ? hi : hj]++;
    }
    int *start = (int *)malloc(sizeof(int) * (nf + 1));
    start[0] = 0;
    for (int c = 0; c < nf; ++c) start[c + 1] = start[c] + cnt[c];
    int *rows = (int *)malloc(sizeof(int) * (start[nf] > 0 ? start[nf] : 1));
    memset(cnt, 0, sizeof(int) * (nf + 1));
    for (int c = 0; c < nf; ++c) rows[start[c] + cnt[c]++] = c;
    for (int k = 0; k < ne; ++k) {
        int hi = p->hidx[p->ev0[k]], hj = p->hidx[p->ev1[k]];
        if (hi < 0 || hj < 0) continue;
        int r = hi < hj ? hi : hj, c = hi < hj ? hj : hi;
        rows[start[c] + cnt[c]++] = r;
    }
    p->colptr = (int *)malloc(sizeof(int) * (nf + 1));
    p->colptr[0] = 0;
    int total = 0;
    for (int c = 0; c < nf; ++c) {
        int *seg = rows + start[c];
        int n = cnt[c];
        qsort(seg, n, sizeof(int), cmp_int);
        int w = 0;
        for (int i = 0; i < n; ++i)
            if (i == 0 || seg[i] != seg[i - 1]) seg[w++] = seg[i];
        cnt[c] = w;
        total += w;
        p->colptr[c + 1] = total;
    }
    p->nblocks = total;
    p->rowidx = (int *)malloc(sizeof(int) * (total > 0 ? total : 1));
    p->diag_slot = (int *)malloc(sizeof(int) * (nf > 0 ? nf : 1));
    for (int c = 0; c < nf; ++c) {
        memcpy(p->rowidx + p->colptr[c], rows + start[c], sizeof(int) * cnt[c]);
        p->diag_slot[c] = p->colptr[c + 1] - 1; /* rows ascending and r<=c: the diagonal is last */
    }
    p->e_slot = (int *)malloc(sizeof(int) * (ne > 0 ? ne : 1));
    for (int k = 0; k < ne; ++k) {
        int hi = p->hidx[p->ev0[k]], hj = p->hidx[p->ev1[k]];
        p->e_slot[k] = -1;
        if (hi < 0 || hj < 0) continue;
        int r = hi < hj ? hi : hj, c = hi < hj ? hj : hi;
        int lo = p->colptr[c], hi2 = p->colptr[c + 1] - 1;
        while (lo < hi2) { int mid = (lo + hi2) / 2; if (p->rowidx[mid] < r) lo = mid + 1; else hi2 = mid; }
        p->e_slot[k] = lo;
    }
    free(cnt); free(start); free(rows);
    p->H = (double *)calloc((size_t)(total > 0 ? total : 1) * d * d, sizeof(double));
    p->b = (double *)calloc((size_t)(nf > 0 ? nf : 1) * d, sizeof(double));
    p->x = (double *)calloc((size_t)(nf > 0 ? nf : 1) * d, sizeof(double));
    p->backup = (double *)malloc(sizeof(double) * (nv > 0 ? nv : 1) * p->est_dim);
    p->built = 1;
    return total;
}

int orc_num_free(const orc_problem *p) { return p->nfree; }
int orc_num_blocks(const orc_problem *p) { return p->nblocks; }
int orc_dim(const orc_problem *p) { return p->d; }
void orc_get_structure(const orc_problem *p, int *colptr, int *rowidx) {
    memcpy(colptr, p->colptr, sizeof(int) * (p->nfree + 1));
    memcpy(rowidx, p->rowidx, sizeof(int) * p->nblocks);
}
void orc_get_hessian_index(const orc_problem *p, int *hidx) { memcpy(hidx, p->hidx, sizeof(int) * p->nv); }

/* ---- per-kind edge model --------------------------------------------------- */
static void quat_rot(const double q[4], const double v[3], double out[3]) {
    double R[9];
    orc_quat_to_rot(q, R);
    for (int i = 0; i < 3; ++i) out[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}

static void edge_error(const orc_problem *p, int k, const double *xi, const double *xj, double *e) {
    const double *m = p->meas + (size_t)k * p->est_dim;
    switch (p->kind) {
    case ORC_KIND_SIM3:
        orc_sim3_edge_error(m, xi, xj, e);
        break;
    case ORC_KIND_SCALE_TRANS: {
        const double *qi = p->aux + 4 * p->ev0[k], *qj = p->aux + 4 * p->ev1[k];
        const double qic[4] = { -qi[0], -qi[1], -qi[2], qi[3] };
        double a[3], b[3];
        quat_rot(qic, xi + 1, a);      /* R_i^T t_i */
        quat_rot(qj, a, b);            /* R_j R_i^T t_i */
        const double sr = xj[0] / xi[0];
        e[0] = p->scale_model ? log(m[0] * xi[0] / xj[0]) : m[0] * xi[0] - xj[0];
        for (int c = 0; c < 3; ++c) e[1 + c] = xj[1 + c] - sr * b[c] - m[1 + c];
        break;
    }
    case ORC_KIND_SCALE:
        e[0] = p->scale_model ? log(m[0] * xi[0] / xj[0]) : m[0] * xi[0] - xj[0];
        break;
    }
}

static void oplus(const orc_problem *p, double *x, const double *delta) {
    if (p->kind == ORC_KIND_SIM3) { /* VertexSim3Expmap::oplusImpl: S <- Sim3(delta) * S (row a9) */
        double U[8], R[8];
        orc_sim3_exp(delta, U);
        orc_sim3_mul(U, x, R);
        memcpy(x, R, sizeof R);
    } else {
        for (int c = 0; c < p->d; ++c) x[c] += delta[c];
        if (p->scale_model) x[0] = (x[0] - delta[0]) * exp(delta[0]);   /* s <- s exp(d_sigma) */
    }
}

static void edge_jacobians(const orc_problem *p, int k, const double *xi, const double *xj,
                           int need_i, int need_j, double *Ji, double *Jj) {
    const int d = p->d, ed = p->est_dim;
    if (p->jac_mode == ORC_JAC_NUMERIC) {
        const double h = p->jac_h, scalar = 1.0 / (2 * h);
        for (int side = 0; side < 2; ++side) {
            if ((side == 0 && !need_i) || (side == 1 && !need_j)) continue;
            double *J = side == 0 ? Ji : Jj;
            for (int c = 0; c < d; ++c) {
                double add[MAXD] = { 0 }, xp[8], e1[MAXD], e2[MAXD];
                add[c] = h;
                memcpy(xp, side == 0 ? xi : xj, sizeof(double) * ed);
                oplus(p, xp, add);
                edge_error(p, k, side == 0 ? xp : xi, side == 0 ? xj : xp, e1);
                add[c] = -h;
                memcpy(xp, side == 0 ? xi : xj, sizeof(double) * ed);
                oplus(p, xp, add);
                edge_error(p, k, side == 0 ? xp : xi, side == 0 ? xj : xp, e2);
                for (int r = 0; r < d; ++r) J[r * d + c] = scalar * (e1[r] - e2[r]);
            }
        }
        return;
    }
    switch (p->kind) {
    case ORC_KIND_SIM3:
        orc_sim3_edge_jac_analytic(p->meas + (size_t)k * 8, xi, xj, Ji, Jj);
        break;
    case ORC_KIND_SCALE_TRANS: {
        const double *m = p->meas + (size_t)k * 4;
        const double *qi = p->aux + 4 * p->ev0[k], *qj = p->aux + 4 * p->ev1[k];
        const double qic[4] = { -qi[0], -qi[1], -qi[2], qi[3] };
        double Ri[9], Rj[9], Q[9], a[3], b[3];
        orc_quat_to_rot(qi, Ri);
        orc_quat_to_rot(qj, Rj);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double acc = 0;
                for (int t = 0; t < 3; ++t) acc += Rj[r * 3 + t] * Ri[c * 3 + t];
                Q[r * 3 + c] = acc;
            }
        quat_rot(qic, xi + 1, a);
        quat_rot(qj, a, b);
        const double si = xi[0], sj = xj[0];
        memset(Ji, 0, sizeof(double) * 16);
        memset(Jj, 0, sizeof(double) * 16);
        Ji[0] = p->scale_model ? 1.0 : m[0];
        Jj[0] = -1;
        for (int r = 0; r < 3; ++r) {
            Ji[(1 + r) * 4] = p->scale_model ? sj / si * b[r] : sj / (si * si) * b[r];
            Jj[(1 + r) * 4] = p->scale_model ? -(sj / si) * b[r] : -b[r] / si;
            for (int c = 0; c < 3; ++c) Ji[(1 + r) * 4 + 1 + c] = -(sj / si) * Q[r * 3 + c];
            Jj[(1 + r) * 4 + 1 + r] = 1;
        }
        break;
    }
    case ORC_KIND_SCALE:
        Ji[0] = p->scale_model ? 1.0 : p->meas[k];
        Jj[0] = -1;
        break;
    }
}

static double edge_chi2(const orc_problem *p, int k, const double *e) {
    const int d = p->d;
    if (!p->info) { double s = 0; for (int i = 0; i < d; ++i) s += e[i] * e[i]; return s; }
    const double *O = p->info + (size_t)k * d * d;
    double s = 0;
    for (int i = 0; i < d; ++i) {
        double acc = 0;
        for (int j = 0; j < d; ++j) acc += O[i * d + j] * e[j];
        s += e[i] * acc;
    }
    return s;
}

void orc_edge_errors(orc_problem *p, double *err) {
    for (int k = 0; k < p->ne; ++k)
        edge_error(p, k, p->est + (size_t)p->ev0[k] * p->est_dim, p->est + (size_t)p->ev1[k] * p->est_dim, err + (size_t)k * p->d);
}

static double edge_rho0(const orc_problem *p, int k) {
    double e[MAXD];
    edge_error(p, k, p->est + (size_t)p->ev0[k] * p->est_dim, p->est + (size_t)p->ev1[k] * p->est_dim, e);
    double c = edge_chi2(p, k, e);
    if (p->robust_kind != ORC_ROBUST_NONE) {
        double rho[3];
        orc_robustify(p->robust_kind, p->robust_param, c, rho);
        c = rho[0];
    }
    return c;
}

typedef struct { const orc_problem *p; double *c; } chi_ctx;
static void chi_range(void *ctx, int lo, int hi) {
    chi_ctx *c = (chi_ctx *)ctx;
    for (int k = lo; k < hi; ++k) c->c[k] = edge_rho0(c->p, k);
}

double orc_chi2(orc_problem *p) { /* computeActiveErrors + activeRobustChi2 */
    double total = 0;
    if (g_threads <= 1) {
        for (int k = 0; k < p->ne; ++k) total += edge_rho0(p, k);
        return total;
    }
    /* threads: per-edge values in parallel, summed serially in edge order (same bits as 1 thread) */
    double *c = (double *)malloc(sizeof(double) * (p->ne > 0 ? p->ne : 1));
    chi_ctx cx = { p, c };
    par_for(p->ne, chi_range, &cx);
    for (int k = 0; k < p->ne; ++k) total += c[k];
    free(c);
    return total;
}

/* ---- buildSystem (rows a11, a12, a16) -------------------------------------- */
/* One edge's linearizeOplus + constructQuadraticForm products (no accumulation):
 * c = [Hii dd | Hij dd | Hjj dd | bi d | bj d].  Hij is A^T O' B as g2o forms it (row side = vertex(0)). */
#define CONTRIB_STRIDE (3 * MAXD * MAXD + 2 * MAXD)
static void edge_contrib(const orc_problem *p, int k, double *c) {
    const int d = p->d, dd = d * d, ed = p->est_dim;
    const int vi = p->ev0[k], vj = p->ev1[k];
    const int hi = p->hidx[vi], hj = p->hidx[vj];
    if (hi < 0 && hj < 0) return;
    const double *xi = p->est + (size_t)vi * ed, *xj = p->est + (size_t)vj * ed;
    double e[MAXD], A[MAXD * MAXD], B[MAXD * MAXD], Oe[MAXD], AtO[MAXD * MAXD], BtO[MAXD * MAXD];
    double *Hii = c, *Hij = c + dd, *Hjj = c + 2 * dd, *bi = c + 3 * dd, *bj = c + 3 * dd + d;
    edge_error(p, k, xi, xj, e);
    edge_jacobians(p, k, xi, xj, hi >= 0, hj >= 0, A, B);
    const double *O = p->info ? p->info + (size_t)k * dd : 0;
    double w = 1.0;
    if (p->robust_kind != ORC_ROBUST_NONE) {
        double rho[3];
        orc_robustify(p->robust_kind, p->robust_param, edge_chi2(p, k, e), rho);
        w = rho[1];
    }
    for (int i = 0; i < d; ++i) { /* omega_r = -rho1 * Omega e */
        double acc = 0;
        if (O) for (int j = 0; j < d; ++j) acc += O[i * d + j] * e[j]; else acc = e[i];
        Oe[i] = -w * acc;
    }
    if (hi >= 0) {
        for (int r = 0; r < d; ++r)
            for (int cc = 0; cc < d; ++cc) {
                double acc = 0;
                if (O) for (int t = 0; t < d; ++t) acc += A[t * d + r] * O[t * d + cc]; else acc = A[cc * d + r];
                AtO[r * d + cc] = w * acc;
            }
        for (int r = 0; r < d; ++r) { double acc = 0; for (int t = 0; t < d; ++t) acc += A[t * d + r] * Oe[t]; bi[r] = acc; }
        for (int r = 0; r < d; ++r)
            for (int cc = 0; cc < d; ++cc) { double acc = 0; for (int t = 0; t < d; ++t) acc += AtO[r * d + t] * A[t * d + cc]; Hii[r * d + cc] = acc; }
        if (hj >= 0)
            for (int r = 0; r < d; ++r)
                for (int cc = 0; cc < d; ++cc) { double acc = 0; for (int t = 0; t < d; ++t) acc += AtO[r * d + t] * B[t * d + cc]; Hij[r * d + cc] = acc; }
    }
    if (hj >= 0) {
        for (int r = 0; r < d; ++r)
            for (int cc = 0; cc < d; ++cc) {
                double acc = 0;
                if (O) for (int t = 0; t < d; ++t) acc += B[t * d + r] * O[t * d + cc]; else acc = B[cc * d + r];
                BtO[r * d + cc] = w * acc;
            }
        for (int r = 0; r < d; ++r) { double acc = 0; for (int t = 0; t < d; ++t) acc += B[t * d + r] * Oe[t]; bj[r] = acc; }
        for (int r = 0; r < d; ++r)
            for (int cc = 0; cc < d; ++cc) { double acc = 0; for (int t = 0; t < d; ++t) acc += BtO[r * d + t] * B[t * d + cc]; Hjj[r * d + cc] = acc; }
    }
}

typedef struct { const orc_problem *p; double *buf; int k0; } lin_ctx;
static void lin_range(void *ctx, int lo, int hi) {
    lin_ctx *c = (lin_ctx *)ctx;
    for (int k = lo; k < hi; ++k) edge_contrib(c->p, c->k0 + k, c->buf + (size_t)k * CONTRIB_STRIDE);
}

/* The per-edge products are independent (g2o's optional OpenMP build parallelises exactly this loop);
 * the accumulation into H and b stays serial and in edge order, so the sums are the same bits for any
 * thread count. */
void orc_linearize(orc_problem *p) {
    if (!p->built) orc_build_structure(p);
    const int d = p->d, dd = d * d;
    memset(p->H, 0, sizeof(double) * (size_t)p->nblocks * dd);
    memset(p->b, 0, sizeof(double) * (size_t)p->nfree * d);
    enum { CHUNK = 16384 };
    double *buf = (double *)malloc(sizeof(double) * CHUNK * CONTRIB_STRIDE);
    for (int k0 = 0; k0 < p->ne; k0 += CHUNK) {
        const int k1 = k0 + CHUNK < p->ne ? k0 + CHUNK : p->ne;
        lin_ctx lc = { p, buf, k0 };
        par_for(k1 - k0, lin_range, &lc);
        for (int k = k0; k < k1; ++k) {
            const int hi = p->hidx[p->ev0[k]], hj = p->hidx[p->ev1[k]];
            if (hi < 0 && hj < 0) continue;
            const double *c = buf + (size_t)(k - k0) * CONTRIB_STRIDE;
            if (hi >= 0) {
                double *bi = p->b + (size_t)hi * d, *Hii = p->H + (size_t)p->diag_slot[hi] * dd;
                for (int r = 0; r < d; ++r) bi[r] += c[3 * dd + r];
                for (int t = 0; t < dd; ++t) Hii[t] += c[t];
                if (hj >= 0) {
                    double *Hij = p->H + (size_t)p->e_slot[k] * dd;
                    if (hi < hj) { for (int t = 0; t < dd; ++t) Hij[t] += c[dd + t]; }
                    else { /* _hessianRowMajor: the stored block is (hj,hi) = (A^T O B)^T */
                        for (int r = 0; r < d; ++r)
                            for (int cc = 0; cc < d; ++cc) Hij[cc * d + r] += c[dd + r * d + cc];
                    }
                }
            }
            if (hj >= 0) {
                double *bj = p->b + (size_t)hj * d, *Hjj = p->H + (size_t)p->diag_slot[hj] * dd;
                for (int r = 0; r < d; ++r) bj[r] += c[3 * dd + d + r];
                for (int t = 0; t < dd; ++t) Hjj[t] += c[2 * dd + t];
            }
        }
    }
    free(buf);
}

void orc_get_H(const orc_problem *p, double *blocks) { memcpy(blocks, p->H, sizeof(double) * (size_t)p->nblocks * p->d * p->d); }
void orc_get_b(const orc_problem *p, double *b) { memcpy(b, p->b, sizeof(double) * (size_t)p->nfree * p->d); }

double orc_max_diag(const orc_problem *p) { /* computeLambdaInit's max |H_jj| */
    const int d = p->d, dd = d * d;
    double mx = 0;
    for (int c = 0; c < p->nfree; ++c) {
        const double *Hd = p->H + (size_t)p->diag_slot[c] * dd;
        for (int j = 0; j < d; ++j) if (fabs(Hd[j * d + j]) > mx) mx = fabs(Hd[j * d + j]);
    }
    return mx;
}

int orc_solve(orc_problem *p, double lambda, double *x) {
    if (!p->ldlt) p->ldlt = orc_ldlt_analyze(p->nfree, p->d, p->colptr, p->rowidx);
    if (orc_ldlt_factor(p->ldlt, p->H, lambda) != 0) return -1;
    orc_ldlt_solve(p->ldlt, p->b, p->x);
    if (x && x != p->x) memcpy(x, p->x, sizeof(double) * (size_t)p->nfree * p->d);
    return 0;
}

void orc_update(orc_problem *p, const double *x) {
    for (int f = 0; f < p->nfree; ++f)
        oplus(p, p->est + (size_t)p->free2v[f] * p->est_dim, x + (size_t)f * p->d);
}

void orc_get_vertices(const orc_problem *p, double *est) { memcpy(est, p->est, sizeof(double) * (size_t)p->nv * p->est_dim); }
void orc_get_timing(const orc_problem *p, double t[4]) { memcpy(t, p->timing, sizeof p->timing); }

/* ---- OptimizationAlgorithmLevenberg::solve inside SparseOptimizer::optimize -- */
int orc_optimize(orc_problem *p, int max_iter, double stop_rel_gain, double *hist, int hist_cap,
                 double *final_chi2, double *final_lambda) {
    if (!p->built) orc_build_structure(p);
    if (p->nfree == 0) return -1;
    const size_t nx = (size_t)p->nfree * p->d;
    const size_t est_bytes = sizeof(double) * (size_t)p->nv * p->est_dim;
    double lambda = 0, ni = 2, chi_last = 0;
    int done = 0;
    memset(p->timing, 0, sizeof p->timing);
    for (int it = 0; it < max_iter; ++it) {
        double t0 = now_s();
        double currentChi = orc_chi2(p);
        double t1 = now_s();
        orc_linearize(p);
        double t2 = now_s();
        p->timing[2] += t1 - t0;
        p->timing[0] += t2 - t1;
        if (it == 0) {
            lambda = p->user_lambda > 0 ? p->user_lambda : p->tau * orc_max_diag(p);
            ni = 2;
        }
        const double chi_start = currentChi;
        double rho = 0, tempChi = currentChi;
        int qmax = 0;
        do {
            memcpy(p->backup, p->est, est_bytes);                 /* push */
            double t3 = now_s();
            int ok2 = orc_solve(p, lambda, 0) == 0;               /* setLambda + solve + restoreDiagonal */
            double t4 = now_s();
            orc_update(p, p->x);
            tempChi = orc_chi2(p);
            double t5 = now_s();
            p->timing[1] += t4 - t3;
            p->timing[2] += t5 - t4;
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (size_t j = 0; j < nx; ++j) scale += p->x[j] * (lambda * p->x[j] + p->b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = alpha < 2. / 3. ? alpha : 2. / 3.;
                double scaleFactor = alpha > 1. / 3. ? alpha : 1. / 3.;
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;                             /* discardTop */
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(p->est, p->backup, est_bytes);             /* pop */
            }
            qmax++;
        } while (rho < 0 && qmax < p->max_trials);
        done = it + 1;
        chi_last = currentChi;
        if (hist && it < hist_cap) {
            hist[it * 4 + 0] = currentChi; hist[it * 4 + 1] = lambda; hist[it * 4 + 2] = qmax; hist[it * 4 + 3] = rho;
        }
        if (qmax == p->max_trials || rho == 0) break;             /* Terminate */
        if (stop_rel_gain > 0) {
            double gain = (chi_start - currentChi) / currentChi;
            if (gain >= 0 && gain < stop_rel_gain) break;
        }
    }
    p->timing[3] = p->timing[0] + p->timing[1] + p->timing[2];
    if (final_chi2) *final_chi2 = chi_last;
    if (final_lambda) *final_lambda = lambda;
    return done;
}
