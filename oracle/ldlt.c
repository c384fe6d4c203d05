/*
 * ldlt.c -- sparse LDL^T for the oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Plays the role of g2o::LinearSolverEigen (Eigen::SimplicialLDLT with a
 * fill-reducing block ordering) that every optimiser of the reference plugs in
 * (kitti_surf.cpp:553-557, :728-732, bal_example.cpp:73-83): symbolic analysis
 * once, numeric up-looking LDL^T per LM trial, failure on an exact zero pivot.
 * The algorithm is the classic elimination-tree up-looking factorisation that
 * SimplicialLDLT itself implements; the ordering is a plain minimum-degree on
 * the block graph (the choice of ordering only moves round-off).
 */
#include "ldlt.h"
#include <stdlib.h>
#include <string.h>

struct orc_ldlt {
    int nb, d, n;       /* blocks, block dim, scalar dim */
    int *perm;          /* new block position -> old block index */
    int *iperm;         /* old block index -> new position */
    int *Ap, *Ai;       /* permuted scalar upper CSC pattern */
    double *Ax;
    int *map;           /* per stored input scalar: slot in Ax or -1 */
    int nblocks;
    int *parent, *Lp, *Li, *Lnz, *flag, *pattern;
    double *Lx, *D, *Y;
    int *diag_slot;     /* slot of each scalar diagonal in Ax */
};

/* ---- minimum degree on the block graph ---------------------------------- */
typedef struct { int *v; int n, cap; } ivec;
static void iv_push(ivec *a, int x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 8; a->v = (int *)realloc(a->v, sizeof(int) * a->cap); }
    a->v[a->n++] = x;
}

typedef struct { int deg, v; } hent;
typedef struct { hent *h; int n, cap; } heap;
static void heap_push(heap *H, int deg, int v) {
    if (H->n == H->cap) { H->cap = H->cap ? H->cap * 2 : 64; H->h = (hent *)realloc(H->h, sizeof(hent) * H->cap); }
    int i = H->n++;
    H->h[i].deg = deg; H->h[i].v = v;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (H->h[p].deg < H->h[i].deg || (H->h[p].deg == H->h[i].deg && H->h[p].v < H->h[i].v)) break;
        hent t = H->h[p]; H->h[p] = H->h[i]; H->h[i] = t; i = p;
    }
}
static hent heap_pop(heap *H) {
    hent top = H->h[0];
    H->h[0] = H->h[--H->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < H->n && (H->h[l].deg < H->h[m].deg || (H->h[l].deg == H->h[m].deg && H->h[l].v < H->h[m].v))) m = l;
        if (r < H->n && (H->h[r].deg < H->h[m].deg || (H->h[r].deg == H->h[m].deg && H->h[r].v < H->h[m].v))) m = r;
        if (m == i) break;
        hent t = H->h[m]; H->h[m] = H->h[i]; H->h[i] = t; i = m;
    }
    return top;
}

static void min_degree_order(int nb, const int *colptr, const int *rowidx, int *perm) {
    ivec *adj = (ivec *)calloc(nb, sizeof(ivec));
    for (int c = 0; c < nb; ++c)
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
            int r = rowidx[p];
            if (r != c) { iv_push(&adj[r], c); iv_push(&adj[c], r); }
        }
    char *dead = (char *)calloc(nb, 1);
    int *mark = (int *)malloc(sizeof(int) * nb);
    for (int i = 0; i < nb; ++i) mark[i] = -1;
    heap H = { 0, 0, 0 };
    for (int v = 0; v < nb; ++v) heap_push(&H, adj[v].n, v);
    int k = 0, stamp = 0;
    ivec nbrs = { 0, 0, 0 };
    while (k < nb) {
        hent e = heap_pop(&H);
        int v = e.v;
        if (dead[v]) continue;
        /* compact v's adjacency (drop dead / duplicate entries) */
        ++stamp;
        nbrs.n = 0;
        mark[v] = stamp;
        for (int i = 0; i < adj[v].n; ++i) {
            int u = adj[v].v[i];
            if (!dead[u] && mark[u] != stamp) { mark[u] = stamp; iv_push(&nbrs, u); }
        }
        if (nbrs.n != e.deg) { /* stale key: reinsert with the true degree unless already minimal */
            if (H.n > 0 && nbrs.n > H.h[0].deg) {
                adj[v].n = 0;
                for (int i = 0; i < nbrs.n; ++i) iv_push(&adj[v], nbrs.v[i]);
                heap_push(&H, nbrs.n, v);
                continue;
            }
        }
        dead[v] = 1;
        perm[k++] = v;
        /* connect the neighbours into a clique */
        for (int i = 0; i < nbrs.n; ++i) {
            int u = nbrs.v[i];
            ++stamp;
            int w = 0;
            mark[u] = stamp;
            for (int j = 0; j < adj[u].n; ++j) {
                int x = adj[u].v[j];
                if (!dead[x] && mark[x] != stamp) { mark[x] = stamp; adj[u].v[w++] = x; }
            }
            adj[u].n = w;
            for (int j = 0; j < nbrs.n; ++j) {
                int x = nbrs.v[j];
                if (mark[x] != stamp) { mark[x] = stamp; iv_push(&adj[u], x); }
            }
            heap_push(&H, adj[u].n, u);
        }
        free(adj[v].v); adj[v].v = 0; adj[v].n = adj[v].cap = 0;
    }
    for (int i = 0; i < nb; ++i) free(adj[i].v);
    free(adj); free(dead); free(mark); free(H.h); free(nbrs.v);
}

orc_ldlt *orc_ldlt_analyze(int nb, int d, const int *colptr, const int *rowidx) {
    orc_ldlt *S = (orc_ldlt *)calloc(1, sizeof(orc_ldlt));
    S->nb = nb; S->d = d; S->n = nb * d;
    S->nblocks = colptr[nb];
    S->perm = (int *)malloc(sizeof(int) * (nb > 0 ? nb : 1));
    S->iperm = (int *)malloc(sizeof(int) * (nb > 0 ? nb : 1));
    min_degree_order(nb, colptr, rowidx, S->perm);
    for (int i = 0; i < nb; ++i) S->iperm[S->perm[i]] = i;
    const int n = S->n, dd = d * d;
    /* count entries per permuted scalar column */
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    for (int c = 0; c < nb; ++c)
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
            int r = rowidx[p];
            int R = S->iperm[r], C = S->iperm[c];
            if (r == c) { for (int b = 0; b < d; ++b) cnt[C * d + b] += b + 1; }
            else if (R < C) { for (int b = 0; b < d; ++b) cnt[C * d + b] += d; }
            else { for (int a = 0; a < d; ++a) cnt[R * d + a] += d; }
        }
    S->Ap = (int *)malloc(sizeof(int) * (n + 1));
    S->Ap[0] = 0;
    for (int j = 0; j < n; ++j) S->Ap[j + 1] = S->Ap[j] + cnt[j];
    const int nnz = S->Ap[n];
    S->Ai = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    S->Ax = (double *)calloc(nnz > 0 ? nnz : 1, sizeof(double));
    S->map = (int *)malloc(sizeof(int) * (size_t)(S->nblocks > 0 ? S->nblocks : 1) * dd);
    S->diag_slot = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    memset(cnt, 0, sizeof(int) * (n + 1));
    for (int c = 0; c < nb; ++c)
        for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
            int r = rowidx[p];
            int R = S->iperm[r], C = S->iperm[c];
            for (int a = 0; a < d; ++a)
                for (int b = 0; b < d; ++b) {
                    int row, col;
                    if (r == c) { if (a > b) { S->map[(size_t)p * dd + a * d + b] = -1; continue; } row = C * d + a; col = C * d + b; }
                    else if (R < C) { row = R * d + a; col = C * d + b; }
                    else { row = C * d + b; col = R * d + a; }
                    int slot = S->Ap[col] + cnt[col]++;
                    S->Ai[slot] = row;
                    S->map[(size_t)p * dd + a * d + b] = slot;
                    if (row == col) S->diag_slot[col] = slot;
                }
        }
    free(cnt);
    /* elimination tree + column counts */
    S->parent = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    S->Lnz = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    S->flag = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    S->pattern = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    S->Lp = (int *)malloc(sizeof(int) * (n + 1));
    for (int k = 0; k < n; ++k) {
        S->parent[k] = -1; S->flag[k] = k; S->Lnz[k] = 0;
        for (int p = S->Ap[k]; p < S->Ap[k + 1]; ++p) {
            int i = S->Ai[p];
            if (i < k)
                for (; S->flag[i] != k; i = S->parent[i]) {
                    if (S->parent[i] == -1) S->parent[i] = k;
                    S->Lnz[i]++;
                    S->flag[i] = k;
                }
        }
    }
    S->Lp[0] = 0;
    for (int k = 0; k < n; ++k) S->Lp[k + 1] = S->Lp[k] + S->Lnz[k];
    const long lnz = S->Lp[n];
    S->Li = (int *)malloc(sizeof(int) * (size_t)(lnz > 0 ? lnz : 1));
    S->Lx = (double *)malloc(sizeof(double) * (size_t)(lnz > 0 ? lnz : 1));
    S->D = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    S->Y = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    return S;
}

long orc_ldlt_lnz(const orc_ldlt *S) { return S->Lp[S->n]; }

int orc_ldlt_factor(orc_ldlt *S, const double *blocks, double lambda) {
    const int n = S->n, dd = S->d * S->d;
    for (size_t i = 0; i < (size_t)S->nblocks * dd; ++i)
        if (S->map[i] >= 0) S->Ax[S->map[i]] = blocks[i];
    for (int j = 0; j < n; ++j) S->Ax[S->diag_slot[j]] += lambda;
    const int *Ap = S->Ap, *Ai = S->Ai, *Lp = S->Lp, *parent = S->parent;
    int *Li = S->Li, *Lnz = S->Lnz, *flag = S->flag, *pattern = S->pattern;
    double *Lx = S->Lx, *D = S->D, *Y = S->Y;
    const double *Ax = S->Ax;
    for (int k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int top = n;
        flag[k] = k;
        Lnz[k] = 0;
        for (int p = Ap[k]; p < Ap[k + 1]; ++p) {
            int i = Ai[p];
            if (i <= k) {
                Y[i] += Ax[p];
                int len = 0;
                for (; flag[i] != k; i = parent[i]) { pattern[len++] = i; flag[i] = k; }
                while (len > 0) pattern[--top] = pattern[--len];
            }
        }
        D[k] = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            const int i = pattern[top];
            const double yi = Y[i];
            Y[i] = 0.0;
            const int p2 = Lp[i] + Lnz[i];
            int p;
            for (p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            const double l_ki = yi / D[i];
            D[k] -= l_ki * yi;
            Li[p] = k;
            Lx[p] = l_ki;
            Lnz[i]++;
        }
        if (D[k] == 0.0) return -1; /* SimplicialLDLT reports NumericalIssue */
    }
    return 0;
}

void orc_ldlt_solve(orc_ldlt *S, const double *b, double *x) {
    const int n = S->n, d = S->d;
    double *y = S->Y;
    for (int nbk = 0; nbk < S->nb; ++nbk)
        for (int k = 0; k < d; ++k) y[nbk * d + k] = b[S->perm[nbk] * d + k];
    for (int j = 0; j < n; ++j) {
        const double yj = y[j];
        const int p2 = S->Lp[j] + S->Lnz[j];
        for (int p = S->Lp[j]; p < p2; ++p) y[S->Li[p]] -= S->Lx[p] * yj;
    }
    for (int j = 0; j < n; ++j) y[j] /= S->D[j];
    for (int j = n - 1; j >= 0; --j) {
        double acc = y[j];
        const int p2 = S->Lp[j] + S->Lnz[j];
        for (int p = S->Lp[j]; p < p2; ++p) acc -= S->Lx[p] * y[S->Li[p]];
        y[j] = acc;
    }
    for (int nbk = 0; nbk < S->nb; ++nbk)
        for (int k = 0; k < d; ++k) x[S->perm[nbk] * d + k] = y[nbk * d + k];
    for (int j = 0; j < n; ++j) y[j] = 0;
}

void orc_ldlt_free(orc_ldlt *S) {
    if (!S) return;
    free(S->perm); free(S->iperm); free(S->Ap); free(S->Ai); free(S->Ax); free(S->map);
    free(S->parent); free(S->Lp); free(S->Li); free(S->Lnz); free(S->flag); free(S->pattern);
    free(S->Lx); free(S->D); free(S->Y); free(S->diag_slot);
    free(S);
}
