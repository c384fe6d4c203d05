// direct.cu -- numeric sparse block Cholesky + triangular solves on the device (see direct.h).
//
// One kernel per LM trial does what LinearSolverEigen::solve [EXT g2o] does per trial (numeric LDL^T of
// Hpp + lambda I, two triangular solves; reference plug-in site kitti_surf.cpp:553-557):
//   scatter   L <- permuted (H + lambda I), fill-in blocks zeroed
//   per level (the independent sets of the multiple-minimum-degree order, direct_host.cpp):
//     gather  every block of the level's columns subtracts its ordered list of L_ij L_kj^T products
//     pivot   Cholesky of the level's d x d diagonal blocks, inverse of the triangular factor kept
//     scale   sub-diagonal blocks of the level's columns times L_kk^-T
//   forward / backward substitution level by level (8-lane group per block row, shuffles inside the group)
// Every block has one writer and a fixed summation order: bitwise reproducible.  Small factors (the KITTI-size
// graphs) run in ONE CTA with __syncthreads between phases -- the whole solve is one launch and no
// host round trip; larger ones run as a cooperative grid (one CTA per SM) with grid barriers.
#include <cooperative_groups.h>

#include "direct.h"
#include "problem.h"
#include "reduce.cuh"

namespace cg = cooperative_groups;

namespace s3o {

struct DirectDev {
    int n, nlev, nL;
    const int32_t *perm, *lev_ptr, *cptr, *brow, *bcol, *src, *upd_ptr, *upd_a, *upd_b, *row_ptr, *row_blk, *row_col;
    double *L;       // [nL][D*D]
    double *Linv;    // [n][D*D]   inverse of the lower-triangular pivot factors
    double *y;       // [n][D]     solve workspace, elimination order
    int *fail;
};

struct DirectState {
    DirectPlan plan;
    bool analyzed = false, available = false;
    int32_t *d_perm = nullptr, *d_lev_ptr = nullptr, *d_cptr = nullptr, *d_brow = nullptr, *d_bcol = nullptr, *d_src = nullptr;
    int32_t *d_upd_ptr = nullptr, *d_upd_a = nullptr, *d_upd_b = nullptr, *d_row_ptr = nullptr, *d_row_blk = nullptr, *d_row_col = nullptr;
    double *d_L = nullptr, *d_Linv = nullptr, *d_y = nullptr;
    int *d_fail = nullptr;
    int coop_grid = 0;          // CTAs of the cooperative launch (<= 1: single CTA)
    int wide_levels = 0;        // leading levels wide enough for the whole grid; the rest runs on CTA 0
    bool factored = false;
};

namespace {

constexpr int kDirectNT = 512;

// loads of data written earlier in the same kernel by other SMs must bypass the (non-coherent) L1
template <bool COOP> __device__ __forceinline__ double ldw(const double *p) { return COOP ? __ldcg(p) : *p; }

// ---- phases (all take the executing thread space: tid in [0, nthreads)) ---------------------------------------
// gather: every block of the level's columns subtracts its ordered list of L[a] L[b]^T products
template <int D, bool COOP>
__device__ __forceinline__ void dk_gather(const DirectDev &P, int lev, int tid, int nthreads) {
    constexpr int DD = D * D;
    const int c0 = P.lev_ptr[lev], c1 = P.lev_ptr[lev + 1];
    const int t0 = P.cptr[c0], t1 = P.cptr[c1];
    for (int item = tid; item < (t1 - t0) * DD; item += nthreads) {
        const int t = t0 + item / DD, e = item % DD;
        const int q0 = P.upd_ptr[t], q1 = P.upd_ptr[t + 1];
        if (q0 == q1) continue;
        const int r = e / D, c = e - r * D;
        double acc = ldw<COOP>(P.L + (size_t)t * DD + e);
        for (int q = q0; q < q1; ++q) {
            const double *La = P.L + (size_t)P.upd_a[q] * DD + r * D;
            const double *Lb = P.L + (size_t)P.upd_b[q] * DD + c * D;
            double s = 0;
#pragma unroll
            for (int m = 0; m < D; ++m) s += ldw<COOP>(La + m) * ldw<COOP>(Lb + m);
            acc -= s;
        }
        P.L[(size_t)t * DD + e] = acc;
    }
}

// lower Cholesky of one d x d pivot block and the inverse of the factor (one thread)
template <int D, bool COOP>
__device__ __forceinline__ void dk_pivot_one(const DirectDev &P, int k) {
    constexpr int DD = D * D;
    double *Ld = P.L + (size_t)P.cptr[k] * DD;
    double *Lik = P.Linv + (size_t)k * DD;
    double A[DD], Li[DD];
#pragma unroll
    for (int i = 0; i < DD; ++i) A[i] = ldw<COOP>(Ld + i);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double djj = A[j * D + j];
#pragma unroll
        for (int m = 0; m < j; ++m) djj -= A[j * D + m] * A[j * D + m];
        if (!(djj > 0) || !isfinite(djj)) { ok = false; djj = 1; }
        const double ljj = sqrt(djj), inv = 1.0 / ljj;
        A[j * D + j] = ljj;
#pragma unroll
        for (int r = j + 1; r < D; ++r) {
            double v = A[r * D + j];
#pragma unroll
            for (int m = 0; m < j; ++m) v -= A[r * D + m] * A[j * D + m];
            A[r * D + j] = v * inv;
        }
    }
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
        for (int r = 0; r < D; ++r) {
            if (r < c) { Li[r * D + c] = 0; continue; }
            double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int m = c; m < r; ++m) v -= A[r * D + m] * Li[m * D + c];
            Li[r * D + c] = v / A[r * D + r];
        }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            Ld[r * D + c] = c <= r ? A[r * D + c] : 0.0;
            Lik[r * D + c] = Li[r * D + c];
        }
    if (!ok) *P.fail = 1;
}

// scale one row of a sub-diagonal block: L_ik[r,:] = A_ik[r,:] L_kk^-T
template <int D, bool COOP>
__device__ __forceinline__ void dk_scale_row(const DirectDev &P, int t, int r, const double *Lik) {
    constexpr int DD = D * D;
    double a[D];
    double *row = P.L + (size_t)t * DD + r * D;
#pragma unroll
    for (int m = 0; m < D; ++m) a[m] = ldw<COOP>(row + m);
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double sacc = 0;
#pragma unroll
        for (int m = 0; m <= c; ++m) sacc += a[m] * ldw<COOP>(Lik + c * D + m);
        row[c] = sacc;
    }
}

// A level with many columns: one thread per pivot (32 pivots per warp instruction), barrier, then one thread per
// block row.  A level with few columns: one warp per column, lane 0 factors, the warp scales -- no barrier between.
template <int D, bool COOP, class Sync>
__device__ __forceinline__ void dk_pivot_scale(const DirectDev &P, int lev, int tid, int nthreads, Sync sync) {
    constexpr int DD = D * D;
    const int c0 = P.lev_ptr[lev], c1 = P.lev_ptr[lev + 1];
    const int nwarps = nthreads >> 5;
    if (c1 - c0 > 2 * nwarps) {
        for (int k = c0 + tid; k < c1; k += nthreads) dk_pivot_one<D, COOP>(P, k);
        sync();
        const int t0 = P.cptr[c0], t1 = P.cptr[c1];
        for (int item = tid; item < (t1 - t0) * D; item += nthreads) {
            const int t = t0 + item / D, r = item % D;
            const int k = P.bcol[t];
            if (P.brow[t] == k) continue;
            dk_scale_row<D, COOP>(P, t, r, P.Linv + (size_t)k * DD);
        }
        return;
    }
    const int warp = tid >> 5, lane = tid & 31;
    for (int k = c0 + warp; k < c1; k += nwarps) {
        if (lane == 0) { dk_pivot_one<D, COOP>(P, k); __threadfence(); }
        __syncwarp();
        const int nsub = P.cptr[k + 1] - P.cptr[k] - 1;
        for (int item = lane; item < nsub * D; item += 32)
            dk_scale_row<D, COOP>(P, P.cptr[k] + 1 + item / D, item % D, P.Linv + (size_t)k * DD);
    }
}

// forward substitution of one level: y_k = L_kk^-1 (b_perm(k) - sum_j L_kj y_j), a GL-lane group per block row
template <int D, bool COOP>
__device__ __forceinline__ void dk_forward(const DirectDev &P, int lev, const double *__restrict__ b, int tid, int nthreads) {
    constexpr int DD = D * D, GL = GroupLanes<D>::value;
    const int group = tid / GL, lane = tid % GL, ngroups = nthreads / GL;
    const int c0 = P.lev_ptr[lev], c1 = P.lev_ptr[lev + 1];
    for (int kb = c0; kb < c1; kb += ngroups) {       // uniform trip count: the shuffles below stay convergent
        const int k = kb + group;
        const bool act = k < c1 && lane < D;
        double acc = 0;
        if (act) {
            acc = b[(size_t)P.perm[k] * D + lane];
            for (int q = P.row_ptr[k]; q < P.row_ptr[k + 1]; ++q) {
                const double *Lt = P.L + (size_t)P.row_blk[q] * DD + lane * D;
                const double *yj = P.y + (size_t)P.row_col[q] * D;
                double s = 0;
#pragma unroll
                for (int m = 0; m < D; ++m) s += ldw<COOP>(Lt + m) * ldw<COOP>(yj + m);
                acc -= s;
            }
        }
        double yl = 0;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const double am = __shfl_sync(0xffffffffu, acc, m, GL);
            if (act && m <= lane) yl += ldw<COOP>(P.Linv + (size_t)k * DD + lane * D + m) * am;
        }
        if (act) P.y[(size_t)k * D + lane] = yl;
    }
}

// backward substitution of one level: x_k = L_kk^-T (y_k - sum_i L_ik^T x_i), written in Hessian order as well
template <int D, bool COOP>
__device__ __forceinline__ void dk_backward(const DirectDev &P, int lev, double *__restrict__ x, int tid, int nthreads) {
    constexpr int DD = D * D, GL = GroupLanes<D>::value;
    const int group = tid / GL, lane = tid % GL, ngroups = nthreads / GL;
    const int c0 = P.lev_ptr[lev], c1 = P.lev_ptr[lev + 1];
    for (int kb = c0; kb < c1; kb += ngroups) {
        const int k = kb + group;
        const bool act = k < c1 && lane < D;
        double acc = 0;
        if (act) {
            acc = ldw<COOP>(P.y + (size_t)k * D + lane);
            for (int t = P.cptr[k] + 1; t < P.cptr[k + 1]; ++t) {
                const double *Lt = P.L + (size_t)t * DD + lane;
                const double *xi = P.y + (size_t)P.brow[t] * D;
                double s = 0;
#pragma unroll
                for (int m = 0; m < D; ++m) s += ldw<COOP>(Lt + m * D) * ldw<COOP>(xi + m);
                acc -= s;
            }
        }
        double xl = 0;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const double am = __shfl_sync(0xffffffffu, acc, m, GL);
            if (act && m >= lane) xl += ldw<COOP>(P.Linv + (size_t)k * DD + m * D + lane) * am;
        }
        if (act) {
            P.y[(size_t)k * D + lane] = xl;
            x[(size_t)P.perm[k] * D + lane] = xl;
        }
    }
}

// Schedule.  The first `wide` levels (the rounds that eliminate many columns at once) run on the whole grid with
// grid barriers; the narrow tail -- most of the rounds, each a handful of blocks -- runs on CTA 0 alone with CTA
// barriers while the other CTAs wait at ONE grid barrier.  Per level: [gather(l) + forward(l-1)] | pivot+scale(l);
// the forward substitution of a level only needs that level's pivots, so it rides in the next gather phase.
//   COOP = false: one CTA, wide = 0.
template <int D, bool COOP>
__global__ void __launch_bounds__(kDirectNT) direct_kernel(DirectDev P, const double *__restrict__ H, double lambda,
                                                           const double *__restrict__ b, double *__restrict__ x,
                                                           DevScalars *sc, int do_factor, int wide, const GridBarrier gb) {
    constexpr int DD = D * D;
    unsigned phase = 0;
    const int ltid = threadIdx.x;
    const int gtid = blockIdx.x * kDirectNT + threadIdx.x, gthreads = gridDim.x * kDirectNT;
    auto gsync = [&]() {
        if constexpr (COOP) grid_barrier(gb, phase);
        else __syncthreads();
    };
    if (do_factor) {
        if (gtid == 0) *P.fail = 0;
        // ---- scatter (H + lambda I) into the factor's storage, elimination order
        for (long long item = gtid; item < (long long)P.nL * DD; item += gthreads) {
            const int t = (int)(item / DD), e = (int)(item - (long long)t * DD);
            const int r = e / D, c = e - r * D;
            const int s = P.src[t];
            double v = 0;
            if (s >= 0) {
                const double *Hb = H + (size_t)(s >> 1) * DD;
                v = (s & 1) ? Hb[c * D + r] : Hb[e];
            }
            if (r == c && P.brow[t] == P.bcol[t]) v += lambda;
            P.L[item] = v;
        }
        gsync();
    }
    // ---- wide levels on the whole grid
    for (int lev = 0; lev < wide; ++lev) {
        if (do_factor) dk_gather<D, COOP>(P, lev, gtid, gthreads);
        if (lev > 0) dk_forward<D, COOP>(P, lev - 1, b, gtid, gthreads);
        if (do_factor || lev > 0) gsync();
        if (do_factor) { dk_pivot_scale<D, COOP>(P, lev, gtid, gthreads, gsync); gsync(); }
    }
    // ---- narrow tail on CTA 0 (CTA barriers), then its share of the backward substitution
    if (blockIdx.x == 0) {
        for (int lev = wide; lev < P.nlev; ++lev) {
            if (do_factor) dk_gather<D, COOP>(P, lev, ltid, kDirectNT);
            if (lev > 0) dk_forward<D, COOP>(P, lev - 1, b, ltid, kDirectNT);
            __syncthreads();
            if (do_factor) { dk_pivot_scale<D, COOP>(P, lev, ltid, kDirectNT, [] { __syncthreads(); }); __syncthreads(); }
        }
        if (P.nlev > 0) {
            if (wide == P.nlev) { /* the last level's forward step still runs on the grid below */ }
            else { dk_forward<D, COOP>(P, P.nlev - 1, b, ltid, kDirectNT); __syncthreads(); }
        }
        for (int lev = P.nlev - 1; lev >= wide; --lev) { dk_backward<D, COOP>(P, lev, x, ltid, kDirectNT); __syncthreads(); }
    }
    if (wide > 0) {
        if (wide == P.nlev) { dk_forward<D, COOP>(P, P.nlev - 1, b, gtid, gthreads); }
        gsync();
        for (int lev = wide - 1; lev >= 0; --lev) { dk_backward<D, COOP>(P, lev, x, gtid, gthreads); gsync(); }
    }
    if (gtid == 0) {     // the PCG bookkeeping the LM driver reads: an exact solve, 0 iterations
        const int failed = *reinterpret_cast<volatile int *>(P.fail);
        sc->done = failed ? 3 : 1;
        sc->iters = 0;
        sc->rr = 0;
        sc->rr0 = 1;
    }
}

template <class T>
int up(s3o_problem *p, T **dst, const std::vector<T> &src) { return upload(p, dst, src); }

void free_direct_arrays(DirectState *D) {
    dev_free(D->d_perm); dev_free(D->d_lev_ptr); dev_free(D->d_cptr); dev_free(D->d_brow); dev_free(D->d_bcol); dev_free(D->d_src);
    dev_free(D->d_upd_ptr); dev_free(D->d_upd_a); dev_free(D->d_upd_b); dev_free(D->d_row_ptr); dev_free(D->d_row_blk);
    dev_free(D->d_row_col); dev_free(D->d_L); dev_free(D->d_Linv); dev_free(D->d_y); dev_free(D->d_fail);
}

template <int D>
int launch_direct(s3o_problem *p, const DirectDev &P, double lambda, int do_factor) {
    DirectState *S = p->direct;
    if (S->coop_grid <= 1) {
        direct_kernel<D, false><<<1, kDirectNT, 0, p->stream>>>(P, p->d_H, lambda, p->d_b, p->d_x, p->d_sc, do_factor, 0, GridBarrier{});
        return S3O_OK;
    }
    const double *H = p->d_H, *b = p->d_b;
    double *x = p->d_x;
    DevScalars *sc = p->d_sc;
    DirectDev Pc = P;
    int wide = S->wide_levels;
    GridBarrier gb{};
    void *args[] = { &Pc, &H, &lambda, &b, &x, &sc, &do_factor, &wide, &gb };
    return launch_persistent(p, (const void *)direct_kernel<D, true>, S->coop_grid, kDirectNT, args, 9);
}

template <int D>
int coop_capacity(int device) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, direct_kernel<D, true>, kDirectNT, 0) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return per_sm > 0 ? sms : 0;      // one CTA per SM
}

}  // namespace

void direct_destroy(s3o_problem *p) {
    if (!p->direct) return;
    free_direct_arrays(p->direct);
    delete p->direct;
    p->direct = nullptr;
}

// Block products above which the AUTO rule stays with PCG / above which the analysis is abandoned.
static constexpr long long kAutoMaxPairs = 400000, kForcedMaxPairs = 60000000;
static constexpr int kAutoMaxLevels = 128;

// Symbolic analysis + upload, once per structure.  Returns S3O_OK also when the factor is too large
// (direct_available() then says no).
int direct_setup(s3o_problem *p) {
    if (p->direct && p->direct->analyzed) return S3O_OK;
    if (!p->direct) p->direct = new DirectState();
    DirectState *D = p->direct;
    D->analyzed = true;
    D->available = false;
    if (p->dist || p->S.nf == 0) return S3O_OK;
    const bool forced = p->linsolver == S3O_LINSOLVER_DIRECT;
    // AUTO: a factor within kAutoMaxPairs block products cannot have more than ~kAutoMaxPairs / 3 columns (a chain
    // costs 3 products per column), so larger graphs go to the PCG without paying for the analysis
    if (!forced && p->S.nf > kAutoMaxPairs / 3) return S3O_OK;
    if (!direct_analyze(p->S.nf, p->S.rowptr, p->S.colidx, forced ? kForcedMaxPairs : kAutoMaxPairs, D->plan)) return S3O_OK;
    const DirectPlan &P = D->plan;
    if (!forced && P.nlev > kAutoMaxLevels) return S3O_OK;
    const int nL = (int)P.brow.size(), d = p->d;
    std::vector<int32_t> bcol(nL);
    for (int j = 0; j < P.n; ++j)
        for (int t = P.cptr[j]; t < P.cptr[j + 1]; ++t) bcol[t] = j;
    int rc = 0;
    rc = rc ? rc : up(p, &D->d_perm, P.perm);
    rc = rc ? rc : up(p, &D->d_lev_ptr, P.lev_ptr);
    rc = rc ? rc : up(p, &D->d_cptr, P.cptr);
    rc = rc ? rc : up(p, &D->d_brow, P.brow);
    rc = rc ? rc : up(p, &D->d_bcol, bcol);
    rc = rc ? rc : up(p, &D->d_src, P.src);
    rc = rc ? rc : up(p, &D->d_upd_ptr, P.upd_ptr);
    rc = rc ? rc : up(p, &D->d_upd_a, P.upd_a);
    rc = rc ? rc : up(p, &D->d_upd_b, P.upd_b);
    rc = rc ? rc : up(p, &D->d_row_ptr, P.row_ptr);
    rc = rc ? rc : up(p, &D->d_row_blk, P.row_blk);
    rc = rc ? rc : up(p, &D->d_row_col, P.row_col);
    rc = rc ? rc : dev_alloc(&D->d_L, (size_t)nL * d * d);
    rc = rc ? rc : dev_alloc(&D->d_Linv, (size_t)P.n * d * d);
    rc = rc ? rc : dev_alloc(&D->d_y, (size_t)P.n * d);
    rc = rc ? rc : dev_alloc(&D->d_fail, 1);
    if (rc) { free_direct_arrays(D); return rc; }
    S3O_CUDA(cudaStreamSynchronize(p->stream));
    // Grid and schedule: a level is "wide" while its gather phase has more than 8 items per thread of one CTA;
    // the leading wide levels run on a cooperative grid sized for the widest one, everything after the last wide
    // level on CTA 0 alone (CTA barriers are ~10x cheaper than grid barriers, and the tail is pure latency).
    D->coop_grid = 0;
    D->wide_levels = 0;
    {
        long long widest = 0;
        for (int l = 0; l < P.nlev; ++l) {
            const long long items = (long long)(P.cptr[P.lev_ptr[l + 1]] - P.cptr[P.lev_ptr[l]]) * d * d;
            if (items > 8 * kDirectNT) D->wide_levels = l + 1;
            widest = std::max(widest, items);
        }
        if (D->wide_levels > 0) {
            int cap = 0;
            switch (d) {
            case 7: cap = coop_capacity<7>(p->device); break;
            case 6: cap = coop_capacity<6>(p->device); break;
            case 4: cap = coop_capacity<4>(p->device); break;
            case 1: cap = coop_capacity<1>(p->device); break;
            }
            const long long want = (widest + 4 * kDirectNT - 1) / (4 * kDirectNT);
            D->coop_grid = (int)std::max<long long>(1, std::min<long long>(cap, want));
            if (D->coop_grid <= 1) D->wide_levels = 0;
        }
    }
    D->available = true;
    D->factored = false;
    p->stats.direct_levels = P.nlev;
    p->stats.direct_blocks = nL;
    return S3O_OK;
}

bool direct_available(const s3o_problem *p) { return p->direct && p->direct->available; }
void direct_invalidate(s3o_problem *p) { if (p->direct) p->direct->factored = false; }

// Solve (H + lambda I) x = b exactly: numeric factorisation (skipped when reuse_factor and a factor of the
// same system is still held) + both triangular solves, one launch.  x in p->d_x; status through DevScalars.
int direct_solve(s3o_problem *p, double lambda, bool reuse_factor) {
    DirectState *D = p->direct;
    if (!D || !D->available) { set_error("direct_solve: no factorisation plan"); return S3O_ERR_INVALID; }
    const DirectPlan &Pl = D->plan;
    DirectDev P{};
    P.n = Pl.n; P.nlev = Pl.nlev; P.nL = (int)Pl.brow.size();
    P.perm = D->d_perm; P.lev_ptr = D->d_lev_ptr; P.cptr = D->d_cptr; P.brow = D->d_brow; P.bcol = D->d_bcol; P.src = D->d_src;
    P.upd_ptr = D->d_upd_ptr; P.upd_a = D->d_upd_a; P.upd_b = D->d_upd_b;
    P.row_ptr = D->d_row_ptr; P.row_blk = D->d_row_blk; P.row_col = D->d_row_col;
    P.L = D->d_L; P.Linv = D->d_Linv; P.y = D->d_y; P.fail = D->d_fail;
    const int do_factor = (reuse_factor && D->factored) ? 0 : 1;
    int rc = S3O_OK;
    switch (p->d) {
    case 7: rc = launch_direct<7>(p, P, lambda, do_factor); break;
    case 6: rc = launch_direct<6>(p, P, lambda, do_factor); break;
    case 4: rc = launch_direct<4>(p, P, lambda, do_factor); break;
    case 1: rc = launch_direct<1>(p, P, lambda, do_factor); break;
    default: set_error("direct_solve: block dimension %d", p->d); return S3O_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    D->factored = true;
    p->stats.direct_solves += 1;
    return check_launch(p, 1);
}

}  // namespace s3o
