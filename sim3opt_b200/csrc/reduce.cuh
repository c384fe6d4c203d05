// reduce.cuh -- deterministic block / grid reductions shared by the kernels.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace s3o {

template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sh /* [32] */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (NT / 32) ? sh[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    }
    return v;  // valid in thread 0
}

template <int NT>
__device__ __forceinline__ double block_max(double v, double *sh) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (NT / 32) ? sh[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
    }
    return v;
}

// Ticket: returns true in every thread of the last CTA to arrive; resets the counter.
__device__ __forceinline__ bool last_block(unsigned *counter) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *counter = 0;
    }
    __syncthreads();
    return s_last != 0;
}

template <int NT>
__device__ __forceinline__ double sum_partials(const double *partials, int n, double *sh) {
    double v = 0;
    for (int i = threadIdx.x; i < n; i += NT) v += __ldcg(partials + i);
    return block_sum<NT>(v, sh);
}


// ---- PCG scalar bookkeeping (run by the last CTA on one GPU, or by a 1-thread kernel after the
// NCCL all-reduce in the partitioned solve) ----------------------------------------------------
__device__ __forceinline__ void fin_init(DevScalars *sc, double tol, int max_iter) {
    sc->rz = sc->rz_new;
    sc->rr0 = sc->rr;
    sc->tol2 = tol * tol;
    sc->iters = 0; sc->max_iter = max_iter; sc->alpha = 0; sc->beta = 0; sc->pq = 0;
    sc->done = (sc->rr == 0.0 || !(sc->rz > 0)) ? 1 : 0;
}
__device__ __forceinline__ void fin_spmv(DevScalars *sc) {
    const double pq = sc->pq;
    if (!(pq > 0) || !isfinite(pq)) { sc->done = 3; sc->alpha = 0; }
    else sc->alpha = sc->rz / pq;
}
__device__ __forceinline__ void fin_update(DevScalars *sc) {
    const double rz = sc->rz_new;
    sc->beta = rz / sc->rz;
    sc->rz = rz;
    sc->iters += 1;
    if (!(sc->rr > sc->tol2 * sc->rr0)) sc->done = 1;            // converged (also catches NaN)
    else if (sc->iters >= sc->max_iter) sc->done = 2;
    else if (!(rz > 0)) sc->done = 3;
}

// Symmetric d x d blocks that are only ever multiplied with vectors (block-Jacobi inverses) are stored as
// their packed upper triangle, row by row: D(D+1)/2 doubles per block instead of D*D.
template <int D> __host__ __device__ constexpr int sym_size() { return D * (D + 1) / 2; }
template <int D>
__device__ __forceinline__ int sym_off(int r, int c) {
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    return lo * D - (lo * (lo - 1)) / 2 + (hi - lo);
}

template <int D> struct GroupLanes { static constexpr int value = D > 4 ? 8 : (D > 2 ? 4 : (D > 1 ? 2 : 1)); };

}  // namespace s3o
