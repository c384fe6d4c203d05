// align.cu -- similarity alignment of two trajectories and its RMSE.
//
// Replaces estimateSimilarityTransform + the RMSE loop of the reference's evaluation modes
// (kitti_surf.cpp:1091-1161 and :1432-1452): Eigen::umeyama(query, train, with_scaling = true)
// [Umeyama 1991: c, R, t minimising sum |train - (c R query + t)|^2], or the reference's
// "only scale" variant (ratio of the coordinate extents of the dominant axes x and z).
// The moments are three deterministic two-stage reductions on the device (means; centred
// cross-covariance and query variance; residuals of the aligned points), the 3x3 SVD in between
// runs on the host (one-sided Jacobi).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "internal.h"
#include "problem.h"
#include "reduce.cuh"

namespace s3o {
namespace {

constexpr int kAlignNT = 256, kAlignMaxGrid = 1024;

// block maximum of signed values (block_max in reduce.cuh serves |H_jj| and pads with 0)
template <int NT>
__device__ __forceinline__ double block_max_signed(double v, double *sh) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (NT / 32) ? sh[lane] : -DBL_MAX;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
    }
    return v;
}

// pass 0: NV = 6  sums of query and train coordinates, and the extremes used by the only-scale variant
// pass 1: NV = 10 centred cross-covariance sum (t - mt)(q - mq)^T (9, row-major) and sum |q - mq|^2
// pass 2: NV = 1  sum |t - (M q + tt)|^2 and max |.| (par holds M (9) and tt (3))
template <int PASS>
__global__ void __launch_bounds__(kAlignNT) align_pass_kernel(int n, const double *__restrict__ q, const double *__restrict__ t,
                                                              const double *__restrict__ par, double *__restrict__ partial,
                                                              double *__restrict__ out, unsigned *counter) {
    constexpr int NV = PASS == 0 ? 18 : (PASS == 1 ? 10 : 2);
    // pass 0 layout: [0..5] sums, [6..8] max q, [9..11] min q, [12..14] max t, [15..17] min t
    __shared__ double sh[32];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0;
    if (PASS == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { acc[6 + k] = -DBL_MAX; acc[9 + k] = DBL_MAX; acc[12 + k] = -DBL_MAX; acc[15 + k] = DBL_MAX; }
    }
    for (int i = blockIdx.x * kAlignNT + threadIdx.x; i < n; i += gridDim.x * kAlignNT) {
        const double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        const double tx = t[3 * i], ty = t[3 * i + 1], tz = t[3 * i + 2];
        if (PASS == 0) {
            const double v[6] = { qx, qy, qz, tx, ty, tz };
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] += v[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[6 + k] = fmax(acc[6 + k], v[k]); acc[9 + k] = fmin(acc[9 + k], v[k]);
                acc[12 + k] = fmax(acc[12 + k], v[3 + k]); acc[15 + k] = fmin(acc[15 + k], v[3 + k]);
            }
        } else if (PASS == 1) {
            const double dq[3] = { qx - par[0], qy - par[1], qz - par[2] };
            const double dt[3] = { tx - par[3], ty - par[4], tz - par[5] };
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[r * 3 + c] += dt[r] * dq[c];
            acc[9] += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
        } else {
            const double ex = tx - (par[0] * qx + par[1] * qy + par[2] * qz + par[9]);
            const double ey = ty - (par[3] * qx + par[4] * qy + par[5] * qz + par[10]);
            const double ez = tz - (par[6] * qx + par[7] * qy + par[8] * qz + par[11]);
            const double d2 = ex * ex + ey * ey + ez * ez;
            acc[0] += d2;
            acc[1] = fmax(acc[1], d2);
        }
    }
    // per-CTA reduction of every accumulator, then the last CTA combines the partials in index order
    auto is_max = [](int k) { return (PASS == 0 && ((k >= 6 && k < 9) || (k >= 12 && k < 15))) || (PASS == 2 && k == 1); };
    auto is_min = [](int k) { return PASS == 0 && ((k >= 9 && k < 12) || k >= 15); };
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v;
        if (is_max(k)) v = block_max_signed<kAlignNT>(acc[k], sh);
        else if (is_min(k)) v = -block_max_signed<kAlignNT>(-acc[k], sh);
        else v = block_sum<kAlignNT>(acc[k], sh);
        if (threadIdx.x == 0) partial[(size_t)k * kAlignMaxGrid + blockIdx.x] = v;
    }
    if (last_block(counter)) {
        if (threadIdx.x < NV) {
            const int k = threadIdx.x;
            double v = __ldcg(partial + (size_t)k * kAlignMaxGrid);
            for (int b = 1; b < (int)gridDim.x; ++b) {
                const double w = __ldcg(partial + (size_t)k * kAlignMaxGrid + b);
                v = is_max(k) ? fmax(v, w) : (is_min(k) ? fmin(v, w) : v + w);
            }
            out[k] = v;
        }
    }
}

void jacobi_svd3(const double A[9], double U[9], double S[3], double V[9]) {
    // one-sided Jacobi on the columns of W = A: W V = U diag(S)
    double W[9];
    for (int i = 0; i < 9; ++i) { W[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double a = 0, b = 0, c = 0;
                for (int r = 0; r < 3; ++r) { a += W[r * 3 + p] * W[r * 3 + p]; b += W[r * 3 + q] * W[r * 3 + q]; c += W[r * 3 + p] * W[r * 3 + q]; }
                off = std::max(off, std::fabs(c) / std::sqrt(std::max(a * b, 1e-300)));
                if (std::fabs(c) < 1e-300) continue;
                const double zeta = (b - a) / (2 * c);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const double cs = 1 / std::sqrt(1 + tt * tt), sn = cs * tt;
                for (int r = 0; r < 3; ++r) {
                    const double wp = W[r * 3 + p], wq = W[r * 3 + q];
                    W[r * 3 + p] = cs * wp - sn * wq; W[r * 3 + q] = sn * wp + cs * wq;
                    const double vp = V[r * 3 + p], vq = V[r * 3 + q];
                    V[r * 3 + p] = cs * vp - sn * vq; V[r * 3 + q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    int order[3] = { 0, 1, 2 };
    double nrm[3];
    for (int c = 0; c < 3; ++c) nrm[c] = std::sqrt(W[c] * W[c] + W[3 + c] * W[3 + c] + W[6 + c] * W[6 + c]);
    std::sort(order, order + 3, [&](int a, int b) { return nrm[a] > nrm[b]; });
    double Vs[9];
    for (int k = 0; k < 3; ++k) {
        const int c = order[k];
        S[k] = nrm[c];
        for (int r = 0; r < 3; ++r) { U[r * 3 + k] = nrm[c] > 0 ? W[r * 3 + c] / nrm[c] : 0.0; Vs[r * 3 + k] = V[r * 3 + c]; }
    }
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
    // complete a rank-deficient U to an orthonormal basis (degenerate trajectories)
    if (S[2] <= 1e-300 * std::max(S[0], 1.0) || nrm[order[2]] == 0) {
        U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

}  // namespace
}  // namespace s3o

using namespace s3o;

extern "C" int s3o_align_similarity(int device, int n, const double *query_xyz, const double *train_xyz, int only_scale,
                                    double *S221, double *rmse, double *max_dev) {
    if (n < 1 || !query_xyz || !train_xyz || !S221) { set_error("s3o_align_similarity: bad arguments"); return S3O_ERR_INVALID; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("s3o_align_similarity: no CUDA device %d (there is no CPU fallback)", device); return S3O_ERR_CUDA; }
    double *d_q = nullptr, *d_t = nullptr, *d_par = nullptr, *d_partial = nullptr, *d_out = nullptr;
    unsigned *d_counter = nullptr;
    int rc = 0;
    rc = rc ? rc : dev_alloc(&d_q, (size_t)n * 3);
    rc = rc ? rc : dev_alloc(&d_t, (size_t)n * 3);
    rc = rc ? rc : dev_alloc(&d_par, 12);
    rc = rc ? rc : dev_alloc(&d_partial, (size_t)18 * kAlignMaxGrid);
    rc = rc ? rc : dev_alloc(&d_out, 18);
    rc = rc ? rc : dev_alloc(&d_counter, 1);
    auto cleanup = [&]() { cudaFree(d_q); cudaFree(d_t); cudaFree(d_par); cudaFree(d_partial); cudaFree(d_out); cudaFree(d_counter); };
    if (rc) { cleanup(); return rc; }
    cudaError_t e = cudaMemcpy(d_q, query_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_t, train_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_counter, 0, sizeof(unsigned));
    const int grid = std::max(1, std::min(kAlignMaxGrid, (n + kAlignNT - 1) / kAlignNT));
    double h[18], par[12];
    auto fetch = [&](int cnt) { return cudaMemcpy(h, d_out, sizeof(double) * cnt, cudaMemcpyDeviceToHost); };
    if (e == cudaSuccess) { align_pass_kernel<0><<<grid, kAlignNT>>>(n, d_q, d_t, d_par, d_partial, d_out, d_counter); e = fetch(18); }
    if (e != cudaSuccess) { set_error("s3o_align_similarity: %s", cudaGetErrorString(e)); cleanup(); return S3O_ERR_CUDA; }
    double M[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }, tt[3] = { 0, 0, 0 };
    if (only_scale) {
        // kitti_surf.cpp:1104-1136: mean ratio of the coordinate extents over the dominant axes x and z;
        // no rotation, no translation (KITTI ground truth starts at the origin with identity orientation)
        const double sx = (h[12] - h[15]) / (h[6] - h[9]), sz = (h[14] - h[17]) / (h[8] - h[11]);
        const double s = 0.5 * (sx + sz);
        M[0] = M[4] = M[8] = s;
    } else {
        for (int k = 0; k < 6; ++k) par[k] = h[k] / n;
        const double mq[3] = { par[0], par[1], par[2] }, mt[3] = { par[3], par[4], par[5] };
        e = cudaMemcpy(d_par, par, sizeof(double) * 6, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { align_pass_kernel<1><<<grid, kAlignNT>>>(n, d_q, d_t, d_par, d_partial, d_out, d_counter); e = fetch(10); }
        if (e != cudaSuccess) { set_error("s3o_align_similarity: %s", cudaGetErrorString(e)); cleanup(); return S3O_ERR_CUDA; }
        double Sigma[9], U[9], Sv[3], V[9];
        for (int k = 0; k < 9; ++k) Sigma[k] = h[k] / n;
        const double var_q = h[9] / n;
        jacobi_svd3(Sigma, U, Sv, V);
        double sgn[3] = { 1, 1, 1 };
        if (det3(U) * det3(V) < 0) sgn[2] = -1;
        double R[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R[r * 3 + c] = U[r * 3] * sgn[0] * V[c * 3] + U[r * 3 + 1] * sgn[1] * V[c * 3 + 1] + U[r * 3 + 2] * sgn[2] * V[c * 3 + 2];
        const double c = var_q > 0 ? (Sv[0] * sgn[0] + Sv[1] * sgn[1] + Sv[2] * sgn[2]) / var_q : 1.0;
        for (int k = 0; k < 9; ++k) M[k] = c * R[k];
        for (int r = 0; r < 3; ++r) tt[r] = mt[r] - (M[r * 3] * mq[0] + M[r * 3 + 1] * mq[1] + M[r * 3 + 2] * mq[2]);
    }
    for (int k = 0; k < 9; ++k) par[k] = M[k];
    for (int k = 0; k < 3; ++k) par[9 + k] = tt[k];
    e = cudaMemcpy(d_par, par, sizeof(double) * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { align_pass_kernel<2><<<grid, kAlignNT>>>(n, d_q, d_t, d_par, d_partial, d_out, d_counter); e = fetch(2); }
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) { set_error("s3o_align_similarity: %s", cudaGetErrorString(e)); return S3O_ERR_CUDA; }
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) S221[r * 4 + c] = M[r * 3 + c];
        S221[r * 4 + 3] = tt[r];
    }
    S221[12] = S221[13] = S221[14] = 0; S221[15] = 1;
    if (rmse) *rmse = std::sqrt(h[0] / n);
    if (max_dev) *max_dev = std::sqrt(h[1]);
    return S3O_OK;
}
