// gridbar.cuh -- grid-wide barrier for persistent kernels launched the ordinary way.
//
// cudaLaunchCooperativeKernel costs ~140 us per launch on this stack (measured in situ with CUDA events: the
// multilevel cooperative kernel took 390 us inside the PCG loop against 246 us of kernel time), which is more than the
// work of the latency-bound kernels that need a grid barrier.  These kernels are launched with an ordinary <<<>>> and
// at most one CTA per SM, so that every CTA is resident as soon as the previous kernel on the stream has drained, and
// synchronise through a counter in global memory: arrival = atomicAdd, wait = spin on the count reaching
// phase * gridDim.x.  The counter is zeroed by a 4-byte memset node before every launch.
// Contract: the library runs one stream per problem and one process per GPU (SURVEY.md 8b "Threading"); a kernel from
// another stream that holds SMs forever would starve the barrier, so the spin carries a time-out that raises `abort`
// (every later barrier then falls through and the caller reports a breakdown instead of hanging).
// S3O_COOP_LAUNCH=1 selects cudaLaunchCooperativeKernel + cooperative_groups::grid_group::sync() instead.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace s3o {

struct GridBarrier {
    unsigned *counter;      // zero at kernel start
    int *abort;             // set on time-out
    int cooperative;        // 1: launched with cudaLaunchCooperativeKernel, use grid.sync()
};

// all threads of all CTAs call this the same number of times; `phase` is a per-thread running count (uniform)
__device__ __forceinline__ void grid_barrier(const GridBarrier &B, unsigned &phase) {
    if (B.cooperative) {
        __threadfence();
        cooperative_groups::this_grid().sync();
        return;
    }
    ++phase;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(B.counter, 1u);
        const unsigned target = phase * gridDim.x;
        long long spins = 0;
        while (*reinterpret_cast<volatile unsigned *>(B.counter) < target) {
            if (*reinterpret_cast<volatile int *>(B.abort)) break;
            if (++spins > (1ll << 28)) { *B.abort = 1; break; }
        }
        __threadfence();
    }
    __syncthreads();
}

}  // namespace s3o
