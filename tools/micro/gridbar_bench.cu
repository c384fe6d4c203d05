// Micro-benchmark: cost of one grid barrier (counter barrier of csrc/gridbar.cuh vs cooperative_groups grid.sync)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sim3opt_b200/csrc gridbar_bench.cu -o /tmp/gridbar_bench && /tmp/gridbar_bench
#include <cstdio>
#include <cooperative_groups.h>
#include "gridbar.cuh"
using namespace s3o;
__global__ void __launch_bounds__(512, 1) k_counter(GridBarrier gb, int n, double *sink) {
    unsigned phase = 0;
    double v = threadIdx.x;
    for (int i = 0; i < n; ++i) { v = v * 1.0000001 + 1; grid_barrier(gb, phase); }
    if (v == 12345.678) *sink = v;
}
__global__ void __launch_bounds__(512, 1) k_coop(int n, double *sink) {
    double v = threadIdx.x;
    for (int i = 0; i < n; ++i) { v = v * 1.0000001 + 1; __threadfence(); cooperative_groups::this_grid().sync(); }
    if (v == 12345.678) *sink = v;
}
int main() {
    unsigned *bar; double *sink;
    cudaMalloc(&bar, 8); cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {8, 32, 117, 148}) {
        const int n = 2000;
        GridBarrier gb{bar, (int *)(bar + 1), 0};
        float ms1 = 0, ms2 = 0, ms3 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(bar, 0, 8);
            cudaEventRecord(e0); k_counter<<<grid, 512>>>(gb, n, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms1, e0, e1);
            void *args[] = {(void *)&n, &sink};
            cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_coop, dim3(grid), dim3(512), args, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms2, e0, e1);
            int one = 1;
            void *args1[] = {(void *)&one, &sink};
            cudaEventRecord(e0);
            for (int i = 0; i < 100; ++i) cudaLaunchCooperativeKernel((void *)k_coop, dim3(grid), dim3(512), args1, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms3, e0, e1);
        }
        printf("grid %3d: counter barrier %.2f us, cg grid.sync %.2f us per barrier; cooperative launch of a 1-barrier kernel %.1f us each (%s)\n",
               grid, ms1 * 1e3 / n, ms2 * 1e3 / n, ms3 * 1e3 / 100, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
