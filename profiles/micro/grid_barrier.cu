// Micro-benchmark: cost of one grid-wide barrier on B200 -- cooperative_groups grid.sync() against a
// hand-written sense-free counter barrier (one atomic per CTA, acquire polling by thread 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_barrier grid_barrier.cu && ./grid_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_cg(int iters, double* sink) {
    cg::grid_group grid = cg::this_grid();
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) { acc += 1.0; grid.sync(); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// counter barrier: generation g completes when the counter reaches g * gridDim.x
__device__ __forceinline__ void my_barrier(unsigned* ctr, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        red_release(ctr, 1u);
        const unsigned target = gen * gridDim.x;
        while (ld_acquire(ctr) < target) { }
    }
    __syncthreads();
}

__global__ void k_my(int iters, double* sink, unsigned* ctr) {
    unsigned gen = 0;
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) { acc += 1.0; my_barrier(ctr, gen); }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* sink; unsigned* ctr;
    cudaMalloc(&sink, 8); cudaMalloc(&ctr, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2000;
    for (int per_sm = 1; per_sm <= 4; per_sm *= 2)
        for (int threads = 256; threads <= 512; threads *= 2) {
            if (per_sm * threads > 2048) continue;
            int grid = per_sm * sms, it = iters;
            void* args1[] = {&it, &sink};
            void* args2[] = {&it, &sink, &ctr};
            float ms1 = 0, ms2 = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(threads), args1, 0, 0);
                cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms1, a, b);
                cudaMemset(ctr, 0, 4);
                cudaEventRecord(a);
                cudaLaunchCooperativeKernel((void*)k_my, dim3(grid), dim3(threads), args2, 0, 0);
                cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms2, a, b);
            }
            printf("grid %4d x %3d threads: cg grid.sync %.2f us, counter barrier %.2f us  (%s)\n", grid, threads,
                   ms1 * 1e3 / iters, ms2 * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
