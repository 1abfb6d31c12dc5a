// l2_bandwidth.cu -- measured L2 read bandwidth of the B200 this runs on (denominator of roofline.l2_frac).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bandwidth l2_bandwidth.cu && ./l2_bandwidth
// Every SM streams a buffer that fits the L2 (16 / 32 / 64 MB, 16-byte loads, 8 independent loads in flight per
// thread) 50 times; the first pass (HBM -> L2) is untimed.  Also a 2 GB stream for the HBM figure of the same kernel.
// Prints one JSON line.
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(512) stream_kernel(const double2* __restrict__ src, size_t n16, int reps, double* sink) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n16; i += 8 * stride) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + i + u * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
        }
        for (; i < n16; i += stride) { double2 v = __ldcg(src + i); acc += v.x + v.y; }
    }
    if (acc == 123.456) *sink = acc;
}

static double run(const double2* d, size_t bytes, int reps, int grid, double* sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    stream_kernel<<<grid, 512>>>(d, bytes / 16, 1, sink);            // warm: fills the L2
    cudaEventRecord(a);
    stream_kernel<<<grid, 512>>>(d, bytes / 16, reps, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return (double)bytes * reps / (ms * 1e-3) / 1e9;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const size_t big = (size_t)2 << 30;
    double2* d = nullptr;
    double* sink = nullptr;
    cudaMalloc(&d, big);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 0, big);
    const int grid = p.multiProcessorCount * 4;
    double best16 = 0, best32 = 0, best64 = 0, hbm = 0;
    for (int it = 0; it < 5; ++it) {
        double v;
        v = run(d, (size_t)16 << 20, 50, grid, sink); if (v > best16) best16 = v;
        v = run(d, (size_t)32 << 20, 50, grid, sink); if (v > best32) best32 = v;
        v = run(d, (size_t)64 << 20, 50, grid, sink); if (v > best64) best64 = v;
        v = run(d, big, 2, grid, sink); if (v > hbm) hbm = v;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"l2_read_gbs_16MB\": %.1f, \"l2_read_gbs_32MB\": %.1f, "
           "\"l2_read_gbs_64MB\": %.1f, \"hbm_read_gbs_2GB\": %.1f}\n",
           p.name, p.multiProcessorCount, p.l2CacheSize, best16, best32, best64, hbm);
    return 0;
}
