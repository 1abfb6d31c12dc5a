// kexp_krylov_kernels.cuh -- single-vector kernels of the shift-and-invert Arnoldi path
// (krylov_exp.py:50-109): warp-per-row CSR SpMV with shuffle row reductions, fixed-order
// multi-dot, a cooperative modified-Gram-Schmidt sweep, and the small dense Hessenberg
// inverse / matrix exponential on one CTA.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kexp {
namespace cg = cooperative_groups;

constexpr int kKryThreads = 256;

// y = A x, one warp per row, lanes stride over the row, xor-shuffle tree (fixed order)
__global__ void __launch_bounds__(kKryThreads)
spmv_warp_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                 const double* __restrict__ x, double* __restrict__ y, int nrows) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * kKryThreads + threadIdx.x) >> 5, nw = (gridDim.x * kKryThreads) >> 5;
    for (int r = gw; r < nrows; r += nw) {
        const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
        double s = 0.0;
        for (int k = k0 + lane; k < k1; k += 32) s = fma(__ldg(val + k), x[__ldg(col + k)], s);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) y[r] = s;
    }
}

// y = blockdiag(P) x, 3x3 blocks
__global__ void __launch_bounds__(kKryThreads)
blk3_apply_kernel(const double* __restrict__ P, const double* __restrict__ x, double* __restrict__ y, int nblocks) {
    for (int b = blockIdx.x * kKryThreads + threadIdx.x; b < nblocks; b += gridDim.x * kKryThreads) {
        const double* B = P + (size_t)9 * b;
        const double x0 = x[3 * b], x1 = x[3 * b + 1], x2 = x[3 * b + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i) y[3 * b + i] = fma(B[3 * i], x0, fma(B[3 * i + 1], x1, B[3 * i + 2] * x2));
    }
}

// out = a*x + b*y + c*z   (null pointers are skipped)
__global__ void __launch_bounds__(kKryThreads)
lincomb3_kernel(double* __restrict__ out, double a, const double* x, double b, const double* y, double c, const double* z, int n) {
    for (int i = blockIdx.x * kKryThreads + threadIdx.x; i < n; i += gridDim.x * kKryThreads) {
        double v = 0.0;
        if (x) v = a * x[i];
        if (y) v = fma(b, y[i], v);
        if (z) v = fma(c, z[i], v);
        out[i] = v;
    }
}

// up to 4 dot products a_k . b_k in one pass; partials[block][k], the last block sums them
// in block order and writes out[k] (device) -- deterministic for a fixed grid.
struct DotArgs {
    const double* a[4];
    const double* b[4];
    int n[4];
    int count;
};

__global__ void __launch_bounds__(kKryThreads)
multidot_kernel(DotArgs args, double* __restrict__ partials, unsigned* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sh[4][kKryThreads / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < args.count; ++k) {
        double s = 0.0;
        for (int i = blockIdx.x * kKryThreads + threadIdx.x; i < args.n[k]; i += gridDim.x * kKryThreads)
            s = fma(args.a[k][i], args.b[k][i], s);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < args.count) {
        double s = 0.0;
        for (int w = 0; w < kKryThreads / 32; ++w) s += sh[threadIdx.x][w];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *counter = 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < args.count) {
        __threadfence();
        double s = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + (size_t)b * 4 + threadIdx.x);
        out[threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------
// Modified Gram-Schmidt sweep of krylov_exp.py:73-77 in the M inner product, one
// cooperative kernel:   for i = 0..j:  h_i = w . Z_i  (Z_i = M V_i);  w -= h_i V_i
// Every block keeps its slice of w in registers; per step the blocks exchange one partial
// sum each through global memory and a grid-wide barrier, and every block adds the
// partials in the same fixed order.  H(:,j) is written by block 0.
// ------------------------------------------------------------------------------------
constexpr int kMgsPerThread = 8;

__global__ void __launch_bounds__(kKryThreads)
mgs_sweep_kernel(double* __restrict__ w, const double* __restrict__ V, const double* __restrict__ Z, int n, int64_t ldv,
                 int j, double* __restrict__ partials, double* __restrict__ Hcol) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[kKryThreads / 32];
    __shared__ double s_h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slice = kKryThreads * kMgsPerThread;
    const int base = blockIdx.x * slice;
    double wv[kMgsPerThread];
#pragma unroll
    for (int e = 0; e < kMgsPerThread; ++e) {
        const int i = base + e * kKryThreads + threadIdx.x;
        wv[e] = (i < n) ? w[i] : 0.0;
    }
    for (int it = 0; it <= j; ++it) {
        const double* Zi = Z + (size_t)it * ldv;
        const double* Vi = V + (size_t)it * ldv;
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < kMgsPerThread; ++e) {
            const int i = base + e * kKryThreads + threadIdx.x;
            if (i < n) s = fma(wv[e], Zi[i], s);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < kKryThreads / 32; ++q) t += sh[q];
            partials[(size_t)(it & 1) * gridDim.x + blockIdx.x] = t;
        }
        grid.sync();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(partials + (size_t)(it & 1) * gridDim.x + b);
            s_h = t;
            if (blockIdx.x == 0) Hcol[it] = t;
        }
        __syncthreads();
        const double h = s_h;
#pragma unroll
        for (int e = 0; e < kMgsPerThread; ++e) {
            const int i = base + e * kKryThreads + threadIdx.x;
            if (i < n) wv[e] = fma(-h, Vi[i], wv[e]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < kMgsPerThread; ++e) {
        const int i = base + e * kKryThreads + threadIdx.x;
        if (i < n) w[i] = wv[e];
    }
}

// y = beta * sum_{j<its} u_j V_j     (krylov_exp.py:102-109)
__global__ void __launch_bounds__(kKryThreads)
combine_kernel(const double* __restrict__ V, int64_t ldv, const double* __restrict__ u, int its, double beta,
               double* __restrict__ y, int n) {
    for (int i = blockIdx.x * kKryThreads + threadIdx.x; i < n; i += gridDim.x * kKryThreads) {
        double s = 0.0;
        for (int j = 0; j < its; ++j) s = fma(u[j], V[(size_t)j * ldv + i], s);
        y[i] = beta * s;
    }
}

// ------------------------------------------------------------------------------------
// Small dense algebra on ONE CTA (m <= 64): matrices live in a global scratch area (L1/L2
// resident), every step is separated by __syncthreads.
//   inv(H)           Gauss-Jordan with partial pivoting            (krylov_exp.py:88)
//   expm(A)          Pade-13 scaling and squaring (Higham 2005)     (krylov_exp.py:96)
// ------------------------------------------------------------------------------------
constexpr int kDenseThreads = 1024;

__device__ inline void dn_matmul(const double* A, const double* B, double* C, int m) {
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int i = e / m, j = e % m;
        double s = 0.0;
        for (int k = 0; k < m; ++k) s = fma(A[i * m + k], B[k * m + j], s);
        C[e] = s;
    }
    __syncthreads();
}

// Solve A X = B in place (A destroyed, B <- X); both m x m row-major
__device__ inline void dn_solve(double* A, double* B, int m) {
    __shared__ int s_piv;
    for (int k = 0; k < m; ++k) {
        if (threadIdx.x == 0) {
            int p = k; double best = fabs(A[k * m + k]);
            for (int i = k + 1; i < m; ++i) { const double v = fabs(A[i * m + k]); if (v > best) { best = v; p = i; } }
            s_piv = p;
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k) {
            for (int j = threadIdx.x; j < 2 * m; j += blockDim.x) {
                double* M = (j < m) ? A : B;
                const int jj = (j < m) ? j : j - m;
                const double t = M[k * m + jj]; M[k * m + jj] = M[p * m + jj]; M[p * m + jj] = t;
            }
            __syncthreads();
        }
        const double piv = A[k * m + k];
        __syncthreads();
        for (int j = threadIdx.x; j < 2 * m; j += blockDim.x) {
            double* M = (j < m) ? A : B;
            const int jj = (j < m) ? j : j - m;
            M[k * m + jj] /= piv;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < m * 2 * m; e += blockDim.x) {
            const int i = e / (2 * m), j = e % (2 * m);
            if (i == k) continue;
            const double f = A[i * m + k];
            // column k of A is still needed by the other threads of this step: it is zeroed after the barrier
            if (j < m) { if (j != k) A[i * m + j] = fma(-f, A[k * m + j], A[i * m + j]); }
            else B[i * m + (j - m)] = fma(-f, B[k * m + (j - m)], B[i * m + (j - m)]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) if (i != k) A[i * m + k] = 0.0;
        __syncthreads();
    }
}

// E = expm(A); scratch: 7 m*m doubles.  A is overwritten.
__device__ inline void dn_expm(double* A, double* E, double* scratch, int m) {
    const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
                          129060195264000.0, 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0,
                          40840800.0, 960960.0, 16380.0, 182.0, 1.0};
    __shared__ double s_norm;
    __shared__ int s_s;
    double *A2 = scratch, *A4 = A2 + m * m, *A6 = A4 + m * m, *U = A6 + m * m, *V = U + m * m, *T1 = V + m * m, *T2 = T1 + m * m;
    if (threadIdx.x == 0) {
        double nrm = 0.0;
        for (int j = 0; j < m; ++j) { double c = 0.0; for (int i = 0; i < m; ++i) c += fabs(A[i * m + j]); nrm = fmax(nrm, c); }
        s_norm = nrm;
        int s = 0;
        const double theta13 = 5.371920351148152;
        if (nrm > theta13) s = (int)ceil(log2(nrm / theta13));
        s_s = s;
    }
    __syncthreads();
    const int s = s_s;
    const double scale = ldexp(1.0, -s);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) A[e] *= scale;
    __syncthreads();
    dn_matmul(A, A, A2, m);
    dn_matmul(A2, A2, A4, m);
    dn_matmul(A4, A2, A6, m);
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        T1[e] = b[13] * A6[e] + b[11] * A4[e] + b[9] * A2[e];
        T2[e] = b[12] * A6[e] + b[10] * A4[e] + b[8] * A2[e];
    }
    __syncthreads();
    dn_matmul(A6, T1, U, m);       // U <- A6*(b13 A6 + b11 A4 + b9 A2)
    dn_matmul(A6, T2, V, m);       // V <- A6*(b12 A6 + b10 A4 + b8 A2)
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const double id = (e / m == e % m) ? 1.0 : 0.0;
        T1[e] = U[e] + b[7] * A6[e] + b[5] * A4[e] + b[3] * A2[e] + b[1] * id;
        V[e] = V[e] + b[6] * A6[e] + b[4] * A4[e] + b[2] * A2[e] + b[0] * id;
    }
    __syncthreads();
    dn_matmul(A, T1, U, m);        // U = A * (...)
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        T1[e] = V[e] - U[e];       // (V - U) X = (V + U)
        E[e] = V[e] + U[e];
    }
    __syncthreads();
    dn_solve(T1, E, m);
    for (int q = 0; q < s; ++q) {
        dn_matmul(E, E, T1, m);
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) E[e] = T1[e];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kDenseThreads)
dense_expm_kernel(double* A, double* E, double* scratch, int m) { dn_expm(A, E, scratch, m); }

// For the leading (j+1)x(j+1) block of H (leading dimension ldh):
//   Hinv = inv(H);  H1 = (Hinv - I)/gamma;  for s in svals[0..ns): u_s = expm(-H1 s t) e1,
//   factor_s = (Hinv u_s)[j]      (krylov_exp.py:87-99)
// out_u <- u of the LAST s (j+1 doubles), out_factor[ns].  scratch: 11 m*m doubles.
__global__ void __launch_bounds__(kDenseThreads)
hessenberg_exp_kernel(const double* __restrict__ H, int ldh, int m, double gamma, double t, const double* __restrict__ svals,
                      int ns, double* __restrict__ out_u, double* __restrict__ out_factor, double* scratch) {
    double *Hinv = scratch, *Hc = Hinv + m * m, *H1 = Hc + m * m, *E = H1 + m * m, *work = E + m * m;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        Hc[e] = H[(e / m) * ldh + (e % m)];
        Hinv[e] = (e / m == e % m) ? 1.0 : 0.0;
    }
    __syncthreads();
    dn_solve(Hc, Hinv, m);
    for (int q = 0; q < ns; ++q) {
        const double st = svals[q] * t;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const double id = (e / m == e % m) ? 1.0 : 0.0;
            H1[e] = -st * (Hinv[e] - id) / gamma;
        }
        __syncthreads();
        dn_expm(H1, E, work, m);
        if (threadIdx.x == 0) {
            double f = 0.0;
            for (int k = 0; k < m; ++k) f = fma(Hinv[(m - 1) * m + k], E[k * m], f);
            out_factor[q] = f;
        }
        if (q == ns - 1)
            for (int i = threadIdx.x; i < m; i += blockDim.x) out_u[i] = E[i * m];
        __syncthreads();
    }
}

}  // namespace kexp
