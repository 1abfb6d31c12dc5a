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
        if (warp == 0) {                                   // lanes stride over the blocks, fixed xor tree
            double t = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(partials + (size_t)(it & 1) * gridDim.x + b);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) {
                s_h = t;
                if (blockIdx.x == 0) Hcol[it] = t;
            }
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

// ------------------------------------------------------------------------------------
// The whole shifted solve  (M_u + gamma K_uu + gamma^2 G M_eta^-1 D) x = b  as ONE cooperative
// kernel: right-preconditioned (3x3 block-Jacobi) BiCGStab, four grid-wide barriers per
// iteration, every scalar (rho, alpha, omega, the norms) reduced in a fixed order by every block
// from per-block partial sums, so no host round trip and a deterministic result for a fixed
// grid.  The matrix (CSR, 8 lanes per row) and the ten work vectors stay in L2.
//   stage 1   p = r + beta (p - omega v);  y = P^-1 p                      (own 3-blocks)
//   stage 2   v = A y;  r0.v                         -> alpha
//   stage 3   s = r - alpha v;  z = P^-1 s;  s.s, r0.s                     (own 3-blocks)
//   stage 4   t = A z;  t.s, t.t, r0.t               -> omega, rho' = r0.s - omega r0.t
//   stage 5   x += alpha y + omega z;  r = s - omega t   (fused with stage 1 of the next iteration)
// ------------------------------------------------------------------------------------
struct BicgArgs {
    const int* rowptr; const int* col; const double* val;   // A (n_u x n_u CSR)
    const double* Pinv;                                     // 9 doubles per 3x3 block
    const double* b;
    double *x, *r, *r0, *p, *v, *s, *t, *y, *z;
    int nu, max_it;
    double rtol;
    double* partials;                                       // 4 slots x gridDim.x x 4
    int* info;                                              // [0] += iterations, [1] sticky status, [2] iterations of this solve
};

template <int NV>
__device__ inline void bicg_grid_sums(cg::grid_group& grid, double (&v)[NV], double* slot, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = v[k];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[k * 8 + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kKryThreads / 32; ++w) s += sh[threadIdx.x * 8 + w];
        __stcg(slot + (size_t)blockIdx.x * 4 + threadIdx.x, s);
    }
    grid.sync();
    if (warp < NV) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(slot + (size_t)b * 4 + warp);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[32 + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sh[32 + k];
    __syncthreads();
}

// out = A in (8 lanes per row), and up to three dots of the new row values with d0/d1 (and with themselves)
template <int MODE>   // 0: acc[0] += d0[row]*out[row];  1: acc[0] += out*d1[row], acc[1] += out*out, acc[2] += d0[row]*out
__device__ inline void bicg_spmv(const BicgArgs& a, const double* in, double* out, const double* d0, const double* d1, double (&acc)[3]) {
    // kBicgRows rows in flight per 8-lane group and three predicated passes per row (rows of the
    // shifted operator hold 12 or 21 entries), so the rowptr -> col/val -> gather chains of several
    // rows overlap; longer rows finish in a tail loop.
    constexpr int kBicgRows = 4;
    const int sub = threadIdx.x & 7;
    const int grp = (blockIdx.x * kKryThreads + threadIdx.x) >> 3, ngrp = (gridDim.x * kKryThreads) >> 3;
    for (int r0 = grp; r0 < a.nu; r0 += kBicgRows * ngrp) {
        int k0[kBicgRows], k1[kBicgRows];
#pragma unroll
        for (int u = 0; u < kBicgRows; ++u) {
            const int r = r0 + u * ngrp;
            k0[u] = (r < a.nu) ? __ldg(a.rowptr + r) : 0;
            k1[u] = (r < a.nu) ? __ldg(a.rowptr + r + 1) : 0;
        }
        int cc[kBicgRows][3];
        double vv[kBicgRows][3];
#pragma unroll
        for (int u = 0; u < kBicgRows; ++u)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int k = k0[u] + sub + 8 * q;
                const bool ok = k < k1[u];
                cc[u][q] = ok ? __ldg(a.col + k) : -1;
                vv[u][q] = ok ? __ldg(a.val + k) : 0.0;
            }
        double s[kBicgRows];
#pragma unroll
        for (int u = 0; u < kBicgRows; ++u) {
            double xg[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) xg[q] = (cc[u][q] >= 0) ? __ldcg(in + cc[u][q]) : 0.0;
            s[u] = fma(vv[u][2], xg[2], fma(vv[u][1], xg[1], vv[u][0] * xg[0]));
        }
#pragma unroll
        for (int u = 0; u < kBicgRows; ++u) {
            for (int k = k0[u] + sub + 24; k < k1[u]; k += 8) s[u] = fma(__ldg(a.val + k), __ldcg(in + __ldg(a.col + k)), s[u]);
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 4);
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 2);
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 1);
            const int r = r0 + u * ngrp;
            if (sub == 0 && r < a.nu) {
                __stcg(out + r, s[u]);
                if (MODE == 0) acc[0] = fma(__ldcg(d0 + r), s[u], acc[0]);
                else { acc[0] = fma(s[u], __ldcg(d1 + r), acc[0]); acc[1] = fma(s[u], s[u], acc[1]); acc[2] = fma(__ldcg(d0 + r), s[u], acc[2]); }
            }
        }
    }
}

__global__ void __launch_bounds__(kKryThreads)
bicgstab_coop_kernel(BicgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[48];
    const int nblk = a.nu / 3;
    const int gt = blockIdx.x * kKryThreads + threadIdx.x, gs = gridDim.x * kKryThreads;
    double* slot0 = a.partials;
    double* slot1 = a.partials + (size_t)gridDim.x * 4;
    double* slot2 = a.partials + (size_t)gridDim.x * 8;
    double* slot3 = a.partials + (size_t)gridDim.x * 12;

    double red1[1] = {0.0};
    for (int i = gt; i < a.nu; i += gs) {
        const double bi = __ldcg(a.b + i);
        a.x[i] = 0.0; a.r[i] = bi; a.r0[i] = bi; a.p[i] = 0.0; a.v[i] = 0.0;
        red1[0] = fma(bi, bi, red1[0]);
    }
    bicg_grid_sums<1>(grid, red1, slot3, sh);
    const double bb = red1[0];
    int status = 0, it = 0;
    if (bb > 0.0) {
        const double tol2 = a.rtol * a.rtol * bb;
        double rho = 1.0, alpha = 1.0, omega = 1.0, rho_new = bb;
        bool have_update = false;                 // stage 5 of the previous iteration still to be applied
        status = 1;
        for (; it < a.max_it; ++it) {
            if (rho_new == 0.0 || omega == 0.0) { status = 2; break; }
            const double beta = (rho_new / rho) * (alpha / omega);
            // stage 5 (previous iteration) + stage 1
            for (int blk = gt; blk < nblk; blk += gs) {
                double rr[3], pp[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int i = 3 * blk + c;
                    if (have_update) {
                        const double sv = __ldcg(a.s + i);
                        a.x[i] = fma(omega, __ldcg(a.z + i), fma(alpha, __ldcg(a.y + i), __ldcg(a.x + i)));
                        rr[c] = fma(-omega, __ldcg(a.t + i), sv);
                        a.r[i] = rr[c];
                    } else rr[c] = __ldcg(a.r + i);
                    pp[c] = fma(beta, fma(-omega, __ldcg(a.v + i), __ldcg(a.p + i)), rr[c]);
                    a.p[i] = pp[c];
                }
                const double* B = a.Pinv + (size_t)9 * blk;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    __stcg(a.y + 3 * blk + c, fma(__ldg(B + 3 * c), pp[0], fma(__ldg(B + 3 * c + 1), pp[1], __ldg(B + 3 * c + 2) * pp[2])));
            }
            have_update = false;
            grid.sync();
            // stage 2
            double acc[3] = {0.0, 0.0, 0.0};
            bicg_spmv<0>(a, a.y, a.v, a.r0, nullptr, acc);
            double red_a[1] = {acc[0]};
            bicg_grid_sums<1>(grid, red_a, slot0, sh);
            if (red_a[0] == 0.0) { status = 2; break; }
            alpha = rho_new / red_a[0];
            // stage 3
            double red_s[2] = {0.0, 0.0};
            for (int blk = gt; blk < nblk; blk += gs) {
                double ss[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int i = 3 * blk + c;
                    ss[c] = fma(-alpha, __ldcg(a.v + i), __ldcg(a.r + i));
                    __stcg(a.s + i, ss[c]);
                    red_s[0] = fma(ss[c], ss[c], red_s[0]);
                    red_s[1] = fma(__ldcg(a.r0 + i), ss[c], red_s[1]);
                }
                const double* B = a.Pinv + (size_t)9 * blk;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    __stcg(a.z + 3 * blk + c, fma(__ldg(B + 3 * c), ss[0], fma(__ldg(B + 3 * c + 1), ss[1], __ldg(B + 3 * c + 2) * ss[2])));
            }
            bicg_grid_sums<2>(grid, red_s, slot1, sh);
            if (red_s[0] <= tol2) {                // s already small enough: x += alpha y
                for (int i = gt; i < a.nu; i += gs) a.x[i] = fma(alpha, __ldcg(a.y + i), __ldcg(a.x + i));
                status = 0; ++it; break;
            }
            // stage 4
            acc[0] = acc[1] = acc[2] = 0.0;
            bicg_spmv<1>(a, a.z, a.t, a.r0, a.s, acc);
            bicg_grid_sums<3>(grid, acc, slot2, sh);
            omega = (acc[1] > 0.0) ? acc[0] / acc[1] : 0.0;
            rho = rho_new;
            rho_new = fma(-omega, acc[2], red_s[1]);                         // r0.r = r0.s - omega r0.t
            have_update = true;
            const double rr = fma(omega, fma(omega, acc[1], -2.0 * acc[0]), red_s[0]);   // |s - omega t|^2
            if (rr <= tol2) { status = 0; ++it; break; }
        }
        if (have_update)                           // apply the pending x update of the last iteration
            for (int i = gt; i < a.nu; i += gs) a.x[i] = fma(omega, __ldcg(a.z + i), fma(alpha, __ldcg(a.y + i), __ldcg(a.x + i)));
    }
    if (gt == 0) {
        a.info[0] += it;
        a.info[2] = it;
        if (status) a.info[1] = status;
    }
}

// ------------------------------------------------------------------------------------
// The same solve with the MATRIX RESIDENT IN SHARED MEMORY for its whole duration (ref <= 5:
// 24 MB of 3x3 blocks over 148 x 227 KB).  One persistent CTA per SM owns a contiguous range of
// block rows, one thread per block row; its slices of x, r, r0, p, v, s, t, y, z never leave
// registers.  Per iteration only y and z are published (the neighbours gather them), plus one
// partial sum per CTA and reduction.  Layout per CTA ("sliced ELL", built on the host): rows sorted
// by block count, slices of 32 rows padded to the slice width, values [slot][9][lane], columns
// [slot][lane]; padding slots carry zero values and the row's own column.
// ------------------------------------------------------------------------------------
constexpr int kBicgSmemThreads = 512;

struct BicgSmemArgs {
    const double* vals;        // all chunks, doubles
    const int* cols;           // all chunks
    const int* chunk_slice;    // [G+1] first slice of chunk c
    const int* slice_slot;     // [nslices+1] first 32-lane slot of slice s
    const int* chunk_row;      // [G+1] first local row of chunk c (index into rowmap)
    const int* rowmap;         // local row -> global block row
    const double* Pinv;        // 9 doubles per global block row
    const double* b;
    double* x;                 // out (n_u)
    double *y, *z;             // published SpMV inputs (n_u)
    int max_it;
    double rtol;
    double* partials;          // 4 slots x gridDim.x x 4
    int* info;
};

template <int NV>
__device__ inline void bicg_smem_sums(cg::grid_group& grid, double (&v)[NV], double* slot, double* sh) {
    constexpr int NW = kBicgSmemThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = v[k];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[k * NW + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += sh[threadIdx.x * NW + w];
        __stcg(slot + (size_t)blockIdx.x * 4 + threadIdx.x, s);
    }
    grid.sync();                                   // its fence invalidates L1: later plain loads see the other CTAs' stores
    if (warp < NV) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(slot + (size_t)b * 4 + warp);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sh[4 * NW + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sh[4 * NW + k];
    __syncthreads();
}

// o = (A in)(own block row): values and columns from shared memory, `in` gathered from global
__device__ inline void bicg_smem_row(const double* sv, const int* sc, int slot0, int width, int lane,
                                     const double* in, double (&o)[3]) {
    o[0] = o[1] = o[2] = 0.0;
    int q = 0;
    for (; q + 4 <= width; q += 4) {
        int c[4];
        double xg[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = sc[(slot0 + q + u) * 32 + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            xg[u][0] = in[3 * c[u]]; xg[u][1] = in[3 * c[u] + 1]; xg[u][2] = in[3 * c[u] + 2];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double* m = sv + (size_t)(slot0 + q + u) * 9 * 32 + lane;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                o[i] = fma(m[(3 * i) * 32], xg[u][0], fma(m[(3 * i + 1) * 32], xg[u][1], fma(m[(3 * i + 2) * 32], xg[u][2], o[i])));
        }
    }
    for (; q < width; ++q) {
        const int c = sc[(slot0 + q) * 32 + lane];
        const double x0 = in[3 * c], x1 = in[3 * c + 1], x2 = in[3 * c + 2];
        const double* m = sv + (size_t)(slot0 + q) * 9 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            o[i] = fma(m[(3 * i) * 32], x0, fma(m[(3 * i + 1) * 32], x1, fma(m[(3 * i + 2) * 32], x2, o[i])));
    }
}

__global__ void __launch_bounds__(kBicgSmemThreads, 1)
bicgstab_smem_kernel(BicgSmemArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem_dyn[];
    __shared__ double sh[5 * (kBicgSmemThreads / 32)];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x;
    const int s_first = a.chunk_slice[c], s_last = a.chunk_slice[c + 1];
    const int slot_base = a.slice_slot[s_first];
    const int nslots = a.slice_slot[s_last] - slot_base;
    const int row_first = a.chunk_row[c];
    const int rows = a.chunk_row[c + 1] - row_first;
    // shared layout: values [nslots][9][32] | Pinv [9][T] | columns [nslots][32]
    double* sv = smem_dyn;
    double* sP = sv + (size_t)nslots * 9 * 32;
    int* sc = reinterpret_cast<int*>(sP + 9 * kBicgSmemThreads);
    {
        const double* gv = a.vals + (size_t)slot_base * 9 * 32;
        for (int i = threadIdx.x; i < nslots * 9 * 32; i += kBicgSmemThreads) sv[i] = __ldg(gv + i);
        const int* gc = a.cols + (size_t)slot_base * 32;
        for (int i = threadIdx.x; i < nslots * 32; i += kBicgSmemThreads) sc[i] = __ldg(gc + i);
    }
    const bool active = (int)threadIdx.x < rows;
    const int g = active ? __ldg(a.rowmap + row_first + threadIdx.x) : 0;        // global block row
    {
#pragma unroll
        for (int e = 0; e < 9; ++e) sP[e * kBicgSmemThreads + threadIdx.x] = active ? __ldg(a.Pinv + (size_t)9 * g + e) : 0.0;
    }
    int my_slot = 0, width = 0;
    if (active) {
        const int sl = s_first + (threadIdx.x >> 5);
        my_slot = __ldg(a.slice_slot + sl) - slot_base;
        width = __ldg(a.slice_slot + sl + 1) - __ldg(a.slice_slot + sl);
    }
    __syncthreads();
    const double* P = sP + threadIdx.x;
    auto precond = [&](const double (&in)[3], double (&out)[3]) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
            out[i] = fma(P[(3 * i) * kBicgSmemThreads], in[0], fma(P[(3 * i + 1) * kBicgSmemThreads], in[1], P[(3 * i + 2) * kBicgSmemThreads] * in[2]));
    };
    double* slot0 = a.partials;
    double* slot1 = a.partials + (size_t)gridDim.x * 4;
    double* slot2 = a.partials + (size_t)gridDim.x * 8;
    double* slot3 = a.partials + (size_t)gridDim.x * 12;

    double x[3] = {0, 0, 0}, r[3] = {0, 0, 0}, r0[3] = {0, 0, 0}, p[3] = {0, 0, 0}, v[3] = {0, 0, 0};
    double s[3] = {0, 0, 0}, t[3] = {0, 0, 0}, y[3] = {0, 0, 0}, z[3] = {0, 0, 0};
    double red1[1] = {0.0};
    if (active) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { r[i] = r0[i] = __ldcg(a.b + 3 * g + i); red1[0] = fma(r[i], r[i], red1[0]); }
    }
    bicg_smem_sums<1>(grid, red1, slot3, sh);
    const double bb = red1[0];
    int status = 0, it = 0;
    if (bb > 0.0) {
        const double tol2 = a.rtol * a.rtol * bb;
        double rho = 1.0, alpha = 1.0, omega = 1.0, rho_new = bb;
        bool have_update = false;
        status = 1;
        for (; it < a.max_it; ++it) {
            if (rho_new == 0.0 || omega == 0.0) { status = 2; break; }
            const double beta = (rho_new / rho) * (alpha / omega);
            if (active) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (have_update) {
                        x[i] = fma(omega, z[i], fma(alpha, y[i], x[i]));
                        r[i] = fma(-omega, t[i], s[i]);
                    }
                    p[i] = fma(beta, fma(-omega, v[i], p[i]), r[i]);
                }
                precond(p, y);
#pragma unroll
                for (int i = 0; i < 3; ++i) __stcg(a.y + 3 * g + i, y[i]);
            }
            have_update = false;
            grid.sync();
            double red_a[1] = {0.0};
            if (active) {
                bicg_smem_row(sv, sc, my_slot, width, lane, a.y, v);
                red_a[0] = fma(r0[0], v[0], fma(r0[1], v[1], r0[2] * v[2]));
            }
            bicg_smem_sums<1>(grid, red_a, slot0, sh);
            if (red_a[0] == 0.0) { status = 2; break; }
            alpha = rho_new / red_a[0];
            double red_s[2] = {0.0, 0.0};
            if (active) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    s[i] = fma(-alpha, v[i], r[i]);
                    red_s[0] = fma(s[i], s[i], red_s[0]);
                    red_s[1] = fma(r0[i], s[i], red_s[1]);
                }
                precond(s, z);
#pragma unroll
                for (int i = 0; i < 3; ++i) __stcg(a.z + 3 * g + i, z[i]);
            }
            bicg_smem_sums<2>(grid, red_s, slot1, sh);
            if (red_s[0] <= tol2) {
#pragma unroll
                for (int i = 0; i < 3; ++i) x[i] = fma(alpha, y[i], x[i]);
                status = 0; ++it; break;
            }
            double acc[3] = {0.0, 0.0, 0.0};
            if (active) {
                bicg_smem_row(sv, sc, my_slot, width, lane, a.z, t);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc[0] = fma(t[i], s[i], acc[0]);
                    acc[1] = fma(t[i], t[i], acc[1]);
                    acc[2] = fma(r0[i], t[i], acc[2]);
                }
            }
            bicg_smem_sums<3>(grid, acc, slot2, sh);
            omega = (acc[1] > 0.0) ? acc[0] / acc[1] : 0.0;
            rho = rho_new;
            rho_new = fma(-omega, acc[2], red_s[1]);
            have_update = true;
            const double rr = fma(omega, fma(omega, acc[1], -2.0 * acc[0]), red_s[0]);
            if (rr <= tol2) { status = 0; ++it; break; }
        }
        if (have_update) {
#pragma unroll
            for (int i = 0; i < 3; ++i) x[i] = fma(omega, z[i], fma(alpha, y[i], x[i]));
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 3; ++i) a.x[3 * g + i] = x[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.info[0] += it;
        a.info[2] = it;
        if (status) a.info[1] = status;
    }
}

// ------------------------------------------------------------------------------------
// Device-side glue of the Arnoldi loop (no host round trip per step):
//   arnoldi_norm_kernel   hn = sqrt(max(w.Mw, 0)) -> H[j+1][j], scal = 1/hn (0 and a breakdown flag when hn = 0)
//   scale2_kernel         V_{j+1} = scal w,  Z_{j+1} = scal Mw        (krylov_exp.py:78-80)
//   residual_kernel       res[j] = max_q |sqrt(c2)/gamma * fac[q]|    (krylov_exp.py:99-100)
// ------------------------------------------------------------------------------------
__global__ void arnoldi_norm_kernel(const double* __restrict__ dot, double* __restrict__ Hentry, double* __restrict__ scal, int* __restrict__ info,
                                    const double* __restrict__ beta) {
    const double hn = sqrt(fmax(dot[0], 0.0));
    if (Hentry) *Hentry = hn;
    scal[0] = (hn > 0.0) ? 1.0 / hn : 0.0;
    scal[1] = hn;
    // an invariant subspace is a breakdown; a zero start vector (beta = 0: exp(.) 0 = 0) is not
    if (!(hn > 0.0) && info && beta && beta[0] > 0.0) info[3] = 1;
}

__global__ void __launch_bounds__(kKryThreads)
scale2_kernel(double* __restrict__ o1, double* __restrict__ o2, const double* __restrict__ x1, const double* __restrict__ x2,
              const double* __restrict__ scal, int n) {
    const double a = scal[0];
    for (int i = blockIdx.x * kKryThreads + threadIdx.x; i < n; i += gridDim.x * kKryThreads) {
        o1[i] = a * x1[i];
        o2[i] = a * x2[i];
    }
}

__global__ void residual_kernel(const double* __restrict__ c2, double gamma, const double* __restrict__ fac, int nfac, double* __restrict__ out) {
    const double C = sqrt(fmax(c2[0], 0.0));
    double r = 0.0;
    for (int q = 0; q < nfac; ++q) r = fmax(r, fabs(C / gamma * fac[q]));
    *out = r;
}

// Y[k][i] = beta * sum_{j<its} u[k][j] V_j[i]  for B coefficient vectors at once (the sample times of one shared input
// live in ONE Krylov space: only the small exponential depends on t_k).  u: [B][64], beta from device memory.
__global__ void __launch_bounds__(kKryThreads)
combine_batch_kernel(const double* __restrict__ V, int64_t ldv, const double* __restrict__ u, int its, int B, const double* __restrict__ scal,
                     double* __restrict__ Y, int64_t ldy, int n) {
    extern __shared__ double su[];                    // [B][its]
    for (int e = threadIdx.x; e < B * its; e += kKryThreads) su[e] = u[(size_t)(e / its) * 64 + (e % its)];
    __syncthreads();
    const double beta = scal[1];
    for (int i = blockIdx.x * kKryThreads + threadIdx.x; i < n; i += gridDim.x * kKryThreads) {
        for (int k0 = 0; k0 < B; k0 += 8) {
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < its; ++j) {
                const double v = V[(size_t)j * ldv + i];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (k0 + q < B) acc[q] = fma(su[(k0 + q) * its + j], v, acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (k0 + q < B) Y[(size_t)(k0 + q) * ldy + i] = (beta > 0.0) ? beta * acc[q] : 0.0;     // zero input: exp(.) 0 = 0
        }
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
    __shared__ int s_s;
    double *A2 = scratch, *A4 = A2 + m * m, *A6 = A4 + m * m, *U = A6 + m * m, *V = U + m * m, *T1 = V + m * m, *T2 = T1 + m * m;
    if (threadIdx.x == 0) {
        double nrm = 0.0;
        for (int j = 0; j < m; ++j) { double c = 0.0; for (int i = 0; i < m; ++i) c += fabs(A[i * m + j]); nrm = fmax(nrm, c); }
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
