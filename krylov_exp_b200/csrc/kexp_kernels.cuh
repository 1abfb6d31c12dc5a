// kexp_kernels.cuh -- device kernels of the batched Chebyshev exp(sL)v path (sm_100a).
//
// Data layout in HBM: a batch "panel" is a row-major [rows][W] array of doubles, W in
// {2,4,8,16,32,64} sample columns contiguous per dof ("batch-minor").  A group of
// LPR = W/2 lanes owns one row and each lane one double2 (two sample columns), so a
// gathered row X[col,:] is one fully coalesced W*8-byte segment (512 B at W = 64) and
// every matrix entry is used for W columns.  All reductions are two-stage and
// fixed-order (warp shuffles -> shared memory -> per-block partials -> last block), so
// results are bit-reproducible for a given panel width and grid.
//
// Reference semantics implemented here (file:line in /root/reference):
//   operator_solver.solve()  averaged_sw_explicit.py:142-164   -> spmm + batched PCG
//   three-term recurrence    cheby_exp.py:138-165              -> cheb_* kernels
//   DU *= wt; allreduce      timestepping_methods.py:10-11     -> weighted_colsum
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kexp {

constexpr int kThreads = 256;
constexpr int kMaxW = 64;

// Per-panel CG scalars, one slot per sample column (device memory).
struct CgScalars {
    double rz[kMaxW];      // r . P^-1 r
    double pq[kMaxW];      // p . A p
    double alpha[kMaxW];
    double beta[kMaxW];
    double bb[kMaxW];      // ||b||^2
    double rr[kMaxW];      // ||r||^2
    int active[kMaxW];     // column still iterating
    int done;              // all columns converged: remaining CG kernels of the graph return at once
    int iters;             // iterations of the current solve
    unsigned long long total_iters;
    unsigned long long total_solves;
    double max_rel_res;    // of the last solve
    double rtol2;
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// streaming variants: vectors that are touched once per kernel should not evict the
// matrix / gathered rows from L1
__device__ __forceinline__ double2 ld2_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Sum over the rows handled by one warp (lanes with equal lane % LPR), fixed order.
template <int LPR>
__device__ __forceinline__ double2 warp_reduce_rows(double2 v) {
#pragma unroll
    for (int off = 16; off >= LPR; off >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
    }
    return v;
}

// Block partial of NACC per-thread double2 accumulators -> partials[(a*grid + block)*W + col].
// Returns true in the last block to arrive (which may then read all partials).
template <int W, int NACC>
__device__ __forceinline__ bool block_partials(double2 (&acc)[NACC], double* __restrict__ partials,
                                               unsigned* __restrict__ counter) {
    constexpr int LPR = W / 2;
    __shared__ double2 sh[NACC][kThreads / 32][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double2 v = warp_reduce_rows<LPR>(acc[a]);
        if (lane < LPR) sh[a][warp][lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < LPR * NACC) {
        const int a = threadIdx.x / LPR, l = threadIdx.x % LPR;
        double2 s = sh[a][0][l];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) { s.x += sh[a][w][l].x; s.y += sh[a][w][l].y; }
        st2(partials + ((size_t)a * gridDim.x + blockIdx.x) * W + 2 * l, s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(counter, 1u);
        s_last = (ticket == gridDim.x - 1);
        if (s_last) *counter = 0;          // ready for the next launch
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// Last block: column sums of the per-block partials in fixed order. Result for column c
// of accumulator a is returned to threads with threadIdx.x == c (c < W) through `out`.
template <int W, int NACC>
__device__ __forceinline__ void final_sums(const double* __restrict__ partials, double (&out)[NACC]) {
    constexpr int NSPLIT = kThreads / W;
    __shared__ double red[NACC][NSPLIT][W];
    const int c = threadIdx.x % W, part = threadIdx.x / W;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double s = 0.0;
        const double* base = partials + (size_t)a * gridDim.x * W + c;
        int b = part;
        for (; b + 7 * NSPLIT < (int)gridDim.x; b += 8 * NSPLIT) {   // 8 loads in flight, summed in order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + (size_t)(b + u * NSPLIT) * W);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < (int)gridDim.x; b += NSPLIT) s += __ldcg(base + (size_t)b * W);
        red[a][part][c] = s;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double s = 0.0;
        if (threadIdx.x < W) {
#pragma unroll
            for (int p = 0; p < NSPLIT; ++p) s += red[a][p][threadIdx.x];
        }
        out[a] = s;
    }
}

// ------------------------------------------------------------------------------------
// CSR SpMM  Y[r,:] = sum_k val[k] * X[col[k],:]   for r in [0,nrows)
// One LPR-lane group per row; the group loads up to 32 (col,val) pairs of the warp's rows
// cooperatively and broadcasts them with shuffles; X rows are gathered as double2.
// MODE 0: plain.  MODE 1 (CG): also accumulates p.q per column, the last block turns the
// partials into pq and alpha = rz/pq  (the dot product fused into the SpMM epilogue).
// ------------------------------------------------------------------------------------
template <int W, int MODE>
__global__ void __launch_bounds__(kThreads)
spmm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
            const double* __restrict__ X, double* __restrict__ Y, int nrows,
            CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    constexpr int LPR = W / 2, RPW = 32 / LPR;
    if (MODE == 1 && sc->done) return;
    const int lane = threadIdx.x & 31;
    const int l = lane % LPR, sub = lane / LPR;
    const int gwarp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
    double2 dot[1] = {make_double2(0.0, 0.0)};

    for (int rowbase = gwarp * RPW; rowbase < nrows; rowbase += nwarps * RPW) {
        const int r = rowbase + sub;
        int k0 = 0, k1 = 0;
        if (r < nrows) { k0 = __ldg(rowptr + r); k1 = __ldg(rowptr + r + 1); }
        int len = k1 - k0, maxlen = len;
#pragma unroll
        for (int off = 16; off >= LPR; off >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
        double2 acc = make_double2(0.0, 0.0);
        for (int kb = 0; kb < maxlen; kb += LPR) {
            const int kk = k0 + kb + l;
            int cc = 0; double aa = 0.0;
            if (kb + l < len) { cc = __ldg(col + kk); aa = __ldg(val + kk); }
            const int cnt = min(LPR, maxlen - kb);
            if (cnt == LPR) {
#pragma unroll
                for (int j = 0; j < LPR; ++j) {
                    const int c = __shfl_sync(0xffffffffu, cc, j, LPR);
                    const double a = __shfl_sync(0xffffffffu, aa, j, LPR);
                    const double2 x = ld2(X + (size_t)c * W + 2 * l);
                    acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y);
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const int c = __shfl_sync(0xffffffffu, cc, j, LPR);
                    const double a = __shfl_sync(0xffffffffu, aa, j, LPR);
                    const double2 x = ld2(X + (size_t)c * W + 2 * l);
                    acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y);
                }
            }
        }
        if (r < nrows) {
            st2(Y + (size_t)r * W + 2 * l, acc);
            if (MODE == 1) {
                const double2 p = ld2(X + (size_t)r * W + 2 * l);
                dot[0].x = fma(p.x, acc.x, dot[0].x); dot[0].y = fma(p.y, acc.y, dot[0].y);
            }
        }
    }
    if (MODE == 1) {
        if (block_partials<W, 1>(dot, partials, counter)) {
            double s[1];
            final_sums<W, 1>(partials, s);
            if (threadIdx.x < W) {
                const int c = threadIdx.x;
                sc->pq[c] = s[0];
                sc->alpha[c] = (sc->active[c] && s[0] > 0.0) ? sc->rz[c] / s[0] : 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// BSR(3x3) SpMM -- the layout BDM2/DG1 matrices have naturally (3 dofs per mesh entity).
// A block is 80 bytes: 9 values row-major + the block-column index in the low word of the
// 10th double, so one block is five aligned 16-byte broadcast loads.  A group of LPB lanes
// owns one block row (3 output rows) and each lane CPT = min(W,8) sample columns: every
// gathered X value feeds 3 FMAs and every matrix value 8, which takes the L1/shared-memory
// data path off the critical path.  Each CTA walks a CONTIGUOUS range of block rows so the
// gathered rows of neighbouring mesh entities are re-used out of L1 instead of L2.
// MODE as in spmm_kernel.
// ------------------------------------------------------------------------------------
template <int W>
struct BsrCfg {
    static constexpr int CPT = (W >= 8) ? 8 : W;   // sample columns per thread
    static constexpr int LPB = W / CPT;            // lanes per block row
    static constexpr int BRW = 32 / LPB;           // block rows per warp
    static constexpr int NV = CPT / 2;             // double2 per row per thread
};

template <int W, int MODE>
__global__ void __launch_bounds__(kThreads, 2)
bsr3_spmm_kernel(const int* __restrict__ browptr, const double* __restrict__ blocks,
                 const double* __restrict__ X, double* __restrict__ Y, int nbr,
                 CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    using C = BsrCfg<W>;
    constexpr int LPB = C::LPB, BRW = C::BRW, NV = C::NV;
    if (MODE == 1 && sc->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const int per = nbr / gridDim.x, rem = nbr % gridDim.x;
    const int start = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int end = start + per + ((int)blockIdx.x < rem ? 1 : 0);
    double2 dot[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) dot[j] = make_double2(0.0, 0.0);

    for (int base = start; base < end; base += (kThreads / 32) * BRW) {
        const int br = base + warp * BRW + g;
        if (br < end) {
            double2 acc[3][NV];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < NV; ++j) acc[i][j] = make_double2(0.0, 0.0);
            const int k0 = __ldg(browptr + br), k1 = __ldg(browptr + br + 1);
            for (int k = k0; k < k1; ++k) {
                const double2* bp = reinterpret_cast<const double2*>(blocks + (size_t)10 * k);
                const double2 v01 = __ldg(bp), v23 = __ldg(bp + 1), v45 = __ldg(bp + 2), v67 = __ldg(bp + 3), v8c = __ldg(bp + 4);
                const int c = __double2loint(v8c.y);
                const double a[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c.x}};
                const double* xr = X + (size_t)(3 * c) * W + 2 * l;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double2 x[NV];
#pragma unroll
                    for (int j = 0; j < NV; ++j) x[j] = ld2(xr + (size_t)r * W + 2 * LPB * j);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < NV; ++j) {
                            acc[i][j].x = fma(a[i][r], x[j].x, acc[i][j].x);
                            acc[i][j].y = fma(a[i][r], x[j].y, acc[i][j].y);
                        }
                }
            }
            double* yr = Y + (size_t)(3 * br) * W + 2 * l;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < NV; ++j) st2(yr + (size_t)i * W + 2 * LPB * j, acc[i][j]);
            if (MODE == 1) {
                const double* pr = X + (size_t)(3 * br) * W + 2 * l;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const double2 p = ld2(pr + (size_t)i * W + 2 * LPB * j);
                        dot[j].x = fma(p.x, acc[i][j].x, dot[j].x);
                        dot[j].y = fma(p.y, acc[i][j].y, dot[j].y);
                    }
            }
        }
    }
    if (MODE == 1) {
        // block partial: lanes with equal l across the warp's block rows, then across warps
        __shared__ double2 sh[kThreads / 32][W / 2];
        __shared__ bool s_last;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double2 v = dot[j];
#pragma unroll
            for (int off = 16; off >= LPB; off >>= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            }
            if (lane < LPB) sh[warp][l + LPB * j] = v;
        }
        __syncthreads();
        if (threadIdx.x < W / 2) {
            double2 s = sh[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < kThreads / 32; ++w) { s.x += sh[w][threadIdx.x].x; s.y += sh[w][threadIdx.x].y; }
            st2(partials + (size_t)blockIdx.x * W + 2 * threadIdx.x, s);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(counter, 1u);
            s_last = (ticket == gridDim.x - 1);
            if (s_last) *counter = 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double s[1];
            final_sums<W, 1>(partials, s);
            if (threadIdx.x < W) {
                const int c = threadIdx.x;
                sc->pq[c] = s[0];
                sc->alpha[c] = (sc->active[c] && s[0] > 0.0) ? sc->rz[c] / s[0] : 0.0;
            }
        }
    }
}

// 3x3 block (row-major, 9 doubles) times three double2 values
__device__ __forceinline__ void blk3_mul(const double* __restrict__ B, const double2 (&r)[3], double2 (&z)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double b0 = __ldg(B + 3 * i), b1 = __ldg(B + 3 * i + 1), b2 = __ldg(B + 3 * i + 2);
        z[i].x = fma(b0, r[0].x, fma(b1, r[1].x, b2 * r[2].x));
        z[i].y = fma(b0, r[0].y, fma(b1, r[1].y, b2 * r[2].y));
    }
}

// ------------------------------------------------------------------------------------
// PCG start:  x = 0, r = b (in place: r aliases b), p = P^-1 r, rz = r.p, bb = r.r
// One LPR-lane group per dof triple (the 3x3 block-Jacobi unit).
// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(const double* __restrict__ r, double* __restrict__ x, double* __restrict__ p,
               const double* __restrict__ Pinv, int ntriples,
               CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    constexpr int LPR = W / 2;
    const int l = threadIdx.x % LPR;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    double2 acc[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3], z[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) rv[i] = ld2_stream(r + o + (size_t)i * W);
        blk3_mul(Pinv + (size_t)9 * g, rv, z);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            st2(x + o + (size_t)i * W, make_double2(0.0, 0.0));
            st2(p + o + (size_t)i * W, z[i]);
            acc[0].x = fma(rv[i].x, z[i].x, acc[0].x); acc[0].y = fma(rv[i].y, z[i].y, acc[0].y);
            acc[1].x = fma(rv[i].x, rv[i].x, acc[1].x); acc[1].y = fma(rv[i].y, rv[i].y, acc[1].y);
        }
    }
    if (block_partials<W, 2>(acc, partials, counter)) {
        double s[2];
        final_sums<W, 2>(partials, s);
        __shared__ int any_active;
        if (threadIdx.x == 0) any_active = 0;
        __syncthreads();
        if (threadIdx.x < W) {
            const int c = threadIdx.x;
            sc->rz[c] = s[0]; sc->bb[c] = s[1]; sc->rr[c] = s[1];
            const int act = (s[1] > 0.0) ? 1 : 0;
            sc->active[c] = act;
            if (act) atomicOr(&any_active, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sc->done = any_active ? 0 : 1;
            sc->iters = 0;
            sc->max_rel_res = 0.0;
            sc->total_solves += 1ull;
        }
    }
}

// ------------------------------------------------------------------------------------
// PCG update:  x += alpha p,  r -= alpha q,  z = P^-1 r,  rz' = r.z,  rr = r.r
// last block: convergence per column, beta = rz'/rz, rz = rz'.
// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
cg_update_xr_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                    const double* __restrict__ q, const double* __restrict__ Pinv, int ntriples,
                    CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    constexpr int LPR = W / 2;
    if (sc->done) return;
    const int l = threadIdx.x % LPR;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    const double2 alpha = make_double2(sc->alpha[2 * l], sc->alpha[2 * l + 1]);
    double2 acc[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3], z[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const size_t oi = o + (size_t)i * W;
            const double2 pv = ld2_stream(p + oi), qv = ld2_stream(q + oi);
            double2 xv = ld2_stream(x + oi);
            rv[i] = ld2_stream(r + oi);
            xv.x = fma(alpha.x, pv.x, xv.x); xv.y = fma(alpha.y, pv.y, xv.y);
            rv[i].x = fma(-alpha.x, qv.x, rv[i].x); rv[i].y = fma(-alpha.y, qv.y, rv[i].y);
            st2(x + oi, xv);
            st2(r + oi, rv[i]);
        }
        blk3_mul(Pinv + (size_t)9 * g, rv, z);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[0].x = fma(rv[i].x, z[i].x, acc[0].x); acc[0].y = fma(rv[i].y, z[i].y, acc[0].y);
            acc[1].x = fma(rv[i].x, rv[i].x, acc[1].x); acc[1].y = fma(rv[i].y, rv[i].y, acc[1].y);
        }
    }
    if (block_partials<W, 2>(acc, partials, counter)) {
        double s[2];
        final_sums<W, 2>(partials, s);
        __shared__ int any_active;
        __shared__ double max_rel[kMaxW];
        if (threadIdx.x == 0) any_active = 0;
        __syncthreads();
        if (threadIdx.x < W) {
            const int c = threadIdx.x;
            double rel = 0.0;
            if (sc->active[c]) {
                const double rz_old = sc->rz[c];
                sc->rr[c] = s[1];
                sc->beta[c] = (rz_old > 0.0) ? s[0] / rz_old : 0.0;
                sc->rz[c] = s[0];
                if (s[1] <= sc->rtol2 * sc->bb[c]) sc->active[c] = 0;
                else atomicOr(&any_active, 1);
            }
            if (sc->bb[c] > 0.0) rel = sc->rr[c] / sc->bb[c];
            max_rel[c] = rel;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int c = 0; c < W; ++c) m = fmax(m, max_rel[c]);
            sc->max_rel_res = sqrt(m);
            sc->iters += 1;
            sc->total_iters += 1ull;
            if (!any_active) sc->done = 1;
        }
    }
}

// p = P^-1 r + beta p   (columns that have converged keep their p; their alpha is 0)
template <int W>
__global__ void __launch_bounds__(kThreads)
cg_update_p_kernel(const double* __restrict__ r, double* __restrict__ p, const double* __restrict__ Pinv,
                   int ntriples, const CgScalars* __restrict__ sc) {
    constexpr int LPR = W / 2;
    if (sc->done) return;
    const int l = threadIdx.x % LPR;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    const double2 beta = make_double2(sc->beta[2 * l], sc->beta[2 * l + 1]);
    const bool a0 = sc->active[2 * l] != 0, a1 = sc->active[2 * l + 1] != 0;
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3], z[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) rv[i] = ld2_stream(r + o + (size_t)i * W);
        blk3_mul(Pinv + (size_t)9 * g, rv, z);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double2 pv = ld2(p + o + (size_t)i * W);
            if (a0) pv.x = fma(beta.x, pv.x, z[i].x);
            if (a1) pv.y = fma(beta.y, pv.y, z[i].y);
            st2(p + o + (size_t)i * W, pv);
        }
    }
}

// ------------------------------------------------------------------------------------
// Chebyshev three-term recurrence, fused with the accumulation into Y:
//   S_n = cs * (t_k/L) * Wv + S_{n-2} (in place in the S_{n-2} buffer),  Y += a_n S_n
// Wv = M^-1 K S_{n-1}, taken from
//   SRC 0: a panel (the CG solution x, or T for rows that need no solve)
//   SRC 1: T + Rt * x   (back-substitution rows of the condensed plan, CSR Rt)
//   SRC 2: Dblk * T     (3x3 block rows: DG1 mass inverse of the plain plan)
// The coefficient table and t_k/L live in device memory; each block stages them in shared
// memory once.
// ------------------------------------------------------------------------------------
struct ChebStep {
    int n;            // term index (coefficient a[n])
    int has_old;      // 0 for n == 1 (S_old not read)
    double cs;        // 1 for n == 1, 2 afterwards
};

template <int W, int SRC>
__global__ void __launch_bounds__(kThreads)
cheb_update_kernel(const double* __restrict__ src, const double* __restrict__ xs,
                   const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                   const double* __restrict__ blk,
                   double* __restrict__ S, double* __restrict__ Y, int nrows,
                   const double* __restrict__ coef, const double* __restrict__ tscale, ChebStep st) {
    constexpr int LPR = W / 2;
    __shared__ double s_ts[kMaxW];
    __shared__ double s_a;
    if (threadIdx.x < W) s_ts[threadIdx.x] = st.cs * tscale[threadIdx.x];
    if (threadIdx.x == 0) s_a = coef[st.n];
    __syncthreads();
    const int l = threadIdx.x % LPR;
    const double2 ts = make_double2(s_ts[2 * l], s_ts[2 * l + 1]);
    const double an = s_a;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    if (SRC == 2) {
        const int ntri = nrows / 3;
        for (int g = g0; g < ntri; g += ng) {
            const size_t o = (size_t)(3 * g) * W + 2 * l;
            double2 tv[3], wv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) tv[i] = ld2_stream(src + o + (size_t)i * W);
            blk3_mul(blk + (size_t)9 * g, tv, wv);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const size_t oi = o + (size_t)i * W;
                double2 s = st.has_old ? ld2_stream(S + oi) : make_double2(0.0, 0.0);
                s.x = fma(ts.x, wv[i].x, s.x); s.y = fma(ts.y, wv[i].y, s.y);
                double2 y = ld2_stream(Y + oi);
                y.x = fma(an, s.x, y.x); y.y = fma(an, s.y, y.y);
                st2(S + oi, s); st2(Y + oi, y);
            }
        }
    } else {
        for (int r = g0; r < nrows; r += ng) {
            const size_t o = (size_t)r * W + 2 * l;
            double2 wv = ld2_stream(src + o);
            if (SRC == 1) {
                const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
                for (int k = k0; k < k1; ++k) {
                    const double a = __ldg(val + k);
                    const double2 x = ld2(xs + (size_t)__ldg(col + k) * W + 2 * l);
                    wv.x = fma(a, x.x, wv.x); wv.y = fma(a, x.y, wv.y);
                }
            }
            double2 s = st.has_old ? ld2_stream(S + o) : make_double2(0.0, 0.0);
            s.x = fma(ts.x, wv.x, s.x); s.y = fma(ts.y, wv.y, s.y);
            double2 y = ld2_stream(Y + o);
            y.x = fma(an, s.x, y.x); y.y = fma(an, s.y, y.y);
            st2(S + o, s); st2(Y + o, y);
        }
    }
}

// Y = a_0 * S_0   (cheby_exp.py:119-123)
template <int W>
__global__ void __launch_bounds__(kThreads)
cheb_init_kernel(const double* __restrict__ S0, double* __restrict__ Y, size_t nelem2, const double* __restrict__ coef) {
    const double a0 = coef[0];
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < nelem2; i += (size_t)gridDim.x * kThreads) {
        double2 v = ld2_stream(S0 + 2 * i);
        v.x *= a0; v.y *= a0;
        st2(Y + 2 * i, v);
    }
}

// Plain W = op output without recurrence (operator_apply): rows as in cheb_update, out = Wv
template <int W, int SRC>
__global__ void __launch_bounds__(kThreads)
op_finish_kernel(const double* __restrict__ src, const double* __restrict__ xs,
                 const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                 const double* __restrict__ blk, double* __restrict__ out, int nrows) {
    constexpr int LPR = W / 2;
    const int l = threadIdx.x % LPR;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    if (SRC == 2) {
        for (int g = g0; g < nrows / 3; g += ng) {
            const size_t o = (size_t)(3 * g) * W + 2 * l;
            double2 tv[3], wv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) tv[i] = ld2(src + o + (size_t)i * W);
            blk3_mul(blk + (size_t)9 * g, tv, wv);
#pragma unroll
            for (int i = 0; i < 3; ++i) st2(out + o + (size_t)i * W, wv[i]);
        }
    } else {
        for (int r = g0; r < nrows; r += ng) {
            const size_t o = (size_t)r * W + 2 * l;
            double2 wv = ld2(src + o);
            if (SRC == 1) {
                const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
                for (int k = k0; k < k1; ++k) {
                    const double a = __ldg(val + k);
                    const double2 x = ld2(xs + (size_t)__ldg(col + k) * W + 2 * l);
                    wv.x = fma(a, x.x, wv.x); wv.y = fma(a, x.y, wv.y);
                }
            }
            st2(out + o, wv);
        }
    }
}

// ------------------------------------------------------------------------------------
// host-layout <-> panel layout.  src: B vectors of n doubles (vector k at src + k*ld),
// dst[perm[i]][k] (perm == nullptr: identity).  Columns k >= B are zero-filled.
// Tiled through shared memory so both sides are coalesced.
// ------------------------------------------------------------------------------------
template <int W>
__global__ void pack_kernel(const double* __restrict__ src, int64_t ld, int shared_x, int B, int col0,
                            const int* __restrict__ perm, double* __restrict__ dst, int n) {
    __shared__ double tile[W][33];
    const int i0 = blockIdx.x * 32;
    // read: threads along i (coalesced in the host layout)
    for (int k = threadIdx.y; k < W; k += blockDim.y) {
        const int i = i0 + threadIdx.x;
        double v = 0.0;
        if (i < n && col0 + k < B) v = src[(shared_x ? 0 : (int64_t)(col0 + k) * ld) + i];
        tile[k][threadIdx.x] = v;
    }
    __syncthreads();
    // write: threads along k (coalesced in the panel layout)
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int e = tid; e < 32 * W; e += nt) {
        const int ii = e / W, k = e % W, i = i0 + ii;
        if (i < n) dst[(size_t)(perm ? perm[i] : i) * W + k] = tile[k][ii];
    }
}

template <int W>
__global__ void unpack_kernel(const double* __restrict__ src, const int* __restrict__ perm, int n,
                              double* __restrict__ dst, int64_t ld, int B, int col0) {
    __shared__ double tile[W][33];
    const int i0 = blockIdx.x * 32;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int e = tid; e < 32 * W; e += nt) {
        const int ii = e / W, k = e % W, i = i0 + ii;
        tile[k][ii] = (i < n) ? src[(size_t)(perm ? perm[i] : i) * W + k] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < W; k += blockDim.y) {
        const int i = i0 + threadIdx.x;
        if (i < n && col0 + k < B) dst[(int64_t)(col0 + k) * ld + i] = tile[k][threadIdx.x];
    }
}

// v[i] (+)= sum_k w[k] * Y[perm[i]][k]   -- the in-GPU part of "DU *= wt; allreduce"
// (timestepping_methods.py:10-11): a weighted row reduction with warp shuffles.
template <int W>
__global__ void __launch_bounds__(kThreads)
weighted_colsum_kernel(const double* __restrict__ Y, const double* __restrict__ w, const int* __restrict__ perm,
                       int n, double* __restrict__ v, int accumulate) {
    constexpr int LPR = W / 2;
    const int l = threadIdx.x % LPR;
    const double2 wk = make_double2(w[2 * l], w[2 * l + 1]);
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    const int niter = (n + ng - 1) / ng;
    for (int it = 0; it < niter; ++it) {          // uniform trip count: shuffles stay converged
        const int i = g0 + it * ng;
        double s = 0.0;
        if (i < n) {
            const double2 y = ld2_stream(Y + (size_t)(perm ? perm[i] : i) * W + 2 * l);
            s = fma(wk.x, y.x, wk.y * y.y);
        }
#pragma unroll
        for (int off = LPR / 2; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (i < n && l == 0) v[i] = accumulate ? v[i] + s : s;
    }
}

}  // namespace kexp
