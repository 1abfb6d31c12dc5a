// kexp_kernels.cuh -- device kernels of the batched Chebyshev exp(sL)v path (sm_100a).
//
// Data layout in HBM: a batch "panel" is a row-major [rows][W] array of doubles, W in
// {2,4,8,16,32,64} sample columns contiguous per dof ("batch-minor").  A gathered row
// X[col,:] is one fully coalesced W*8-byte segment (512 B at W = 64) and every matrix
// entry is used for W columns.  All reductions are two-stage and fixed-order (warp
// shuffles -> shared memory -> per-block partials -> last block), so results are
// bit-reproducible for a given panel width and grid.
//
// Reference semantics implemented here (file:line in /root/reference):
//   operator_solver.solve()  averaged_sw_explicit.py:142-164   -> spmm + batched PCG
//   three-term recurrence    cheby_exp.py:138-165              -> cheb_* kernels
//   DU *= wt; allreduce      timestepping_methods.py:10-11     -> weighted_colsum
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kexp {

constexpr int kThreads = 256;
// A/B switches of the cooperative mass solve (defaults = measured best, profiles/r01e_summary.md)
#ifndef KEXP_GRID_FENCE
#define KEXP_GRID_FENCE 0      // 1: every thread fences again after grid.sync() (redundant: the barrier's own fence invalidates L1)
#endif
#ifndef KEXP_COOP_PREFETCH
#define KEXP_COOP_PREFETCH 16  // panels at least this wide prefetch b / x_{k-1} rows into L1 and load them after the gather (no
                               // spills); narrower (latency-bound) panels and 0 load them first
#endif
constexpr int kMaxW = 64;
constexpr int kMaxTerms = 1024;                       // rows of the per-panel coefficient table (terms 0..1023)
constexpr int kPinWeights = kMaxW * (1 + kMaxTerms);  // offset of the averaging weights in the pinned staging buffer

// Per-panel CG scalars, one slot per sample column (device memory).
struct CgScalars {
    double rz[kMaxW];      // r . P^-1 r of the current residual (measured, not recurred)
    double pq[kMaxW];      // p . A p
    double alpha[kMaxW];
    double beta[kMaxW];
    double bb[kMaxW];      // ||b||^2
    double rr[kMaxW];      // ||r||^2
    int active[kMaxW];     // column still iterating
    int done;              // all columns converged: remaining CG kernels of the graph return at once
    int iters;             // iterations of the current solve
    int max_iters;         // largest iteration count of any solve since reset
    int unconverged;       // solves that hit the iteration cap since reset
    unsigned long long total_iters;
    unsigned long long total_solves;
    double max_rel_res;    // of the last solve
    double rtol2;
    double worst_rel_res;  // largest max_rel_res of any solve that missed its tolerance since reset
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization
// attribute may start while its predecessor in the stream is still draining; everything it does
// before pdl_wait() must be independent of the predecessor's output (here: staging the constant
// matrix blocks in shared memory).  pdl_wait() returns once the predecessor has completed and its
// writes are visible.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ void prefetch_l1(const double* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// streaming variant: vectors that are touched once per kernel should not displace the
// gathered rows / matrix blocks in L1
__device__ __forceinline__ double2 ld2_stream(const double* p) {
    return __ldcs(reinterpret_cast<const double2*>(p));     // ld.global.cs: evict-first
}

// Thread -> column mapping shared by all panel kernels: a group of LPB lanes owns one row
// block (3 rows: one mesh entity) and each lane CPT = min(W, CPTMAX) sample columns as
// NV = CPT/2 double2 values; lane l holds column pairs l + LPB*j, so for fixed j the group
// reads one contiguous LPB*16-byte segment.
template <int W, int CPTMAX>
struct PanelCfg {
    static constexpr int CPT = (W >= CPTMAX) ? CPTMAX : W;
    static constexpr int LPB = W / CPT;            // lanes per row block
    static constexpr int BRW = 32 / LPB;           // row blocks per warp
    static constexpr int NV = CPT / 2;             // double2 per row per thread
    static constexpr int BPC = (kThreads / 32) * BRW;   // row blocks per CTA pass
};

// Deterministic grid-wide column sums in two ticketed levels.
// Level 1: every block writes its partial; the last block of each group of kGroup
// consecutive blocks adds the group's partials (fixed order) into a group partial.
// Level 2: the last group to finish returns true; that block then adds the group partials
// (final_sums).  The serial tail after the last block is ~gridDim/kGroup values per
// column instead of gridDim.
// Layout: partials[0 .. NACC*G*W) block partials, then NACC*NG*W group partials;
// counters[0] = level-2 ticket, counters[1 + g] = level-1 ticket of group g.
constexpr int kGroup = 8;

// Ticket part: the block partials of this block are already in global memory.
template <int W, int NACC>
__device__ __forceinline__ bool partials_tickets(double* __restrict__ partials, unsigned* __restrict__ counters) {
    __shared__ bool s_flag;
    const int G = gridDim.x, NG = (G + kGroup - 1) / kGroup;
    const int grp = blockIdx.x / kGroup;
    const int gsize = min(kGroup, G - grp * kGroup);
    double* gpart = partials + (size_t)NACC * G * W;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(counters + 1 + grp, 1u);
        s_flag = (ticket == (unsigned)gsize - 1u);
        if (s_flag) counters[1 + grp] = 0;      // ready for the next launch
    }
    __syncthreads();
    if (!s_flag) return false;
    __threadfence();
    for (int e = threadIdx.x; e < NACC * W; e += kThreads) {
        const int a = e / W, c = e % W;
        const double* base = partials + ((size_t)a * G + (size_t)grp * kGroup) * W + c;
        double v[kGroup];
#pragma unroll
        for (int u = 0; u < kGroup; ++u) v[u] = (u < gsize) ? __ldcg(base + (size_t)u * W) : 0.0;
        double s = v[0];
#pragma unroll
        for (int u = 1; u < kGroup; ++u) s += v[u];
        gpart[((size_t)a * NG + grp) * W + c] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(counters, 1u);
        s_flag = (ticket == (unsigned)NG - 1u);
        if (s_flag) counters[0] = 0;
    }
    __syncthreads();
    if (s_flag) __threadfence();
    return s_flag;
}

template <int W, int CPTMAX, int NACC>
__device__ __forceinline__ bool block_partials(double2 (&acc)[NACC][PanelCfg<W, CPTMAX>::NV],
                                               double* __restrict__ partials, unsigned* __restrict__ counters) {
    using C = PanelCfg<W, CPTMAX>;
    constexpr int LPB = C::LPB, NV = C::NV;
    __shared__ double2 sh[NACC][kThreads / 32][W / 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB;
    const int G = gridDim.x;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double2 v = acc[a][j];
#pragma unroll
            for (int off = 16; off >= LPB; off >>= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            }
            if (lane < LPB) sh[a][warp][l + LPB * j] = v;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < NACC * (W / 2); e += kThreads) {
        const int a = e / (W / 2), cp = e % (W / 2);
        double2 s = sh[a][0][cp];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) { s.x += sh[a][w][cp].x; s.y += sh[a][w][cp].y; }
        st2(partials + ((size_t)a * G + blockIdx.x) * W + 2 * cp, s);
    }
    return partials_tickets<W, NACC>(partials, counters);
}

// Last block: column sums of the group partials in fixed order.  The sum for column c of
// accumulator a is returned to the thread with threadIdx.x == c (c < W).
template <int W, int NACC>
__device__ __forceinline__ void final_sums(const double* __restrict__ partials, double (&out)[NACC]) {
    constexpr int NSPLIT = kThreads / W;
    __shared__ double red[NACC][NSPLIT][W];
    const int c = threadIdx.x % W, part = threadIdx.x / W;
    const int G = gridDim.x, NG = (G + kGroup - 1) / kGroup;
    const double* gpart = partials + (size_t)NACC * G * W;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        double s = 0.0;
        const double* base = gpart + (size_t)a * NG * W + c;
        int b = part;
        for (; b + 7 * NSPLIT < NG; b += 8 * NSPLIT) {   // 8 loads in flight, summed in order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + (size_t)(b + u * NSPLIT) * W);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < NG; b += NSPLIT) s += __ldcg(base + (size_t)b * W);
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
// Generic CSR SpMM  Y[r,:] = sum_k val[k] * X[col[k],:]   (matrices without the 3x3 entity
// structure).  One W/2-lane group per row; the group loads (col,val) pairs cooperatively
// and broadcasts them with shuffles; X rows are gathered as double2.
// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
csr_spmm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                const double* __restrict__ X, double* __restrict__ Y, int nrows) {
    constexpr int LPR = W / 2, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int l = lane % LPR, sub = lane / LPR;
    const int gwarp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kThreads) >> 5;
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
            for (int j = 0; j < cnt; ++j) {
                const int c = __shfl_sync(0xffffffffu, cc, j, LPR);
                const double a = __shfl_sync(0xffffffffu, aa, j, LPR);
                const double2 x = ld2(X + (size_t)c * W + 2 * l);
                acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y);
            }
        }
        if (r < nrows) st2(Y + (size_t)r * W + 2 * l, acc);
    }
}

// 3x3 block (row-major, 9 doubles) times three rows of NV double2 values
template <int NV>
__device__ __forceinline__ void blk3_mul(const double* __restrict__ B, const double2 (&r)[3][NV], double2 (&z)[3][NV]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double b0 = __ldg(B + 3 * i), b1 = __ldg(B + 3 * i + 1), b2 = __ldg(B + 3 * i + 2);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            z[i][j].x = fma(b0, r[0][j].x, fma(b1, r[1][j].x, b2 * r[2][j].x));
            z[i][j].y = fma(b0, r[0][j].y, fma(b1, r[1][j].y, b2 * r[2][j].y));
        }
    }
}

// ------------------------------------------------------------------------------------
// BSR(3x3) SpMM -- the layout BDM2/DG1 matrices have naturally (3 dofs per mesh entity).
// A block is 80 bytes: 9 values row-major + the block-column index in the low word of the
// 10th double.  Each CTA owns a CONTIGUOUS range of block rows; the blocks of that range
// are contiguous too and are staged in shared memory first (one coalesced pass), so the
// per-row work has a short dependency chain: block columns from shared memory -> ALL
// gathered X rows of up to GB blocks issued at once (plus the row's own p and r in the CG
// product) -> FMAs with matrix values broadcast from shared memory.  A group of W/2 lanes
// owns a block row and each lane one double2 (two sample columns), so a gathered row is
// one coalesced W*8-byte segment and neighbouring mesh entities re-use rows out of L1.
//
// MODE 0: Y = A X.
// MODE 1: the CG product q = A p with the WHOLE scalar work of the iteration fused into
//   the epilogue.  The group that owns a block row also owns its 3x3 block-Jacobi block,
//   so it can form, per column,   pq = p.q,   qz = q.P^-1 r,   qPq = q.P^-1 q.
//   The last block turns the partials into
//       alpha = rz/pq,   rz' = rz - 2 alpha qz + alpha^2 qPq,   beta = rz'/rz
//   (rz' is r'.P^-1 r' for r' = r - alpha q, exactly), so a single vector kernel can do
//   x, r, z and p afterwards: two kernels and two grid-wide reductions per iteration
//   instead of three.  rz itself is re-measured by that vector kernel every iteration,
//   so the predicted value is only ever used for one beta and errors do not accumulate.
// ------------------------------------------------------------------------------------
// Recurrence update of the solved (edge) rows fused into the LAST iteration of the mass solve, where x_K is in registers:
//   S_n = cs (t_k/L) x_K + S_{n-2}  (in place),  Y += a_n S_n      -- what cheb_update_kernel<W,0> does for these rows,
// without its launch and without re-reading x_K.  S == nullptr: not fused (plain operator application).
struct FusedUpd {
    double* S; double* Y;
    const double* coef_row;      // coefficient of this term per column (device)
    const double* tscale;        // operator scale per column (device)
    double cs; int has_old;
};

template <int W>
__device__ __forceinline__ void fused_row_update(const FusedUpd& f, size_t o, int l, double2 xn) {
    const double2 ts = make_double2(f.cs * __ldg(f.tscale + 2 * l), f.cs * __ldg(f.tscale + 2 * l + 1));
    const double2 an = make_double2(__ldg(f.coef_row + 2 * l), __ldg(f.coef_row + 2 * l + 1));
    double2 s = f.has_old ? ld2_stream(f.S + o) : make_double2(0.0, 0.0);
    s.x = fma(ts.x, xn.x, s.x); s.y = fma(ts.y, xn.y, s.y);
    double2 y = ld2_stream(f.Y + o);
    y.x = fma(an.x, s.x, y.x); y.y = fma(an.y, s.y, y.y);
    st2(f.S + o, s); st2(f.Y + o, y);
}

constexpr int kSpmmSmemBlocks = 520;      // 41 600 B of staged matrix blocks per CTA
constexpr int kSpmmSmemRows = 128;        // block rows per staged sub-chunk (row pointers + 3x3 preconditioner blocks)
constexpr int kSpmmSmemBytes = kSpmmSmemBlocks * 80 + kSpmmSmemRows * 72 + (kSpmmSmemRows + 4) * 4;
constexpr int kSpmmCpt = 2;               // PanelCfg CPTMAX of the SpMM (one double2 per lane)

#ifndef KEXP_SPMM0_MINB
#define KEXP_SPMM0_MINB 3      // resident CTAs per SM the plain product (MODE 0, T = K~ S) is compiled for: with 3 gathered
                               // blocks in flight (KEXP_GB0) it fits 80 registers; 59 -> 52 us per 32-column launch at ref 5
#endif
#ifndef KEXP_SPMM_MINB
#define KEXP_SPMM_MINB 2
#endif
template <int W, int MODE, int GB>
__global__ void __launch_bounds__(kThreads, (MODE == 0 ? KEXP_SPMM0_MINB : KEXP_SPMM_MINB))
bsr3_spmm_kernel(const int* __restrict__ browptr, const double* __restrict__ blocks,
                 const double* __restrict__ X, double* __restrict__ Y, int nbr, int max_bpr,
                 const double* __restrict__ R, const double* __restrict__ Pinv, const int* __restrict__ rowmap,
                 CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter,
                 const FusedUpd fu, int fused_row0, const double* __restrict__ fused_dblk) {
    using C = PanelCfg<W, kSpmmCpt>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    extern __shared__ double2 sblk[];        // 5 double2 per staged block
    double* sP = reinterpret_cast<double*>(sblk + 5 * kSpmmSmemBlocks);          // 9 doubles per staged block row
    int* sptr = reinterpret_cast<int*>(sP + 9 * kSpmmSmemRows);                  // row pointers relative to the sub-chunk
    bool first = true;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const int per = nbr / gridDim.x, rem = nbr % gridDim.x;
    const int start = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int end = start + per + ((int)blockIdx.x < rem ? 1 : 0);
    const int rows_sub = max(1, min(kSpmmSmemRows, kSpmmSmemBlocks / max_bpr));
    double2 dot[3][1] = {{make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0)}};

    for (int sub = start; sub < end; sub += rows_sub) {
        const int sub_end = min(sub + rows_sub, end);
        const int kb0 = __ldg(browptr + sub), kb1 = __ldg(browptr + sub_end);
        const double2* gsrc = reinterpret_cast<const double2*>(blocks) + (size_t)5 * kb0;
        for (int i = threadIdx.x; i < 5 * (kb1 - kb0); i += kThreads) sblk[i] = __ldg(gsrc + i);
        for (int i = threadIdx.x; i <= sub_end - sub; i += kThreads) sptr[i] = __ldg(browptr + sub + i) - kb0;
        if (MODE == 1)
            for (int i = threadIdx.x; i < 9 * (sub_end - sub); i += kThreads) sP[i] = __ldg(Pinv + (size_t)9 * sub + i);
        __syncthreads();
        if (first) {                 // matrix staged: from here on we read what the previous kernel wrote
            first = false;
            pdl_wait();
            if (MODE == 1 && sc->done) return;
            pdl_trigger();
        }
        for (int base = sub; base < sub_end; base += C::BPC) {
            const int br = base + warp * BRW + g;
            if (br < sub_end) {
                const int k0 = sptr[br - sub], k1 = sptr[br - sub + 1];
                const int orow = (MODE == 0 && rowmap) ? __ldg(rowmap + br) : br;      // MODE 1 matrices keep their row order
                const size_t o = (size_t)(3 * orow) * W + 2 * l;
                double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
                double2 pv[3], rv[3];
                if (MODE == 1) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) { pv[i] = ld2(X + o + (size_t)i * W); rv[i] = ld2_stream(R + o + (size_t)i * W); }
                }
                for (int kg = k0; kg < k1; kg += GB) {
                    double2 x[GB][3];
                    // (defined on every path: ptxas treats a conditionally defined array as live from the top of the kernel,
                    // around all enclosing loops, and its registers are lost everywhere -- see kexp_narrow.cu, phase A)
#pragma unroll
                    for (int j = 0; j < GB; ++j)
#pragma unroll
                        for (int r = 0; r < 3; ++r) x[j][r] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const int c = __double2loint(sblk[5 * (kg + j) + 4].y);
                            const double* xr = X + (size_t)(3 * c) * W + 2 * l;
#pragma unroll
                            for (int r = 0; r < 3; ++r) x[j][r] = ld2(xr + (size_t)r * W);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const double2* bp = sblk + 5 * (kg + j);
                            const double2 v01 = bp[0], v23 = bp[1], v45 = bp[2], v67 = bp[3], v8c = bp[4];
                            const double a[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c.x}};
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    acc[i].x = fma(a[i][r], x[j][r].x, acc[i].x);
                                    acc[i].y = fma(a[i][r], x[j][r].y, acc[i].y);
                                }
                        }
                    }
                }
                if (MODE == 0 && fu.S && 3 * orow >= fused_row0) {
                    // rows that need no solve (DG1 depth): W = T (or Dblk T), so the recurrence update happens right here
                    // and T is never written / re-read for them (MODE 0 with fu.S != nullptr: T = K~ S_{n-1} inside an apply)
                    if (fused_dblk) {
                        const double* D = fused_dblk + (size_t)9 * (orow - fused_row0 / 3);
                        const double2 t0 = acc[0], t1 = acc[1], t2 = acc[2];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double d0 = __ldg(D + 3 * i), d1 = __ldg(D + 3 * i + 1), d2 = __ldg(D + 3 * i + 2);
                            acc[i].x = fma(d0, t0.x, fma(d1, t1.x, d2 * t2.x));
                            acc[i].y = fma(d0, t0.y, fma(d1, t1.y, d2 * t2.y));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) fused_row_update<W>(fu, o + (size_t)i * W, l, acc[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) st2(Y + o + (size_t)i * W, acc[i]);
                }
                if (MODE == 1) {
                    const double* Pb = sP + 9 * (br - sub);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double b0 = Pb[3 * i], b1 = Pb[3 * i + 1], b2 = Pb[3 * i + 2];
                        double2 Pq;     // row i of P^-1 q
                        Pq.x = fma(b0, acc[0].x, fma(b1, acc[1].x, b2 * acc[2].x));
                        Pq.y = fma(b0, acc[0].y, fma(b1, acc[1].y, b2 * acc[2].y));
                        dot[0][0].x = fma(pv[i].x, acc[i].x, dot[0][0].x); dot[0][0].y = fma(pv[i].y, acc[i].y, dot[0][0].y);
                        dot[1][0].x = fma(Pq.x, rv[i].x, dot[1][0].x);     dot[1][0].y = fma(Pq.y, rv[i].y, dot[1][0].y);
                        dot[2][0].x = fma(Pq.x, acc[i].x, dot[2][0].x);    dot[2][0].y = fma(Pq.y, acc[i].y, dot[2][0].y);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (first) {
        pdl_wait();
        if (MODE == 1 && sc->done) return;
        pdl_trigger();
    }
    if (MODE == 1) {
        if (block_partials<W, kSpmmCpt, 3>(dot, partials, counter)) {
            double s[3];
            final_sums<W, 3>(partials, s);
            if (threadIdx.x < W) {
                const int c = threadIdx.x;
                const double rz = sc->rz[c];
                double alpha = 0.0, beta = 0.0;
                if (sc->active[c] && s[0] > 0.0 && rz > 0.0) {
                    alpha = rz / s[0];
                    const double rz_new = fma(alpha * alpha, s[2], fma(-2.0 * alpha, s[1], rz));
                    beta = fmax(rz_new, 0.0) / rz;
                }
                sc->pq[c] = s[0];
                sc->alpha[c] = alpha;
                sc->beta[c] = beta;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Chebyshev semi-iteration for the mass system S w = b with the 3x3 block-Jacobi P:
// the eigenvalues of P^-1 S lie in [lmin, lmax] (kappa ~ 5.2, mesh independent; estimated once
// on the host), so
//     z_k = P^-1 (b - S x_k),   x_{k+1} = x_{k-1} + omega_{k+1} (x_k + alpha z_k - x_{k-1})
// (three-term Chebyshev acceleration, alpha = 2/(lmax+lmin), omega from the host) converges
// like CG (31 vs 29 iterations for 1e-12) but needs NO inner products: the whole iteration is
// block-row local after the sparse product and runs as ONE kernel -- gather x_k, stream b and
// the own rows of x_{k-1}, overwrite them with x_{k+1} (32 n_s B bytes instead of CG's 80 n_s B
// in two kernels with two grid-wide reductions).  Neighbours only ever gather x_k, which lives
// in the other panel, so the in-place update of the x_{k-1} panel is race free.
// LAST = 1 additionally accumulates ||b - S x_k||^2 per column (the residual one step behind,
// hence conservative) for the convergence check of the fixed-length iteration.
// Structure of the product as in bsr3_spmm_kernel.
// ------------------------------------------------------------------------------------
template <int W, int LAST, int GB>
__global__ void __launch_bounds__(kThreads, KEXP_SPMM_MINB)
bsr3_cheb_iter_kernel(const int* __restrict__ browptr, const double* __restrict__ blocks,
                      const double* __restrict__ Xold, double* __restrict__ Xnew, const double* __restrict__ Bv,
                      const double* __restrict__ Pinv, int nbr, int max_bpr,
                      double omega, double alpha, int iters_done, double rtol2,
                      CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter, const FusedUpd fu) {
    using C = PanelCfg<W, kSpmmCpt>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    extern __shared__ double2 sblk[];
    double* sP = reinterpret_cast<double*>(sblk + 5 * kSpmmSmemBlocks);
    int* sptr = reinterpret_cast<int*>(sP + 9 * kSpmmSmemRows);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const int per = nbr / gridDim.x, rem = nbr % gridDim.x;
    const int start = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int end = start + per + ((int)blockIdx.x < rem ? 1 : 0);
    const int rows_sub = max(1, min(kSpmmSmemRows, kSpmmSmemBlocks / max_bpr));
    double2 rr[1][1] = {{make_double2(0.0, 0.0)}};
    bool first = true;

    for (int sub = start; sub < end; sub += rows_sub) {
        const int sub_end = min(sub + rows_sub, end);
        const int kb0 = __ldg(browptr + sub), kb1 = __ldg(browptr + sub_end);
        const double2* gsrc = reinterpret_cast<const double2*>(blocks) + (size_t)5 * kb0;
        for (int i = threadIdx.x; i < 5 * (kb1 - kb0); i += kThreads) sblk[i] = __ldg(gsrc + i);
        for (int i = threadIdx.x; i <= sub_end - sub; i += kThreads) sptr[i] = __ldg(browptr + sub + i) - kb0;
        for (int i = threadIdx.x; i < 9 * (sub_end - sub); i += kThreads) sP[i] = __ldg(Pinv + (size_t)9 * sub + i);
        __syncthreads();
        if (first) { first = false; pdl_wait(); pdl_trigger(); }
        for (int base = sub; base < sub_end; base += C::BPC) {
            const int br = base + warp * BRW + g;
            if (br < sub_end) {
                const int k0 = sptr[br - sub], k1 = sptr[br - sub + 1];
                const size_t o = (size_t)(3 * br) * W + 2 * l;
                double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
                double2 xo[3], bv[3], xm[3];          // x_k (own rows), b, x_{k-1} (own rows; Xnew still holds it)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    xo[i] = ld2(Xold + o + (size_t)i * W);
                    bv[i] = ld2_stream(Bv + o + (size_t)i * W);
                    xm[i] = ld2_stream(Xnew + o + (size_t)i * W);
                }
                for (int kg = k0; kg < k1; kg += GB) {
                    double2 x[GB][3];
                    // (defined on every path: ptxas treats a conditionally defined array as live from the top of the kernel,
                    // around all enclosing loops, and its registers are lost everywhere -- see kexp_narrow.cu, phase A)
#pragma unroll
                    for (int j = 0; j < GB; ++j)
#pragma unroll
                        for (int r = 0; r < 3; ++r) x[j][r] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const int c = __double2loint(sblk[5 * (kg + j) + 4].y);
                            const double* xr = Xold + (size_t)(3 * c) * W + 2 * l;
#pragma unroll
                            for (int r = 0; r < 3; ++r) x[j][r] = ld2(xr + (size_t)r * W);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const double2* bp = sblk + 5 * (kg + j);
                            const double2 v01 = bp[0], v23 = bp[1], v45 = bp[2], v67 = bp[3], v8c = bp[4];
                            const double a[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c.x}};
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    acc[i].x = fma(a[i][r], x[j][r].x, acc[i].x);
                                    acc[i].y = fma(a[i][r], x[j][r].y, acc[i].y);
                                }
                        }
                    }
                }
                double2 rv[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    rv[i].x = bv[i].x - acc[i].x; rv[i].y = bv[i].y - acc[i].y;
                    if (LAST) { rr[0][0].x = fma(rv[i].x, rv[i].x, rr[0][0].x); rr[0][0].y = fma(rv[i].y, rv[i].y, rr[0][0].y); }
                }
                const double* Pb = sP + 9 * (br - sub);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double2 z;
                    z.x = fma(Pb[3 * i], rv[0].x, fma(Pb[3 * i + 1], rv[1].x, Pb[3 * i + 2] * rv[2].x));
                    z.y = fma(Pb[3 * i], rv[0].y, fma(Pb[3 * i + 1], rv[1].y, Pb[3 * i + 2] * rv[2].y));
                    double2 xn;
                    xn.x = fma(omega, fma(alpha, z.x, xo[i].x) - xm[i].x, xm[i].x);
                    xn.y = fma(omega, fma(alpha, z.y, xo[i].y) - xm[i].y, xm[i].y);
                    st2(Xnew + o + (size_t)i * W, xn);
                    if (LAST && fu.S) fused_row_update<W>(fu, o + (size_t)i * W, l, xn);
                }
            }
        }
        __syncthreads();
    }
    if (first) { pdl_wait(); pdl_trigger(); }
    if (LAST) {
        if (block_partials<W, kSpmmCpt, 1>(rr, partials, counter)) {
            double s[1];
            final_sums<W, 1>(partials, s);
            __shared__ int bad;
            __shared__ double max_rel[kMaxW];
            if (threadIdx.x == 0) bad = 0;
            __syncthreads();
            if (threadIdx.x < W) {
                const int c = threadIdx.x;
                sc->rr[c] = s[0];
                double rel = 0.0;
                if (sc->bb[c] > 0.0) {
                    rel = s[0] / sc->bb[c];
                    if (s[0] > rtol2 * sc->bb[c]) atomicOr(&bad, 1);
                }
                max_rel[c] = rel;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double m = 0.0;
                for (int c = 0; c < W; ++c) m = fmax(m, max_rel[c]);
                sc->max_rel_res = sqrt(m);
                sc->iters = iters_done;
                sc->total_iters += (unsigned long long)iters_done;
                if (iters_done > sc->max_iters) sc->max_iters = iters_done;
                sc->done = bad ? 0 : 1;        // cheb_update records an unconverged solve when done == 0
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// The whole Chebyshev mass solve as ONE cooperative kernel: every CTA stages the matrix blocks,
// row pointers and preconditioner blocks of its contiguous row range in shared memory once and
// keeps them for all K-1 iterations; iterations are separated by a grid-wide barrier instead of a
// kernel boundary (no relaunch, no re-staging of the 12 MB matrix per iteration).  The barrier is
// followed by __threadfence(), whose gpu-scope fence invalidates L1 (CCTL.IVALL), so the gathered
// rows written by other SMs in the previous iteration are re-fetched while re-use inside an
// iteration still hits L1.  Needs (rows per CTA) <= kSpmmSmemRows and their blocks <= kSpmmSmemBlocks
// (checked on the host; otherwise one kernel per iteration, bsr3_cheb_iter_kernel).
// ------------------------------------------------------------------------------------
// b and x_{k-1} of the own rows are touched once per iteration (L1 is invalidated by every grid barrier): KEXP_COOP_STREAM=1
// loads them evict-first so that they do not displace the gathered rows of x_k in L1
#ifndef KEXP_COOP_STREAM
#define KEXP_COOP_STREAM 0
#endif
#if KEXP_COOP_STREAM
#define KEXP_COOP_LD ld2_stream
#else
#define KEXP_COOP_LD ld2
#endif
template <int W, int GB>
__global__ void __launch_bounds__(kThreads, KEXP_SPMM_MINB)
cheb_solve_coop_kernel(const int* __restrict__ browptr, const double* __restrict__ blocks,
                       const double* __restrict__ Bv, double* xk, double* xm,
                       const double* __restrict__ Pinv, int nbr, int K, double alpha, double mu, double rtol2, int diag_first,
                       CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter, const FusedUpd fu) {
    namespace cg = cooperative_groups;
    using C = PanelCfg<W, kSpmmCpt>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr bool kPrefetch = (KEXP_COOP_PREFETCH > 0) && (W >= KEXP_COOP_PREFETCH);
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double2 sblk[];
    double* sP = reinterpret_cast<double*>(sblk + 5 * kSpmmSmemBlocks);
    int* sptr = reinterpret_cast<int*>(sP + 9 * kSpmmSmemRows);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const int per = nbr / gridDim.x, rem = nbr % gridDim.x;
    const int start = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int end = start + per + ((int)blockIdx.x < rem ? 1 : 0);
    {
        const int kb0 = __ldg(browptr + start), kb1 = __ldg(browptr + end);
        const double2* gsrc = reinterpret_cast<const double2*>(blocks) + (size_t)5 * kb0;
        for (int i = threadIdx.x; i < 5 * (kb1 - kb0); i += kThreads) sblk[i] = __ldg(gsrc + i);
        for (int i = threadIdx.x; i <= end - start; i += kThreads) sptr[i] = __ldg(browptr + start + i) - kb0;
        for (int i = threadIdx.x; i < 9 * (end - start); i += kThreads) sP[i] = __ldg(Pinv + (size_t)9 * start + i);
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    double2 acc2[2][1] = {{make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0)}};     // ||r||^2 (last iteration), ||b||^2
    // start: x_0 = 0 (xm), x_1 = alpha P^-1 b (xk)
    for (int base = start; base < end; base += C::BPC) {
        const int br = base + warp * BRW + g;
        if (br < end) {
            const size_t o = (size_t)(3 * br) * W + 2 * l;
            double2 bv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                bv[i] = ld2(Bv + o + (size_t)i * W);
                acc2[1][0].x = fma(bv[i].x, bv[i].x, acc2[1][0].x); acc2[1][0].y = fma(bv[i].y, bv[i].y, acc2[1][0].y);
            }
            const double* Pb = sP + 9 * (br - start);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double2 z;
                z.x = alpha * fma(Pb[3 * i], bv[0].x, fma(Pb[3 * i + 1], bv[1].x, Pb[3 * i + 2] * bv[2].x));
                z.y = alpha * fma(Pb[3 * i], bv[0].y, fma(Pb[3 * i + 1], bv[1].y, Pb[3 * i + 2] * bv[2].y));
                st2(xk + o + (size_t)i * W, z);
                st2(xm + o + (size_t)i * W, make_double2(0.0, 0.0));
            }
        }
    }
    grid.sync();
    if (KEXP_GRID_FENCE) __threadfence();
    double omega = 1.0;
    for (int it = 1; it < K; ++it) {
        omega = (it == 1) ? 1.0 / (1.0 - 0.5 * mu * mu) : 1.0 / (1.0 - 0.25 * mu * mu * omega);
        const bool last = (it == K - 1);
        for (int base = start; base < end; base += C::BPC) {
            const int br = base + warp * BRW + g;
            if (br < end) {
                const int k0 = sptr[br - start], k1 = sptr[br - start + 1];
                const size_t o = (size_t)(3 * br) * W + 2 * l;
                double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
                double2 xo[3], bv[3], xmv[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) xo[i] = bv[i] = xmv[i] = make_double2(0.0, 0.0);      // defined on every path
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (!diag_first) xo[i] = ld2(xk + o + (size_t)i * W);
                    if (kPrefetch) {                       // needed only after the gather: keep the registers free
                        prefetch_l1(Bv + o + (size_t)i * W);
                        prefetch_l1(xm + o + (size_t)i * W);
                    } else {
                        bv[i] = KEXP_COOP_LD(Bv + o + (size_t)i * W);
                        xmv[i] = KEXP_COOP_LD(xm + o + (size_t)i * W);
                    }
                }
                for (int kg = k0; kg < k1; kg += GB) {
                    double2 x[GB][3];
                    // (defined on every path: ptxas treats a conditionally defined array as live from the top of the kernel,
                    // around all enclosing loops, and its registers are lost everywhere -- see kexp_narrow.cu, phase A)
#pragma unroll
                    for (int j = 0; j < GB; ++j)
#pragma unroll
                        for (int r = 0; r < 3; ++r) x[j][r] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const int c = __double2loint(sblk[5 * (kg + j) + 4].y);
                            const double* xr = xk + (size_t)(3 * c) * W + 2 * l;
#pragma unroll
                            for (int r = 0; r < 3; ++r) x[j][r] = ld2(xr + (size_t)r * W);
                        }
                    }
                    if (diag_first && kg == k0) {          // the row's own x_k: the diagonal block is gathered first
#pragma unroll
                        for (int i = 0; i < 3; ++i) xo[i] = x[0][i];
                    }
#pragma unroll
                    for (int j = 0; j < GB; ++j) {
                        if (kg + j < k1) {
                            const double2* bp = sblk + 5 * (kg + j);
                            const double2 v01 = bp[0], v23 = bp[1], v45 = bp[2], v67 = bp[3], v8c = bp[4];
                            const double a[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c.x}};
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    acc[i].x = fma(a[i][r], x[j][r].x, acc[i].x);
                                    acc[i].y = fma(a[i][r], x[j][r].y, acc[i].y);
                                }
                        }
                    }
                }
                if (kPrefetch) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        bv[i] = KEXP_COOP_LD(Bv + o + (size_t)i * W);
                        xmv[i] = KEXP_COOP_LD(xm + o + (size_t)i * W);
                    }
                }
                double2 rv[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    rv[i].x = bv[i].x - acc[i].x; rv[i].y = bv[i].y - acc[i].y;
                    if (last) { acc2[0][0].x = fma(rv[i].x, rv[i].x, acc2[0][0].x); acc2[0][0].y = fma(rv[i].y, rv[i].y, acc2[0][0].y); }
                }
                const double* Pb = sP + 9 * (br - start);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double2 z, xn;
                    z.x = fma(Pb[3 * i], rv[0].x, fma(Pb[3 * i + 1], rv[1].x, Pb[3 * i + 2] * rv[2].x));
                    z.y = fma(Pb[3 * i], rv[0].y, fma(Pb[3 * i + 1], rv[1].y, Pb[3 * i + 2] * rv[2].y));
                    xn.x = fma(omega, fma(alpha, z.x, xo[i].x) - xmv[i].x, xmv[i].x);
                    xn.y = fma(omega, fma(alpha, z.y, xo[i].y) - xmv[i].y, xmv[i].y);
                    st2(xm + o + (size_t)i * W, xn);
                    if (last && fu.S) fused_row_update<W>(fu, o + (size_t)i * W, l, xn);
                }
            }
        }
        if (!last) {
            grid.sync();
            if (KEXP_GRID_FENCE) __threadfence();
        }
        double* tsw = xk; xk = xm; xm = tsw;
    }
    if (block_partials<W, kSpmmCpt, 2>(acc2, partials, counter)) {
        double s[2];
        final_sums<W, 2>(partials, s);
        __shared__ int bad;
        __shared__ double max_rel[kMaxW];
        if (threadIdx.x == 0) bad = 0;
        __syncthreads();
        if (threadIdx.x < W) {
            const int c = threadIdx.x;
            sc->rr[c] = s[0]; sc->bb[c] = s[1];
            sc->active[c] = (s[1] > 0.0) ? 1 : 0;
            double rel = 0.0;
            if (s[1] > 0.0) {
                rel = s[0] / s[1];
                if (s[0] > rtol2 * s[1]) atomicOr(&bad, 1);
            }
            max_rel[c] = rel;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int c = 0; c < W; ++c) m = fmax(m, max_rel[c]);
            sc->max_rel_res = sqrt(m);
            sc->iters = K;
            sc->total_iters += (unsigned long long)K;
            sc->total_solves += 1ull;
            if (K > sc->max_iters) sc->max_iters = K;
            sc->done = bad ? 0 : 1;
        }
    }
}

// Chebyshev start:  x_0 = 0,  x_1 = alpha P^-1 b,  bb = ||b||^2
template <int W, int CPTMAX>
__global__ void __launch_bounds__(kThreads)
cheb_solve_init_kernel(const double* __restrict__ b, double* __restrict__ x1, double* __restrict__ x0,
                       const double* __restrict__ Pinv, int ntriples, double alpha,
                       CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    using C = PanelCfg<W, CPTMAX>;
    constexpr int LPB = C::LPB, NV = C::NV;
    pdl_wait();
    pdl_trigger();
    const int l = threadIdx.x % LPB;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPB;
    const int ng = gridDim.x * kThreads / LPB;
    double2 acc[1][NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[0][j] = make_double2(0.0, 0.0);
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3][NV], z[3][NV];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) rv[i][j] = ld2_stream(b + o + (size_t)i * W + 2 * LPB * j);
        blk3_mul<NV>(Pinv + (size_t)9 * g, rv, z);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t oi = o + (size_t)i * W + 2 * LPB * j;
                z[i][j].x *= alpha; z[i][j].y *= alpha;
                st2(x1 + oi, z[i][j]);
                st2(x0 + oi, make_double2(0.0, 0.0));
                acc[0][j].x = fma(rv[i][j].x, rv[i][j].x, acc[0][j].x); acc[0][j].y = fma(rv[i][j].y, rv[i][j].y, acc[0][j].y);
            }
    }
    if (block_partials<W, CPTMAX, 1>(acc, partials, counter)) {
        double s[1];
        final_sums<W, 1>(partials, s);
        if (threadIdx.x < W) {
            const int c = threadIdx.x;
            sc->bb[c] = s[0]; sc->rr[c] = s[0];
            sc->active[c] = (s[0] > 0.0) ? 1 : 0;
        }
        if (threadIdx.x == 0) {
            sc->done = 1;          // set to 0 by the last iteration if the residual test fails
            sc->iters = 0;
            sc->max_rel_res = 0.0;
            sc->total_solves += 1ull;
        }
    }
}

// ------------------------------------------------------------------------------------
// PCG start:  x = 0, r = b (in place: r aliases b), p = P^-1 r, rz = r.p, bb = r.r
// ------------------------------------------------------------------------------------
template <int W, int CPTMAX>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(const double* __restrict__ r, double* __restrict__ x, double* __restrict__ p,
               const double* __restrict__ Pinv, int ntriples,
               CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    using C = PanelCfg<W, CPTMAX>;
    constexpr int LPB = C::LPB, NV = C::NV;
    pdl_wait();
    pdl_trigger();
    const int l = threadIdx.x % LPB;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPB;
    const int ng = gridDim.x * kThreads / LPB;
    double2 acc[2][NV];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[a][j] = make_double2(0.0, 0.0);
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3][NV], z[3][NV];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) rv[i][j] = ld2_stream(r + o + (size_t)i * W + 2 * LPB * j);
        blk3_mul<NV>(Pinv + (size_t)9 * g, rv, z);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t oi = o + (size_t)i * W + 2 * LPB * j;
                st2(x + oi, make_double2(0.0, 0.0));
                st2(p + oi, z[i][j]);
                acc[0][j].x = fma(rv[i][j].x, z[i][j].x, acc[0][j].x); acc[0][j].y = fma(rv[i][j].y, z[i][j].y, acc[0][j].y);
                acc[1][j].x = fma(rv[i][j].x, rv[i][j].x, acc[1][j].x); acc[1][j].y = fma(rv[i][j].y, rv[i][j].y, acc[1][j].y);
            }
    }
    if (block_partials<W, CPTMAX, 2>(acc, partials, counter)) {
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
// PCG vector update, all of it in one pass (alpha and beta are both known, see above):
//   x += alpha p,  r -= alpha q,  z = P^-1 r,  p = z + beta p;   rz = r.z,  rr = r.r
// last block: convergence per column, rz <- measured r.z.
// Columns that have converged are frozen (alpha = 0, p untouched), so a column's result
// does not depend on which other columns share its batch.
// ------------------------------------------------------------------------------------
template <int W, int CPTMAX>
__global__ void __launch_bounds__(kThreads)
cg_update_kernel(double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                 const double* __restrict__ q, const double* __restrict__ Pinv, int ntriples,
                 CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    using C = PanelCfg<W, CPTMAX>;
    constexpr int LPB = C::LPB, NV = C::NV;
    pdl_wait();
    if (sc->done) return;
    pdl_trigger();
    const int l = threadIdx.x % LPB;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPB;
    const int ng = gridDim.x * kThreads / LPB;
    double2 alpha[NV], beta[NV];
    bool act0[NV], act1[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = 2 * (l + LPB * j);
        alpha[j] = make_double2(sc->alpha[c], sc->alpha[c + 1]);
        beta[j] = make_double2(sc->beta[c], sc->beta[c + 1]);
        act0[j] = sc->active[c] != 0; act1[j] = sc->active[c + 1] != 0;
    }
    double2 acc[2][NV];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[a][j] = make_double2(0.0, 0.0);
    for (int g = g0; g < ntriples; g += ng) {
        const size_t o = (size_t)(3 * g) * W + 2 * l;
        double2 rv[3][NV], pv[3][NV], qv[3][NV], xv[3][NV], z[3][NV];
        // phase 1: every load of the triple is issued before anything is stored (12 x 16 B
        // in flight per thread); the compiler will not hoist loads over the in-place stores
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t oi = o + (size_t)i * W + 2 * LPB * j;
                pv[i][j] = ld2_stream(p + oi);
                qv[i][j] = ld2_stream(q + oi);
                xv[i][j] = ld2_stream(x + oi);
                rv[i][j] = ld2_stream(r + oi);
            }
        double Pb[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) Pb[e] = __ldg(Pinv + (size_t)9 * g + e);
        // phase 2: arithmetic
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                xv[i][j].x = fma(alpha[j].x, pv[i][j].x, xv[i][j].x); xv[i][j].y = fma(alpha[j].y, pv[i][j].y, xv[i][j].y);
                rv[i][j].x = fma(-alpha[j].x, qv[i][j].x, rv[i][j].x); rv[i][j].y = fma(-alpha[j].y, qv[i][j].y, rv[i][j].y);
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                z[i][j].x = fma(Pb[3 * i], rv[0][j].x, fma(Pb[3 * i + 1], rv[1][j].x, Pb[3 * i + 2] * rv[2][j].x));
                z[i][j].y = fma(Pb[3 * i], rv[0][j].y, fma(Pb[3 * i + 1], rv[1][j].y, Pb[3 * i + 2] * rv[2][j].y));
                acc[0][j].x = fma(rv[i][j].x, z[i][j].x, acc[0][j].x); acc[0][j].y = fma(rv[i][j].y, z[i][j].y, acc[0][j].y);
                acc[1][j].x = fma(rv[i][j].x, rv[i][j].x, acc[1][j].x); acc[1][j].y = fma(rv[i][j].y, rv[i][j].y, acc[1][j].y);
                if (act0[j]) pv[i][j].x = fma(beta[j].x, pv[i][j].x, z[i][j].x);
                if (act1[j]) pv[i][j].y = fma(beta[j].y, pv[i][j].y, z[i][j].y);
            }
        // phase 3: stores
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const size_t oi = o + (size_t)i * W + 2 * LPB * j;
                st2(x + oi, xv[i][j]);
                st2(r + oi, rv[i][j]);
                st2(p + oi, pv[i][j]);
            }
    }
    if (block_partials<W, CPTMAX, 2>(acc, partials, counter)) {
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
                sc->rr[c] = s[1];
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
            if (sc->iters > sc->max_iters) sc->max_iters = sc->iters;
            if (!any_active) sc->done = 1;
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
// The coefficient table (one row of kMaxW per term: every column may carry its own series, see
// the economised per-column expansions in kexp_api.cu) and the per-column operator scale live in
// device memory; each block stages its row in shared memory once.
// ------------------------------------------------------------------------------------
struct ChebStep {
    int n;            // term index (coefficient a[n])
    int has_old;      // 0 for n == 1 (S_old not read)
    int check_cg;     // 1: this launch follows a CG loop; record a solve that hit the cap
    double cs;        // 1 for n == 1, 2 afterwards
};

template <int W, int SRC>
__global__ void __launch_bounds__(kThreads)
cheb_update_kernel(const double* __restrict__ src, const double* __restrict__ xs,
                   const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                   const double* __restrict__ blk,
                   double* __restrict__ S, double* __restrict__ Y, int nrows,
                   const double* __restrict__ coef, const double* __restrict__ tscale, ChebStep st,
                   CgScalars* __restrict__ sc) {
    constexpr int LPR = W / 2;
    __shared__ double s_ts[kMaxW];
    __shared__ double s_a[kMaxW];
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x < W) {
        s_ts[threadIdx.x] = st.cs * tscale[threadIdx.x];
        s_a[threadIdx.x] = coef[(size_t)st.n * kMaxW + threadIdx.x];      // per-column series coefficient of term n
    }
    if (threadIdx.x == 0) {
        if (st.check_cg && blockIdx.x == 0 && !sc->done) {
            sc->unconverged += 1;
            sc->worst_rel_res = fmax(sc->worst_rel_res, sc->max_rel_res);
        }
    }
    __syncthreads();
    const int l = threadIdx.x % LPR;
    const double2 ts = make_double2(s_ts[2 * l], s_ts[2 * l + 1]);
    const double2 an = make_double2(s_a[2 * l], s_a[2 * l + 1]);
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    if (SRC == 2) {
        const int ntri = nrows / 3;
        for (int g = g0; g < ntri; g += ng) {
            const size_t o = (size_t)(3 * g) * W + 2 * l;
            double2 tv[3][1], wv[3][1];
#pragma unroll
            for (int i = 0; i < 3; ++i) tv[i][0] = ld2_stream(src + o + (size_t)i * W);
            blk3_mul<1>(blk + (size_t)9 * g, tv, wv);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const size_t oi = o + (size_t)i * W;
                double2 s = st.has_old ? ld2_stream(S + oi) : make_double2(0.0, 0.0);
                s.x = fma(ts.x, wv[i][0].x, s.x); s.y = fma(ts.y, wv[i][0].y, s.y);
                double2 y = ld2_stream(Y + oi);
                y.x = fma(an.x, s.x, y.x); y.y = fma(an.y, s.y, y.y);
                st2(S + oi, s); st2(Y + oi, y);
            }
        }
    } else {
        for (int r = g0; r < nrows; r += ng) {
            const size_t o = (size_t)r * W + 2 * l;
            double2 wv = ld2_stream(src + o);
            if (SRC == 1) {
                // back-substitution row: 9 entries (an interior dof against the 9 edge dofs of its cell); all gathers of a
                // chunk are issued before the first use -- one L2 round trip per row instead of one per entry
                constexpr int CH = 9;
                const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
                for (int k = k0; k < k1; k += CH) {
                    double2 x[CH];
                    double a[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) { x[j] = make_double2(0.0, 0.0); a[j] = 0.0; }      // defined on every path
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (k + j < k1) {
                            a[j] = __ldg(val + k + j);
                            x[j] = ld2(xs + (size_t)__ldg(col + k + j) * W + 2 * l);
                        }
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (k + j < k1) { wv.x = fma(a[j], x[j].x, wv.x); wv.y = fma(a[j], x[j].y, wv.y); }
                }
            }
            double2 s = st.has_old ? ld2_stream(S + o) : make_double2(0.0, 0.0);
            s.x = fma(ts.x, wv.x, s.x); s.y = fma(ts.y, wv.y, s.y);
            double2 y = ld2_stream(Y + o);
            y.x = fma(an.x, s.x, y.x); y.y = fma(an.y, s.y, y.y);
            st2(S + o, s); st2(Y + o, y);
        }
    }
}

// Y = a_0 * S_0   (cheby_exp.py:119-123)
template <int W>
__global__ void __launch_bounds__(kThreads)
cheb_init_kernel(const double* __restrict__ S0, double* __restrict__ Y, size_t nelem2, const double* __restrict__ coef) {
    pdl_wait();
    pdl_trigger();
    constexpr int LPR = W / 2;
    // kThreads and the grid stride are multiples of LPR: a thread keeps its column pair
    const int l = threadIdx.x % LPR;
    const double2 a0 = make_double2(coef[2 * l], coef[2 * l + 1]);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < nelem2; i += (size_t)gridDim.x * kThreads) {
        double2 v = ld2_stream(S0 + 2 * i);
        v.x *= a0.x; v.y *= a0.y;
        st2(Y + 2 * i, v);
    }
}

// ------------------------------------------------------------------------------------
// Shared-input batches (the forward half of an averaged evaluation: every ensemble member starts
// from the same U, timestepping_methods.py:7).  All columns then live in one polynomial space:
// with U_m = S_m((t*/L) A) x for t* = max|t_k|, and s_k = t_k/t*,
//     y_k = sum_n a_n S_n(s_k (t*/L) A) x = sum_m D[m][k] U_m,
// because S_n(s b) = sum_m E_nm(s) S_m(b) with |E_nm| <= 1 for |s| <= 1 (host recurrence).
// So the recurrence runs for ONE column and the 64 results are a small dense combination.
// basis_store: column 0 of a panel -> basis[:, m];  basis_combine: Y[i,k] = sum_m basis[i][m] D[m][k].
// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
basis_store_kernel(const double* __restrict__ S, int n, double* __restrict__ basis, int ld, int m) {
    pdl_wait();
    pdl_trigger();
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
        basis[(size_t)i * ld + m] = S[(size_t)i * W];
}

// Long series are combined in chunks of kCombineChunk terms (the coefficient tile lives in shared memory):
// `basis` and `D` then point at the chunk's first term and accumulate != 0 adds to Y.
constexpr int kCombineChunk = 256;
template <int W>
__global__ void __launch_bounds__(kThreads)
basis_combine_kernel(const double* __restrict__ basis, int ld, int nb, const double* __restrict__ D,
                     double* __restrict__ Y, int n, int accumulate) {
    constexpr int LPR = W / 2;
    extern __shared__ double sD[];                   // nb x W
    for (int e = threadIdx.x; e < nb * W; e += kThreads) sD[e] = D[e];
    __syncthreads();
    const int l = threadIdx.x % LPR;
    const int g0 = (blockIdx.x * kThreads + threadIdx.x) / LPR;
    const int ng = gridDim.x * kThreads / LPR;
    for (int i = g0; i < n; i += ng) {
        const double* bi = basis + (size_t)i * ld;
        double2 acc = accumulate ? ld2(Y + (size_t)i * W + 2 * l) : make_double2(0.0, 0.0);
        for (int m = 0; m < nb; ++m) {
            const double u = __ldg(bi + m);
            acc.x = fma(u, sD[m * W + 2 * l], acc.x);
            acc.y = fma(u, sD[m * W + 2 * l + 1], acc.y);
        }
        st2(Y + (size_t)i * W + 2 * l, acc);
    }
}

// Convergence record of a mass solve that is not followed by a recurrence update (kexp_operator_apply_dev)
static __global__ void record_solve_kernel(CgScalars* __restrict__ sc) {
    if (!sc->done) {
        sc->unconverged += 1;
        sc->worst_rel_res = fmax(sc->worst_rel_res, sc->max_rel_res);
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
            double2 tv[3][1], wv[3][1];
#pragma unroll
            for (int i = 0; i < 3; ++i) tv[i][0] = ld2(src + o + (size_t)i * W);
            blk3_mul<1>(blk + (size_t)9 * g, tv, wv);
#pragma unroll
            for (int i = 0; i < 3; ++i) st2(out + o + (size_t)i * W, wv[i][0]);
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
    for (int k = threadIdx.y; k < W; k += blockDim.y) {
        const int i = i0 + threadIdx.x;
        double v = 0.0;
        if (i < n && col0 + k < B) v = src[(shared_x ? 0 : (int64_t)(col0 + k) * ld) + i];
        tile[k][threadIdx.x] = v;
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int e = tid; e < 32 * W; e += nt) {
        const int ii = e / W, k = e % W, i = i0 + ii;
        if (i < n) dst[(size_t)(perm ? perm[i] : i) * W + k] = tile[k][ii];
    }
}

// ------------------------------------------------------------------------------------
// Fused weighted average (the adjoint of the shared-input trick above):
//     v = sum_k w_k exp(t_k L) X_k = sum_k w_k sum_m b_m^(k) S_m(C) X_k = sum_m S_m(C) z_m,
//     z_m = sum_k w_k b_m^(k) X_k,
// because after the re-expansion every column is a polynomial in the SAME operator C.  The B input
// columns are first combined into the M+1 coefficient vectors z_m (clenshaw_combine: a small dense
// product, one pass over the inputs), then ONE Clenshaw recurrence
//     u_m = z_m + 2 C u_{m+1} + u_{m+2},      v = z_0 + C u_1 + u_2
// runs on a single column (basis_axpy adds z_m, the usual operator + update kernels do the rest).
// The B individual results exp(t_k L) X_k are never formed (timestepping_methods.py:9-11 only uses
// their weighted sum).
// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
clenshaw_combine_kernel(const double* __restrict__ X, int n, const double* __restrict__ G, int nb,
                        double* __restrict__ Z, int ld, int accumulate) {
    extern __shared__ double s_G[];                 // nb x W:  G[m][c] = w_c b_m^(c)
    pdl_wait();
    pdl_trigger();
    for (int i = threadIdx.x; i < nb * W; i += kThreads) s_G[i] = G[i];
    __syncthreads();
    constexpr int CH = (W < 16) ? W : 16;           // columns held in registers at a time
    for (int r = blockIdx.x * kThreads + threadIdx.x; r < n; r += gridDim.x * kThreads) {
        double* z = Z + (size_t)r * ld;
#pragma unroll
        for (int c0 = 0; c0 < W; c0 += CH) {
            double x[CH];
#pragma unroll
            for (int c = 0; c < CH; c += 2) {
                const double2 v = ld2_stream(X + (size_t)r * W + c0 + c);
                x[c] = v.x; x[c + 1] = v.y;
            }
            for (int m = 0; m < nb; ++m) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < CH; ++c) acc = fma(x[c], s_G[m * W + c0 + c], acc);
                z[m] = (accumulate || c0 > 0) ? z[m] + acc : acc;
            }
        }
    }
}

// S[:, 0] += Z[:, m]   (column 0 of a panel of width W)
template <int W>
__global__ void __launch_bounds__(kThreads)
basis_axpy_kernel(double* __restrict__ S, int n, const double* __restrict__ Z, int ld, int m) {
    pdl_wait();
    pdl_trigger();
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
        S[(size_t)i * W] += Z[(size_t)i * ld + m];
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
