// kexp_narrow.cu -- the whole cheby_exp.apply (cheby_exp.py:110-165) of a NARROW panel (2..16 sample columns) as ONE
// persistent cooperative kernel.
//
// Why: with few columns per GPU -- the strong-scaling regime (64 samples over 8 GPUs = 8 columns each), the
// single-column recurrences of the shared-input forward pass and of the fused Clenshaw average, and the reference's
// own un-batched apply(x, y, t) -- an apply is a chain of ~25 terms x ~22 sparse products with a grid-wide dependency
// after each one, and its cost is LATENCY (L2 round trip of the gather + barrier), not bandwidth.  Five kernel
// launches per term (K~ product, mass solve, three row ranges of the recurrence update) each pay a launch boundary,
// re-stage the mass matrix in shared memory and serialise the gathers of a block row.  Here:
//   * one launch per apply: 1 CTA per SM, 512 threads, matrix S (+ the index arrays of K~ and R) of the CTA's row
//     ranges resident in shared memory for ALL terms and iterations;
//   * phases separated by grid-wide barriers only:   per term   A: T = K~ S_{n-1} (depth rows updated in the epilogue)
//     | B: K-1 Chebyshev iterations of the mass solve, the edge-row recurrence update fused into the last one
//     | C: interior rows  w_I = T_I + R w_E  and their update   -> K + 2 barriers per term;
//   * inside the mass solve the iterate lives in SHARED memory: x_k for the CTA's own rows plus the halo rows its
//     blocks reference (local slot numbers precomputed on the host), x_{k-1} / x_{k+1} for the own rows, and b.  Per
//     iteration a CTA reads only its halo (each row once, coalesced, ld.global.cg) and publishes its own rows; the
//     products read shared memory, so there is no gather latency chain, no L1 coherence question and hardly any
//     register pressure;
//   * data written by other CTAs is read with ld.global.cg (L2, the coherence point), everything else normally;
//   * the convergence record of every mass solve (||b - S x||/||b|| per column against the relaxed tolerance of
//     its term) is reduced in fixed order by CTA 0 behind the barrier that follows the solve.
// The arithmetic per row is the same as in the wide-panel kernels (kexp_kernels.cuh): same three-term Chebyshev
// acceleration, same recurrence, same order of operations inside a row.
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "kexp_internal.h"

namespace cg = cooperative_groups;

namespace kexp {

template <int W> struct NarrowCfg {
    static constexpr int NT = 512;                   // 1 CTA per SM, 128 registers per thread
    static constexpr int GK = 4;                     // gathered blocks of K~ in flight per row
};
constexpr int kNTmax = 512;
constexpr int kNWarpsMax = kNTmax / 32;
constexpr int kRChunk = 9;               // nonzeros of a row of R in flight (interior dof x the 9 edge dofs of its cell)

struct NarrowArgs {
    // K~ (BSR 3x3, stored rows in locality order, rowmap = output block row)
    const int* kt_browptr; const double* kt_blocks; const int* kt_rowmap; int kt_nbr;
    // S (BSR 3x3, diagonal block first), block-Jacobi inverse
    const int* s_browptr; const double* s_blocks; const double* Pinv; int s_nbr;
    // s_blocks here is the LOCALLY indexed copy: the 10th double of a block holds the shared-memory slot of its column
    // (own rows first, then the CTA's halo rows); halo_ptr[c] .. halo_ptr[c+1]: global block rows of CTA c's halo
    const int* halo_ptr; const int* halo_cols;
    // R (CSR, n_b x n_s) or null; Dblk (n_d/3 blocks) or null
    const int* r_rowptr; const int* r_col; const double* r_val; const double* Dblk;
    int n, n_s, n_b, n_d;
    // panels [rows][W]
    double *SA, *SB, *Y, *T, *X, *P;
    // per-panel tables (device): tscale[W], coef[(M+1)][kMaxW]; per-term iteration counts and squared tolerances
    const double* tscale; const double* coef; const int* kterm; const double* rtol2;
    int M;                               // highest term index
    int mode;                            // 0: forward recurrence y = sum a_n S_n;  1: Clenshaw (fused weighted average)
    double* basis; int basis_ld;         // mode 0: store column 0 of S_n (or null);  mode 1: z_m (required)
    double alpha, mu;                    // Chebyshev acceleration of the mass solve
    CgScalars* sc; double* partials;     // [gridDim][2][W] block sums of ||r||^2, ||b||^2
    unsigned* barrier;                   // grid barrier counter (zeroed before the launch)
    // shared-memory layout (element counts, maxima over the CTAs)
    int cap_s_blocks, cap_s_rows, cap_kt_rows, cap_kt_blocks, cap_r_rows, cap_r_nnz, cap_slots, cap_halo;
    int xs_offset;                       // first double of the iterate buffers in dynamic shared memory
};

// The arguments live in constant memory (one slot per context, round-robin): the phases below are separate, NOT
// inlined functions -- each gets its own register allocation instead of one body that keeps the loop invariants of all
// phases alive and spills -- and read their arguments from the constant bank.  (Passing the ADDRESS of a by-value
// kernel parameter to out-of-line functions makes the compiler keep a per-thread copy in local memory: every pointer
// load of the hot loops became a dependent L1/L2 access, measured 4x slower.)
constexpr int kArgSlots = 8;
__constant__ NarrowArgs c_nargs[kArgSlots];

// Grid-wide barrier of the persistent kernel.  Every value that crosses CTAs is read with ld.global.cg (L2, the coherence
// point), so the barrier does NOT need the L1 invalidation that cooperative_groups' grid.sync() performs (CCTL.IVALL behind
// its acquire) -- and must not have it: the invalidation also drops the thread-private stack lines, after which every
// spilled value of the loop state costs an L2 round trip per phase.  Release: bar.sync + one gpu-scope fence + atomic
// arrive by thread 0; wait: volatile poll of the monotonic counter (zeroed at launch), then bar.sync.
#ifndef KEXP_NARROW_SYNC
#define KEXP_NARROW_SYNC 1         // 0: cooperative_groups grid.sync()
#endif
#ifndef KEXP_NARROW_INLINE
#define KEXP_NARROW_INLINE 0       // 1: phases inlined into one kernel body
#endif
#if KEXP_NARROW_INLINE
#define KEXP_PHASE_FN __device__ __forceinline__
#else
#define KEXP_PHASE_FN __device__ __noinline__
#endif
__device__ __forceinline__ void narrow_barrier(cg::grid_group& grid, unsigned* counter, unsigned& target) {
#if KEXP_NARROW_SYNC
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
    }
    __syncthreads();
#else
    grid.sync();
#endif
}

__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ void cta_range(int nrows, int& start, int& end) {
    const int per = nrows / gridDim.x, rem = nrows % gridDim.x;
    start = blockIdx.x * per + min((int)blockIdx.x, rem);
    end = start + per + ((int)blockIdx.x < rem ? 1 : 0);
}

// recurrence update of one row segment (two columns of a row):  s = cs*ts*w + s_old (+ z),  y += a*s
struct UpdCtx {
    double2 ts, an;          // cs * tscale, coefficient of this term (per column pair)
    int has_old, mode, first_lane, term;
    double* Snew; double* Y; const double* zsrc; double* bstore; int ld;
};

__device__ __forceinline__ void row_update(const UpdCtx& u, size_t o /* element offset of the pair */, size_t row, double2 w) {
    double2 s = u.has_old ? *reinterpret_cast<const double2*>(u.Snew + o) : make_double2(0.0, 0.0);
    if (u.mode == 1 && u.first_lane) s.x += __ldg(u.zsrc + row * u.ld);
    s.x = fma(u.ts.x, w.x, s.x); s.y = fma(u.ts.y, w.y, s.y);
    *reinterpret_cast<double2*>(u.Snew + o) = s;
    if (u.mode == 0) {
        double2 y = *reinterpret_cast<const double2*>(u.Y + o);
        y.x = fma(u.an.x, s.x, y.x); y.y = fma(u.an.y, s.y, y.y);
        *reinterpret_cast<double2*>(u.Y + o) = y;
        if (u.bstore && u.first_lane) u.bstore[row * u.ld] = s.x;
    }
}

// Per-step constants shared by the three phases.
struct StepCtx {
    double* Sin; double* Snew;           // S_{n-1};  S_{n-2} -> S_n
    double cs; int has_old, term, step;
    const double* crow;                  // coefficient row of this term
};

template <int W>
__device__ __forceinline__ UpdCtx make_upd(const NarrowArgs& a, const StepCtx& sx, int lane_pair) {   // a: constant bank
    UpdCtx u;
    u.has_old = sx.has_old; u.mode = a.mode; u.term = sx.term; u.Snew = sx.Snew; u.Y = a.Y; u.ld = a.basis_ld;
    u.zsrc = (a.mode == 1) ? a.basis + sx.term : nullptr;
    u.bstore = (a.mode == 0 && a.basis) ? a.basis + sx.term : nullptr;
    u.ts = make_double2(sx.cs * a.tscale[2 * lane_pair], sx.cs * a.tscale[2 * lane_pair + 1]);
    u.an = make_double2(__ldg(sx.crow + 2 * lane_pair), __ldg(sx.crow + 2 * lane_pair + 1));
    u.first_lane = (lane_pair == 0);
    return u;
}

// ---- phase A: T = K~ S_{n-1}; depth rows (no solve) take their recurrence update in the epilogue
template <int W>
KEXP_PHASE_FN void narrow_phase_a(int slot, const StepCtx sx, const int* __restrict__ kptr, const int* __restrict__ kcol,
                                         int k0r, int k1r, int kt_b0) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int kNT = NarrowCfg<W>::NT, kNWarps = kNT / 32, kGatherK = NarrowCfg<W>::GK;
    constexpr int BPC = kNWarps * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const UpdCtx u = make_upd<W>(a, sx, l);
    const double* Sin = sx.Sin;
    const int row_eta = a.n_s + a.n_b;
    for (int base = k0r; base < k1r; base += BPC) {
        const int br = base + warp * BRW + g;
        if (br < k1r) {
            const int q0 = kptr[br - k0r], q1 = kptr[br - k0r + 1];
            const int orow = a.kt_rowmap ? __ldg(a.kt_rowmap + br) : br;
            double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
            for (int q = q0; q < q1; q += kGatherK) {
                double2 x[kGatherK][3];
#pragma unroll
                for (int j = 0; j < kGatherK; ++j)
                    if (q + j < q1) {
                        const double* xr = Sin + (size_t)(3 * kcol[q + j]) * W + 2 * l;
#pragma unroll
                        for (int r = 0; r < 3; ++r) x[j][r] = ldcg2(xr + (size_t)r * W);
                    }
#pragma unroll
                for (int j = 0; j < kGatherK; ++j)
                    if (q + j < q1) {
                        const double2* bp = reinterpret_cast<const double2*>(a.kt_blocks) + (size_t)5 * (kt_b0 + q + j);
                        const double2 v01 = __ldg(bp), v23 = __ldg(bp + 1), v45 = __ldg(bp + 2), v67 = __ldg(bp + 3), v8c = __ldg(bp + 4);
                        const double m[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c.x}};
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                acc[i].x = fma(m[i][r], x[j][r].x, acc[i].x);
                                acc[i].y = fma(m[i][r], x[j][r].y, acc[i].y);
                            }
                    }
            }
            const size_t o = (size_t)(3 * orow) * W + 2 * l;
            if (3 * orow >= row_eta) {
                if (a.Dblk) {                      // plain plan: W_eta = M_eta^-1 T_eta (3x3 block)
                    const double* D = a.Dblk + (size_t)9 * (orow - row_eta / 3);
                    const double2 t3[3] = {acc[0], acc[1], acc[2]};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double d0 = __ldg(D + 3 * i), d1 = __ldg(D + 3 * i + 1), d2 = __ldg(D + 3 * i + 2);
                        acc[i].x = fma(d0, t3[0].x, fma(d1, t3[1].x, d2 * t3[2].x));
                        acc[i].y = fma(d0, t3[0].y, fma(d1, t3[1].y, d2 * t3[2].y));
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) row_update(u, o + (size_t)i * W, (size_t)(3 * orow + i), acc[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) *reinterpret_cast<double2*>(a.T + o + (size_t)i * W) = acc[i];
            }
        }
    }
}

// Shared-memory state of the mass solve (per CTA): xa = x_k on [own rows | halo rows], xb = x_{k-1} on the own rows (it
// becomes x_{k+1}; its halo part is filled after the next barrier), bs = b on the own rows.  Row r of a slot starts at
// (3*slot + r) * W doubles.
// ---- phase B, start: x_0 = 0, x_1 = alpha P^-1 b on the own rows (shared memory + published); returns ||b||^2 share
template <int W>
KEXP_PHASE_FN double2 narrow_solve_start(int slot, double* xg, double* xa, double* xb, double* bs, const double* __restrict__ sP, int s0, int s1) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int BPC = (NarrowCfg<W>::NT / 32) * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    double2 accb = make_double2(0.0, 0.0);
    for (int base = s0; base < s1; base += BPC) {
        const int br = base + warp * BRW + g;
        if (br < s1) {
            const size_t o = (size_t)(3 * br) * W + 2 * l;
            const int ls = (3 * (br - s0)) * W + 2 * l;
            double2 bv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                bv[i] = ldcg2(a.T + o + (size_t)i * W);
                *reinterpret_cast<double2*>(bs + ls + i * W) = bv[i];
                accb.x = fma(bv[i].x, bv[i].x, accb.x); accb.y = fma(bv[i].y, bv[i].y, accb.y);
            }
            const double* Pb = sP + 9 * (br - s0);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double2 z;
                z.x = a.alpha * fma(Pb[3 * i], bv[0].x, fma(Pb[3 * i + 1], bv[1].x, Pb[3 * i + 2] * bv[2].x));
                z.y = a.alpha * fma(Pb[3 * i], bv[0].y, fma(Pb[3 * i + 1], bv[1].y, Pb[3 * i + 2] * bv[2].y));
                *reinterpret_cast<double2*>(xa + ls + i * W) = z;
                *reinterpret_cast<double2*>(xb + ls + i * W) = make_double2(0.0, 0.0);
                *reinterpret_cast<double2*>(xg + o + (size_t)i * W) = z;
            }
        }
    }
    return accb;
}

// ---- phase B, halo: the rows of x_k this CTA's blocks reference outside its own range, global -> shared (each row once,
// 16-byte pieces, coalesced within a row)
template <int W>
KEXP_PHASE_FN void narrow_load_halo(const double* xg, double* xa, const int* __restrict__ hcols, int n_own, int n_halo) {
    constexpr int PPR = 3 * W / 2;                     // 16-byte pieces per block row (3 rows x W doubles, contiguous)
    constexpr int NT = NarrowCfg<W>::NT;
    for (int e = threadIdx.x; e < n_halo * PPR; e += NT) {
        const int h = e / PPR, p = e % PPR;
        const double2 v = ldcg2(xg + (size_t)(3 * hcols[h]) * W + 2 * p);
        *reinterpret_cast<double2*>(xa + (size_t)(3 * (n_own + h)) * W + 2 * p) = v;
    }
}

// ---- phase B, one Chebyshev iteration on the own rows, everything read from shared memory: x_{k+1} overwrites x_{k-1}
// (shared) and is published to xg_out; LAST also forms ||b - S x_k||^2 and the recurrence update of the edge rows (W = x_K)
template <int W, bool LAST>
KEXP_PHASE_FN double2 narrow_solve_iter(int slot, const StepCtx sx, const double* xa, double* xb, const double* bs, double* xg_out, double omega,
                                        const double2* __restrict__ sblk, const double* __restrict__ sP, const int* __restrict__ sptr,
                                        int s0, int s1) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int BPC = (NarrowCfg<W>::NT / 32) * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    UpdCtx u;
    if (LAST) u = make_upd<W>(a, sx, l);
    double2 accr = make_double2(0.0, 0.0);
    for (int base = s0; base < s1; base += BPC) {
        const int br = base + warp * BRW + g;
        if (br < s1) {
            const int q0 = sptr[br - s0], q1 = sptr[br - s0 + 1];
            const int ls = (3 * (br - s0)) * W + 2 * l;
            double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
            for (int q = q0; q < q1; ++q) {
                const double2* bp = sblk + 5 * q;
                const double2 v01 = bp[0], v23 = bp[1], v45 = bp[2], v67 = bp[3], v8c = bp[4];
                const double* xr = xa + (3 * __double2loint(v8c.y)) * W + 2 * l;
                const double2 x0 = *reinterpret_cast<const double2*>(xr), x1 = *reinterpret_cast<const double2*>(xr + W),
                              x2 = *reinterpret_cast<const double2*>(xr + 2 * W);
                acc[0].x = fma(v01.x, x0.x, fma(v01.y, x1.x, fma(v23.x, x2.x, acc[0].x)));
                acc[0].y = fma(v01.x, x0.y, fma(v01.y, x1.y, fma(v23.x, x2.y, acc[0].y)));
                acc[1].x = fma(v23.y, x0.x, fma(v45.x, x1.x, fma(v45.y, x2.x, acc[1].x)));
                acc[1].y = fma(v23.y, x0.y, fma(v45.x, x1.y, fma(v45.y, x2.y, acc[1].y)));
                acc[2].x = fma(v67.x, x0.x, fma(v67.y, x1.x, fma(v8c.x, x2.x, acc[2].x)));
                acc[2].y = fma(v67.x, x0.y, fma(v67.y, x1.y, fma(v8c.x, x2.y, acc[2].y)));
            }
            double2 rv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 bv = *reinterpret_cast<const double2*>(bs + ls + i * W);
                rv[i].x = bv.x - acc[i].x; rv[i].y = bv.y - acc[i].y;
                if (LAST) { accr.x = fma(rv[i].x, rv[i].x, accr.x); accr.y = fma(rv[i].y, rv[i].y, accr.y); }
            }
            const double* Pb = sP + 9 * (br - s0);
            const size_t o = (size_t)(3 * br) * W + 2 * l;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 xo = *reinterpret_cast<const double2*>(xa + ls + i * W);
                const double2 xmv = *reinterpret_cast<const double2*>(xb + ls + i * W);
                double2 z, xn;
                z.x = fma(Pb[3 * i], rv[0].x, fma(Pb[3 * i + 1], rv[1].x, Pb[3 * i + 2] * rv[2].x));
                z.y = fma(Pb[3 * i], rv[0].y, fma(Pb[3 * i + 1], rv[1].y, Pb[3 * i + 2] * rv[2].y));
                xn.x = fma(omega, fma(a.alpha, z.x, xo.x) - xmv.x, xmv.x);
                xn.y = fma(omega, fma(a.alpha, z.y, xo.y) - xmv.y, xmv.y);
                *reinterpret_cast<double2*>(xb + ls + i * W) = xn;
                *reinterpret_cast<double2*>(xg_out + o + (size_t)i * W) = xn;
                if (LAST) row_update(u, o + (size_t)i * W, (size_t)(3 * br + i), xn);     // edge rows: W = x_K
            }
        }
    }
    return accr;
}

// ---- phase C: interior rows  w_I = T_I + R x_K  and their recurrence update
template <int W>
KEXP_PHASE_FN void narrow_phase_c(int slot, const StepCtx sx, const int* __restrict__ rptr, const int* __restrict__ rcol,
                                         int r0, int r1) {
    const NarrowArgs& a = c_nargs[slot];
    constexpr int LPR = W / 2;
    constexpr int RPC = NarrowCfg<W>::NT / LPR;
    const int lr = threadIdx.x % LPR;
    const UpdCtx u = make_upd<W>(a, sx, lr);
    const int qbase = __ldg(a.r_rowptr + r0);
    for (int r = r0 + threadIdx.x / LPR; r < r1; r += RPC) {
        const int q0 = rptr[r - r0], q1 = rptr[r - r0 + 1];
        const size_t row = (size_t)a.n_s + r;
        const size_t o = row * W + 2 * lr;
        double2 wv = ldcg2(a.T + o);
        for (int q = q0; q < q1; q += kRChunk) {
            double2 x[kRChunk];
            double v[kRChunk];
#pragma unroll
            for (int j = 0; j < kRChunk; ++j)
                if (q + j < q1) {
                    x[j] = ldcg2(a.X + (size_t)rcol[q + j] * W + 2 * lr);
                    v[j] = __ldg(a.r_val + qbase + q + j);
                }
#pragma unroll
            for (int j = 0; j < kRChunk; ++j)
                if (q + j < q1) { wv.x = fma(v[j], x[j].x, wv.x); wv.y = fma(v[j], x[j].y, wv.y); }
        }
        row_update(u, o, row, wv);
    }
}

// ---- convergence record of one solve (CTA 0, fixed summation order over the per-CTA block sums)
template <int W>
KEXP_PHASE_FN void narrow_check(int slot, double* fin, int K, double rtol2, int* s_bad) {
    const NarrowArgs& a = c_nargs[slot];
    constexpr int kNT = NarrowCfg<W>::NT;
    constexpr int TPV = kNT / (2 * W) > 32 ? 32 : kNT / (2 * W);      // threads per value
    const int val = threadIdx.x / TPV, part = threadIdx.x % TPV;
    if (val < 2 * W) {
        double s = 0.0;
        for (int b = part; b < (int)gridDim.x; b += TPV) s += __ldcg(a.partials + (size_t)b * 2 * W + val);
        fin[val * 32 + part] = s;
    }
    if (threadIdx.x == 0) *s_bad = 0;
    __syncthreads();
    double* tot = fin + 2 * W * 32;
    if (threadIdx.x < 2 * W) {
        double s = 0.0;
        for (int p = 0; p < TPV; ++p) s += fin[threadIdx.x * 32 + p];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    double* relv = tot + 2 * W;
    if (threadIdx.x < W) {
        const double rr = tot[threadIdx.x], bb = tot[W + threadIdx.x];
        double rel = 0.0;
        if (bb > 0.0) {
            rel = rr / bb;
            if (rr > rtol2 * bb) atomicOr(s_bad, 1);
        }
        relv[threadIdx.x] = rel;
        a.sc->rr[threadIdx.x] = rr; a.sc->bb[threadIdx.x] = bb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int c = 0; c < W; ++c) m = fmax(m, relv[c]);
        CgScalars* sc = a.sc;
        sc->max_rel_res = sqrt(m);
        sc->iters = K;
        sc->total_iters += (unsigned long long)K;
        sc->total_solves += 1ull;
        if (K > sc->max_iters) sc->max_iters = K;
        sc->done = *s_bad ? 0 : 1;
        if (*s_bad) { sc->unconverged += 1; sc->worst_rel_res = fmax(sc->worst_rel_res, sc->max_rel_res); }
    }
}

template <int W>
__global__ void __launch_bounds__(NarrowCfg<W>::NT, 1)
cheb_apply_narrow_kernel(const int slot) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int kNT = NarrowCfg<W>::NT, kNWarps = kNT / 32;
    constexpr int LPB = C::LPB;
    constexpr int LPR = W / 2;                         // lanes per scalar row
    constexpr int RPC = kNT / LPR;                     // scalar rows per CTA pass
    cg::grid_group grid = cg::this_grid();

    extern __shared__ double2 smem2[];
    double2* sblk = smem2;                                                     // S blocks of the own rows
    double* sP = reinterpret_cast<double*>(sblk + 5 * (size_t)a.cap_s_blocks); // block-Jacobi inverses of the own rows
    double* sred = sP + 9 * (size_t)a.cap_s_rows;                              // reduction scratch: 2 x warps x W, then 2 x W x 32
    int* sptr = reinterpret_cast<int*>(sred + 2 * kNWarpsMax * W + 2 * W * 32 + 4 * W);
    int* kptr = sptr + a.cap_s_rows + 1;
    int* kcol = kptr + a.cap_kt_rows + 1;
    int* rptr = kcol + a.cap_kt_blocks;
    int* rcol = rptr + a.cap_r_rows + 1;
    int* hcol = rcol + a.cap_r_nnz;
    // mass-solve state: two iterate buffers over [own | halo] slots and b on the own rows (16-byte aligned)
    double* xs0 = reinterpret_cast<double*>(smem2) + a.xs_offset;
    double* xs1 = xs0 + (size_t)3 * W * a.cap_slots;
    double* bsm = xs1 + (size_t)3 * W * a.cap_slots;
    __shared__ int s_bad;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB;
    int s0, s1, k0r, k1r, r0, r1;
    cta_range(a.s_nbr, s0, s1);
    cta_range(a.kt_nbr, k0r, k1r);
    cta_range(a.n_b, r0, r1);
    const int kt_b0 = __ldg(a.kt_browptr + k0r);

    // ---- stage the constant data of this CTA's row ranges
    {
        const int kb0 = __ldg(a.s_browptr + s0), kb1 = __ldg(a.s_browptr + s1);
        const double2* gsrc = reinterpret_cast<const double2*>(a.s_blocks) + (size_t)5 * kb0;
        for (int i = threadIdx.x; i < 5 * (kb1 - kb0); i += kNT) sblk[i] = __ldg(gsrc + i);
        for (int i = threadIdx.x; i <= s1 - s0; i += kNT) sptr[i] = __ldg(a.s_browptr + s0 + i) - kb0;
        for (int i = threadIdx.x; i < 9 * (s1 - s0); i += kNT) sP[i] = __ldg(a.Pinv + (size_t)9 * s0 + i);
        const int kb1k = __ldg(a.kt_browptr + k1r);
        for (int i = threadIdx.x; i <= k1r - k0r; i += kNT) kptr[i] = __ldg(a.kt_browptr + k0r + i) - kt_b0;
        for (int i = threadIdx.x; i < kb1k - kt_b0; i += kNT)
            kcol[i] = __double2loint(__ldg(a.kt_blocks + (size_t)10 * (kt_b0 + i) + 9));
        if (a.n_b > 0) {
            const int q0 = __ldg(a.r_rowptr + r0), q1 = __ldg(a.r_rowptr + r1);
            for (int i = threadIdx.x; i <= r1 - r0; i += kNT) rptr[i] = __ldg(a.r_rowptr + r0 + i) - q0;
            for (int i = threadIdx.x; i < q1 - q0; i += kNT) rcol[i] = __ldg(a.r_col + q0 + i);
        }
    }
    const int h0 = __ldg(a.halo_ptr + blockIdx.x), n_halo = __ldg(a.halo_ptr + blockIdx.x + 1) - h0;
    for (int i = threadIdx.x; i < n_halo; i += kNT) hcol[i] = __ldg(a.halo_cols + h0 + i);
    const int n_own = s1 - s0;
    // ---- prologue on a linear slice of the rows:  mode 0: Y = a_0 S_0 (and basis[:,0]);  mode 1: SA = z_M e_0, SB = 0
    {
        int p0, p1;
        cta_range(a.n, p0, p1);
        const int lp = threadIdx.x % LPR;
        const double2 a0 = make_double2(a.coef[2 * lp], a.coef[2 * lp + 1]);
        for (int r = p0 + threadIdx.x / LPR; r < p1; r += RPC) {
            const size_t o = (size_t)r * W + 2 * lp;
            if (a.mode == 0) {
                const double2 v = *reinterpret_cast<const double2*>(a.SA + o);
                *reinterpret_cast<double2*>(a.Y + o) = make_double2(a0.x * v.x, a0.y * v.y);
                if (a.basis && lp == 0) a.basis[(size_t)r * a.basis_ld] = v.x;
            } else {
                *reinterpret_cast<double2*>(a.SA + o) = make_double2(lp == 0 ? __ldg(a.basis + (size_t)r * a.basis_ld + a.M) : 0.0, 0.0);
                *reinterpret_cast<double2*>(a.SB + o) = make_double2(0.0, 0.0);
            }
        }
    }
    unsigned bar_target = 0;
    narrow_barrier(grid, a.barrier, bar_target);

    StepCtx sx;
    sx.Sin = a.SA;               // S_{n-1}
    sx.Snew = a.SB;              // holds S_{n-2}, becomes S_n
    for (int step = 1; step <= a.M; ++step) {
        // term index / coefficient row / scale of this step (Clenshaw runs m = M-1 ... 0)
        sx.step = step;
        sx.term = (a.mode == 0) ? step : a.M - step;
        sx.cs = (a.mode == 0) ? (step == 1 ? 1.0 : 2.0) : (sx.term == 0 ? 1.0 : 2.0);
        sx.has_old = (a.mode == 0) ? (step >= 2) : 1;
        sx.crow = a.coef + (size_t)((a.mode == 0) ? sx.term : 0) * kMaxW;

        narrow_phase_a<W>(slot, sx, kptr, kcol, k0r, k1r, kt_b0);
        narrow_barrier(grid, a.barrier, bar_target);

        // mass solve on the own rows of S (Chebyshev semi-iteration), K - 1 iterations behind x_1
        const int K = a.kterm[step];
        double* xk = (K % 2 == 1) ? a.X : a.P;               // x_1 is published here, x_K ends up in a.X
        double* xm = (xk == a.X) ? a.P : a.X;
        double* xa = xs0;                                    // shared: x_k on [own | halo]
        double* xb = xs1;                                    // shared: x_{k-1} on the own rows -> x_{k+1}
        double2 accb = narrow_solve_start<W>(slot, xk, xa, xb, bsm, sP, s0, s1);
        double2 accr = make_double2(0.0, 0.0);
        double omega = 1.0;
        for (int it = 1; it < K; ++it) {
            narrow_barrier(grid, a.barrier, bar_target);     // x_k of every CTA is published (and this CTA's shared copy complete)
            narrow_load_halo<W>(xk, xa, hcol, n_own, n_halo);
            __syncthreads();
            omega = (it == 1) ? 1.0 / (1.0 - 0.5 * a.mu * a.mu) : 1.0 / (1.0 - 0.25 * a.mu * a.mu * omega);
            if (it == K - 1) accr = narrow_solve_iter<W, true>(slot, sx, xa, xb, bsm, xm, omega, sblk, sP, sptr, s0, s1);
            else narrow_solve_iter<W, false>(slot, sx, xa, xb, bsm, xm, omega, sblk, sP, sptr, s0, s1);
            double* tsw = xk; xk = xm; xm = tsw;
            tsw = xa; xa = xb; xb = tsw;
        }
        // block sums of ||r||^2 and ||b||^2 per column (fixed order): shuffles over the groups of a warp, then the warps
        {
            double2 v[2] = {accr, accb};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int off = 16; off >= LPB; off >>= 1) {
                    v[q].x += __shfl_xor_sync(0xffffffffu, v[q].x, off);
                    v[q].y += __shfl_xor_sync(0xffffffffu, v[q].y, off);
                }
                if (lane < LPB) { sred[(q * kNWarps + warp) * W + 2 * l] = v[q].x; sred[(q * kNWarps + warp) * W + 2 * l + 1] = v[q].y; }
            }
            __syncthreads();
            if (threadIdx.x < 2 * W) {
                const int q = threadIdx.x / W, c = threadIdx.x % W;
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kNWarps; ++w) s += sred[(q * kNWarps + w) * W + c];
                a.partials[((size_t)blockIdx.x * 2 + q) * W + c] = s;
            }
        }
        narrow_barrier(grid, a.barrier, bar_target);                                           // x_K, the edge rows of S_n and the block sums are visible
        if (blockIdx.x == 0) narrow_check<W>(slot, sred + 2 * kNWarpsMax * W, K, a.rtol2[step], &s_bad);

        if (a.n_b > 0) {
            narrow_phase_c<W>(slot, sx, rptr, rcol, r0, r1);
            narrow_barrier(grid, a.barrier, bar_target);                                       // S_n complete
        }
        double* tsw = sx.Sin; sx.Sin = sx.Snew; sx.Snew = tsw;
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct NarrowCaps { int s_blocks, s_rows, kt_rows, kt_blocks, r_rows, r_nnz, slots, halo; size_t smem; };

static void range_of(int nrows, int c, int G, int& start, int& end) {
    const int per = nrows / G, rem = nrows % G;
    start = c * per + std::min(c, rem);
    end = start + per + (c < rem ? 1 : 0);
}

// bytes of the constant part (matrix, index arrays, scratch), rounded to 16: the iterate buffers start there
template <int W>
static size_t narrow_smem_fixed(const NarrowCaps& c) {
    size_t b = sizeof(double2) * 5 * (size_t)c.s_blocks + sizeof(double) * 9 * (size_t)c.s_rows;
    b += sizeof(double) * (2 * kNWarpsMax * W + 2 * W * 32 + 4 * W);
    b += sizeof(int) * ((size_t)c.s_rows + 1 + c.kt_rows + 1 + c.kt_blocks + c.r_rows + 1 + c.r_nnz + c.halo);
    return (b + 15) & ~(size_t)15;
}

template <int W>
static size_t narrow_smem(const NarrowCaps& c) {
    return narrow_smem_fixed<W>(c) + sizeof(double) * 3 * W * ((size_t)2 * c.slots + c.s_rows);
}

// Shared-memory capacities for a grid of G CTAs (from the host copies of the row pointers kept at upload time).
// Per-CTA local numbering of S for a grid of G CTAs: own rows take slots 0..n_own-1, the halo (block columns referenced
// outside the own range, sorted) the slots behind them; the device gets a copy of the blocks with the slot in place of the
// global block column.  Built once per operator.
static int build_narrow_plan(kexp_ctx* ctx, int G) {
    if (ctx->narrow_plan_G == G && ctx->d_nlblocks) return 0;
    const std::vector<int>& sb = ctx->As.h_browptr;
    const std::vector<double>& blk = ctx->As.h_blocks;
    if (sb.empty() || blk.empty()) return kexp::fail(ctx, "narrow apply: host copy of S not available");
    const int nbr = (int)sb.size() - 1;
    std::vector<double> lb(blk);
    std::vector<int> hptr((size_t)G + 1, 0), hcols, halo;
    int max_slots = 0, max_halo = 0;
    for (int c = 0; c < G; ++c) {
        int r0, r1;
        range_of(nbr, c, G, r0, r1);
        halo.clear();
        for (int k = sb[r0]; k < sb[r1]; ++k) {
            int64_t bits; memcpy(&bits, &blk[(size_t)k * 10 + 9], 8);
            const int col = (int)bits;
            if (col < r0 || col >= r1) halo.push_back(col);
        }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        for (int k = sb[r0]; k < sb[r1]; ++k) {
            int64_t bits; memcpy(&bits, &blk[(size_t)k * 10 + 9], 8);
            const int col = (int)bits;
            int64_t slot = (col >= r0 && col < r1) ? col - r0
                                                   : (r1 - r0) + (int)(std::lower_bound(halo.begin(), halo.end(), col) - halo.begin());
            memcpy(&lb[(size_t)k * 10 + 9], &slot, 8);
        }
        hcols.insert(hcols.end(), halo.begin(), halo.end());
        hptr[c + 1] = (int)hcols.size();
        max_slots = std::max(max_slots, r1 - r0 + (int)halo.size());
        max_halo = std::max(max_halo, (int)halo.size());
    }
    cudaFree(ctx->d_nlblocks); cudaFree(ctx->d_nhalo_ptr); cudaFree(ctx->d_nhalo_cols);
    ctx->d_nlblocks = nullptr; ctx->d_nhalo_ptr = ctx->d_nhalo_cols = nullptr;
    KEXP_CUDA(cudaMalloc(&ctx->d_nlblocks, sizeof(double) * lb.size()));
    KEXP_CUDA(cudaMalloc(&ctx->d_nhalo_ptr, sizeof(int) * hptr.size()));
    KEXP_CUDA(cudaMalloc(&ctx->d_nhalo_cols, sizeof(int) * std::max<size_t>(hcols.size(), 1)));
    KEXP_CUDA(cudaMemcpy(ctx->d_nlblocks, lb.data(), sizeof(double) * lb.size(), cudaMemcpyHostToDevice));
    KEXP_CUDA(cudaMemcpy(ctx->d_nhalo_ptr, hptr.data(), sizeof(int) * hptr.size(), cudaMemcpyHostToDevice));
    KEXP_CUDA(cudaMemcpy(ctx->d_nhalo_cols, hcols.data(), sizeof(int) * hcols.size(), cudaMemcpyHostToDevice));
    ctx->narrow_plan_G = G; ctx->narrow_max_slots = max_slots; ctx->narrow_max_halo = max_halo;
    return 0;
}

bool narrow_caps(kexp_ctx* ctx, int G, NarrowCaps& c) {
    const std::vector<int>& sb = ctx->As.h_browptr;
    const std::vector<int>& kb = ctx->Kt.h_browptr;
    const std::vector<int>& rp = ctx->Rt.h_rowptr;
    if (sb.empty() || kb.empty()) return false;
    if (build_narrow_plan(ctx, G)) return false;
    c = NarrowCaps{0, 0, 0, 0, 0, 0, ctx->narrow_max_slots, ctx->narrow_max_halo, 0};
    for (int i = 0; i < G; ++i) {
        int a, b;
        range_of((int)sb.size() - 1, i, G, a, b);
        c.s_rows = std::max(c.s_rows, b - a); c.s_blocks = std::max(c.s_blocks, sb[b] - sb[a]);
        range_of((int)kb.size() - 1, i, G, a, b);
        c.kt_rows = std::max(c.kt_rows, b - a); c.kt_blocks = std::max(c.kt_blocks, kb[b] - kb[a]);
        if (!rp.empty()) {
            range_of((int)rp.size() - 1, i, G, a, b);
            c.r_rows = std::max(c.r_rows, b - a); c.r_nnz = std::max(c.r_nnz, rp[b] - rp[a]);
        }
    }
    return true;
}

template <int W>
static int launch_narrow_w(kexp_ctx* ctx, NarrowArgs& a, const NarrowCaps& caps) {
    auto k = cheb_apply_narrow_kernel<W>;
    const size_t smem = narrow_smem<W>(caps);
    static size_t attr_smem = 0;           // per instantiation
    if (smem > attr_smem) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return kexp::fail(ctx, "narrow apply kernel: cannot opt in to the shared memory it needs");
        }
        attr_smem = smem;
    }
    int slot = ctx->narrow_slot;
    if (cudaMemcpyToSymbolAsync(c_nargs, &a, sizeof(NarrowArgs), sizeof(NarrowArgs) * (size_t)slot, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return kexp::fail(ctx, "narrow apply kernel: argument upload failed");
    void* args[] = {&slot};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(ctx->num_sms), dim3(NarrowCfg<W>::NT), args, smem, ctx->stream);
    if (e != cudaSuccess) return kexp::fail(ctx, std::string("narrow apply kernel launch: ") + cudaGetErrorString(e));
    return 0;
}

// Can a panel of width W run through the persistent kernel?  (Chebyshev mass solve, both matrices in 3x3 blocks with the
// diagonal block present, and the CTA's share of S, index arrays and scratch within 227 KB of shared memory.)
bool narrow_eligible(kexp_ctx* ctx, int W) {
    if (!ctx->use_narrow || W > ctx->narrow_max_w || W > 16) return false;
    if (ctx->solver != 1 || !ctx->Kt.bsr() || !ctx->As.bsr() || !ctx->As.diag_first) return false;
    if (ctx->n_b > 0 && (ctx->Rt.rowptr == nullptr)) return false;
    if (ctx->n % 3 || ctx->n_s % 3 || ctx->n_b % 3) return false;
    NarrowCaps caps;
    if (!narrow_caps(ctx, ctx->num_sms, caps)) return false;
    size_t smem = 0;
    switch (W) {
        case 2: smem = narrow_smem<2>(caps); break;
        case 4: smem = narrow_smem<4>(caps); break;
        case 8: smem = narrow_smem<8>(caps); break;
        case 16: smem = narrow_smem<16>(caps); break;
        default: return false;
    }
    return smem <= 227u * 1024u - 1024u;
}

// Enqueue one whole apply of the panel held in ctx->S0 (mode 0: result in ctx->Y; mode 1: Clenshaw, result in the buffer
// returned through *result).  kterm / rtol2: per-step iteration counts and squared tolerances (host, M+1 entries).
int launch_narrow_apply(kexp_ctx* ctx, int W, int M, int mode, const int* kterm, const double* rtol2, bool with_basis, const double** result) {
    NarrowCaps caps;
    if (!narrow_caps(ctx, ctx->num_sms, caps)) return kexp::fail(ctx, "narrow apply: operator structure not available");
    if (M + 1 > kMaxTerms) return kexp::fail(ctx, "series too long");
    if (!ctx->d_nk) {
        KEXP_CUDA(cudaMalloc(&ctx->d_nk, sizeof(int) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_nrtol2, sizeof(double) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_npartials, sizeof(double) * 2 * kMaxW * (size_t)ctx->num_sms + 64));
    }
    // pageable sources: the driver stages them before the call returns, so the host vectors may be reused at once
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nk, kterm, sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nrtol2, rtol2, sizeof(double) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    NarrowArgs a;
    a.kt_browptr = ctx->Kt.browptr; a.kt_blocks = ctx->Kt.blocks; a.kt_rowmap = ctx->Kt.rowmap; a.kt_nbr = ctx->Kt.nbr;
    a.s_browptr = ctx->As.browptr; a.s_blocks = ctx->d_nlblocks; a.Pinv = ctx->d_Pinv; a.s_nbr = ctx->As.nbr;
    a.halo_ptr = ctx->d_nhalo_ptr; a.halo_cols = ctx->d_nhalo_cols;
    a.r_rowptr = ctx->Rt.rowptr; a.r_col = ctx->Rt.col; a.r_val = ctx->Rt.val; a.Dblk = ctx->d_Dblk;
    a.n = (int)ctx->n; a.n_s = (int)ctx->n_s; a.n_b = (int)ctx->n_b; a.n_d = (int)ctx->n_d;
    a.SA = ctx->S0; a.SB = ctx->S1; a.Y = ctx->Y; a.T = ctx->T; a.X = ctx->X; a.P = ctx->P;
    a.tscale = ctx->d_tscale; a.coef = ctx->d_tscale + kMaxW; a.kterm = ctx->d_nk; a.rtol2 = ctx->d_nrtol2;
    a.M = M; a.mode = mode;
    a.basis = (mode == 1 || with_basis) ? ctx->d_basis : nullptr; a.basis_ld = ctx->basis_ld;
    a.alpha = 2.0 / (ctx->cheb_hi + ctx->cheb_lo);
    a.mu = (ctx->cheb_hi - ctx->cheb_lo) / (ctx->cheb_hi + ctx->cheb_lo);
    a.sc = ctx->d_sc; a.partials = ctx->d_npartials;
    a.barrier = reinterpret_cast<unsigned*>(ctx->d_npartials + 2 * kMaxW * (size_t)ctx->num_sms);
    KEXP_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), ctx->stream));
    a.cap_s_blocks = caps.s_blocks; a.cap_s_rows = caps.s_rows; a.cap_kt_rows = caps.kt_rows; a.cap_kt_blocks = caps.kt_blocks;
    a.cap_r_rows = caps.r_rows; a.cap_r_nnz = caps.r_nnz; a.cap_slots = caps.slots; a.cap_halo = caps.halo;
    switch (W) {
        case 2: a.xs_offset = (int)(narrow_smem_fixed<2>(caps) / sizeof(double)); break;
        case 4: a.xs_offset = (int)(narrow_smem_fixed<4>(caps) / sizeof(double)); break;
        case 8: a.xs_offset = (int)(narrow_smem_fixed<8>(caps) / sizeof(double)); break;
        default: a.xs_offset = (int)(narrow_smem_fixed<16>(caps) / sizeof(double)); break;
    }
    int rc = 1;
    switch (W) {
        case 2: rc = launch_narrow_w<2>(ctx, a, caps); break;
        case 4: rc = launch_narrow_w<4>(ctx, a, caps); break;
        case 8: rc = launch_narrow_w<8>(ctx, a, caps); break;
        case 16: rc = launch_narrow_w<16>(ctx, a, caps); break;
        default: return kexp::fail(ctx, "narrow apply: unsupported panel width");
    }
    if (rc) return rc;
    if (result) *result = (M % 2 == 0) ? ctx->S0 : ctx->S1;       // the buffer written by the last step
    ctx->solve_path = 3;
    return 0;
}

}  // namespace kexp
