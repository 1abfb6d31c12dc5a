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
//   * inside the mass solve the iterate lives in SHARED memory: x_k on the CTA's own rows plus the halo rows its
//     blocks reference (local slot numbers precomputed on the host), x_{k-1} / x_{k+1}, and b.  Per exchange a CTA
//     reads only its halo (each row once, coalesced, ld.global.cg) and publishes its own rows; the products read
//     shared memory, so there is no gather latency chain, no L1 coherence question, hardly any register pressure;
//   * TWO Chebyshev iterations per grid barrier where shared memory allows (single-column panels at ref 5): the CTA
//     also holds the matrix rows of its first halo ring H1 and the iterate on a second ring H2, computes x_{k+1} on
//     own + H1 redundantly (bit-identical to what the owners compute) and then x_{k+2} on its own rows -- half the
//     barriers and exchanges of the solve for ~45 % more local work in a latency-bound kernel;
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
#include <type_traits>
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
    // S as per-CTA LOCAL problems (build_narrow_plan): the blocks of the rows a CTA computes -- its own rows, then (two-level
    // plan) the rows of its first halo ring H1 -- packed per CTA with the shared-memory SLOT of each block's column in the
    // 10th double (slots: own rows | H1 | H2); lrowptr: local row pointers; halo_cols: global block rows of H1 then H2
    const double* lblocks; const int* lrowptr; const int* cta_blk_ptr; const int* cta_row_ptr;
    const int* halo_ptr; const int* halo_cols; const int* halo1_cnt;
    const double* Pinv; int s_nbr; int levels;             // levels: 1 (one iteration per barrier) or 2
    const int* kt_cols;                                      // block columns of K~ (global copy of the 10th doubles)
    // R (CSR, n_b x n_s) or null; Dblk (n_d/3 blocks) or null
    const int* r_rowptr; const int* r_col; const double* r_val; const double* Dblk;
    int n, n_s, n_b, n_d;
    // panels [rows][W]
    double *SA, *SB, *Y, *T, *X;
    double* xg[4];                       // rotating publication buffers of the mass-solve iterates (n_s x W each)
    // per-panel tables (device): tscale[W], coef[(M+1)][kMaxW]; per-term iteration counts and squared tolerances
    const double* tscale; const double* coef; const int* kterm; const double* rtol2;
    int M;                               // highest term index
    int mode;                            // 0: forward recurrence y = sum a_n S_n;  1: Clenshaw (fused weighted average)
    double* basis; int basis_ld;         // mode 0: store column 0 of S_n (or null);  mode 1: z_m (required)
    double alpha, mu;                    // Chebyshev acceleration of the mass solve
    CgScalars* sc; double* partials;     // [gridDim][2][W] block sums of ||r||^2, ||b||^2
    unsigned* barrier;                   // grid barrier counter (zeroed before the launch)
    // shared-memory layout (element counts, maxima over the CTAs)
    int cap_bpr, cap_s_rows, cap_kt_rows, cap_r_rows, cap_slots, cap_halo;     // cap_bpr: most blocks in a computed row
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
KEXP_PHASE_FN void narrow_phase_a(int slot, const StepCtx sx, const int* __restrict__ kptr, int k0r, int k1r, int kt_b0) {
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
                        const double* xr = Sin + (size_t)(3 * __ldg(a.kt_cols + kt_b0 + q + j)) * W + 2 * l;
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
// Shared-memory state of the mass solve (per CTA), in block-row SLOTS: own rows | first halo ring H1 | second ring H2.
// Row r of a slot starts at (3*slot + r) * W doubles.  Two iterate arrays over all slots and b on the rows the CTA computes
// (own rows, plus H1 in the two-level plan).
//
// ---- phase B, start: on the computed rows  b <- T (global), x_1 = alpha P^-1 b -> cur, x_0 = 0 -> prev; the own rows of
// x_1 and x_0 are published (g1, g0).  x_1 on H1 is computed locally (bit-identical to the owner's).  Returns ||b||^2 (own).
template <int W>
KEXP_PHASE_FN double2 narrow_solve_start(int slot, double* g1, double* g0, double* cur, double* prev, double* bs, const double* __restrict__ sP,
                                         const int* __restrict__ hcol, int s0, int n_own, int n_rows) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int BPC = (NarrowCfg<W>::NT / 32) * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    double2 accb = make_double2(0.0, 0.0);
    for (int base = 0; base < n_rows; base += BPC) {
        const int lr = base + warp * BRW + g;
        if (lr < n_rows) {
            const bool own = lr < n_own;
            const int gr = own ? s0 + lr : hcol[lr - n_own];
            const size_t o = (size_t)(3 * gr) * W + 2 * l;
            const int ls = (3 * lr) * W + 2 * l;
            double2 bv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                bv[i] = ldcg2(a.T + o + (size_t)i * W);
                *reinterpret_cast<double2*>(bs + ls + i * W) = bv[i];
                if (own) { accb.x = fma(bv[i].x, bv[i].x, accb.x); accb.y = fma(bv[i].y, bv[i].y, accb.y); }
            }
            const double* Pb = sP + 9 * lr;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double2 z;
                z.x = a.alpha * fma(Pb[3 * i], bv[0].x, fma(Pb[3 * i + 1], bv[1].x, Pb[3 * i + 2] * bv[2].x));
                z.y = a.alpha * fma(Pb[3 * i], bv[0].y, fma(Pb[3 * i + 1], bv[1].y, Pb[3 * i + 2] * bv[2].y));
                *reinterpret_cast<double2*>(cur + ls + i * W) = z;
                *reinterpret_cast<double2*>(prev + ls + i * W) = make_double2(0.0, 0.0);
                if (own) {
                    *reinterpret_cast<double2*>(g1 + o + (size_t)i * W) = z;
                    *reinterpret_cast<double2*>(g0 + o + (size_t)i * W) = make_double2(0.0, 0.0);
                }
            }
        }
    }
    return accb;
}

// ---- phase B, halo exchange: halo slots [h_lo, h_hi) of a published iterate, global -> shared (each row once, 16-byte
// pieces, coalesced within a row)
template <int W>
KEXP_PHASE_FN void narrow_load_halo(const double* xg, double* xs, const int* __restrict__ hcol, int n_own, int h_lo, int h_hi) {
    constexpr int PPR = 3 * W / 2;                     // 16-byte pieces per block row (3 rows x W doubles, contiguous)
    constexpr int NT = NarrowCfg<W>::NT;
    for (int e = h_lo * PPR + threadIdx.x; e < h_hi * PPR; e += NT) {
        const int h = e / PPR, p = e % PPR;
        const double2 v = ldcg2(xg + (size_t)(3 * hcol[h]) * W + 2 * p);
        *reinterpret_cast<double2*>(xs + (size_t)(3 * (n_own + h)) * W + 2 * p) = v;
    }
}

// ---- phase B, one Chebyshev iteration on the local rows [0, n_rows), everything read from shared memory:
//   x_{k+1} = x_{k-1} + omega (x_k + alpha P^-1 (b - S x_k) - x_{k-1}),   x_k = cur (all slots the rows reference), x_{k-1} = prev,
// written over prev; the own rows (lr < n_own) are published to gout (and to gout2: the final iterate also goes to a.X).
// LAST also forms ||b - S x_k||^2 and the recurrence update of the edge rows (W = x_K).
template <int W, bool LAST>
KEXP_PHASE_FN double2 narrow_solve_iter(int slot, const StepCtx sx, const double* cur, double* prev, const double* bs, double* gout, double* gout2,
                                        double omega, const double* __restrict__ sv, const int* __restrict__ ss, const double* __restrict__ sP,
                                        const int* __restrict__ scnt, int ell_ld, int s0, int n_own, int n_rows) {
    const NarrowArgs& a = c_nargs[slot];
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int BPC = (NarrowCfg<W>::NT / 32) * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    UpdCtx u;
    if (LAST) u = make_upd<W>(a, sx, l);
    double2 accr = make_double2(0.0, 0.0);
    for (int base = 0; base < n_rows; base += BPC) {
        const int lr = base + warp * BRW + g;
        if (lr < n_rows) {
            const int nq = scnt[lr];
            const int ls = (3 * lr) * W + 2 * l;
            double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
            // the matrix is stored TRANSPOSED in shared memory (element e of block q of local row lr at sv[(9 q + e) ell_ld + lr],
            // its column slot at ss[q ell_ld + lr]): the rows of a warp read consecutive words -- no bank conflicts, where the
            // 80-byte blocks of one row per lane conflicted 4-way
            for (int q = 0; q < nq; ++q) {
                const double* vp = sv + (size_t)(9 * q) * ell_ld + lr;
                const double* xr = cur + (3 * ss[q * ell_ld + lr]) * W + 2 * l;
                const double2 x0 = *reinterpret_cast<const double2*>(xr), x1 = *reinterpret_cast<const double2*>(xr + W),
                              x2 = *reinterpret_cast<const double2*>(xr + 2 * W);
                const double m00 = vp[0], m01 = vp[ell_ld], m02 = vp[2 * ell_ld], m10 = vp[3 * ell_ld], m11 = vp[4 * ell_ld], m12 = vp[5 * ell_ld],
                             m20 = vp[6 * ell_ld], m21 = vp[7 * ell_ld], m22 = vp[8 * ell_ld];
                // same order as the wide-panel kernels: for each block, columns r = 0, 1, 2 of the 3x3 block in turn
                acc[0].x = fma(m00, x0.x, acc[0].x); acc[0].y = fma(m00, x0.y, acc[0].y);
                acc[1].x = fma(m10, x0.x, acc[1].x); acc[1].y = fma(m10, x0.y, acc[1].y);
                acc[2].x = fma(m20, x0.x, acc[2].x); acc[2].y = fma(m20, x0.y, acc[2].y);
                acc[0].x = fma(m01, x1.x, acc[0].x); acc[0].y = fma(m01, x1.y, acc[0].y);
                acc[1].x = fma(m11, x1.x, acc[1].x); acc[1].y = fma(m11, x1.y, acc[1].y);
                acc[2].x = fma(m21, x1.x, acc[2].x); acc[2].y = fma(m21, x1.y, acc[2].y);
                acc[0].x = fma(m02, x2.x, acc[0].x); acc[0].y = fma(m02, x2.y, acc[0].y);
                acc[1].x = fma(m12, x2.x, acc[1].x); acc[1].y = fma(m12, x2.y, acc[1].y);
                acc[2].x = fma(m22, x2.x, acc[2].x); acc[2].y = fma(m22, x2.y, acc[2].y);
            }
            double2 rv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 bv = *reinterpret_cast<const double2*>(bs + ls + i * W);
                rv[i].x = bv.x - acc[i].x; rv[i].y = bv.y - acc[i].y;
                if (LAST) { accr.x = fma(rv[i].x, rv[i].x, accr.x); accr.y = fma(rv[i].y, rv[i].y, accr.y); }
            }
            const double* Pb = sP + 9 * lr;
            const bool own = lr < n_own;
            const size_t o = (size_t)(3 * (s0 + lr)) * W + 2 * l;          // only meaningful for own rows
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double2 xo = *reinterpret_cast<const double2*>(cur + ls + i * W);
                const double2 xmv = *reinterpret_cast<const double2*>(prev + ls + i * W);
                double2 z, xn;
                z.x = fma(Pb[3 * i], rv[0].x, fma(Pb[3 * i + 1], rv[1].x, Pb[3 * i + 2] * rv[2].x));
                z.y = fma(Pb[3 * i], rv[0].y, fma(Pb[3 * i + 1], rv[1].y, Pb[3 * i + 2] * rv[2].y));
                xn.x = fma(omega, fma(a.alpha, z.x, xo.x) - xmv.x, xmv.x);
                xn.y = fma(omega, fma(a.alpha, z.y, xo.y) - xmv.y, xmv.y);
                *reinterpret_cast<double2*>(prev + ls + i * W) = xn;
                if (own) {
                    if (gout) *reinterpret_cast<double2*>(gout + o + (size_t)i * W) = xn;
                    if (gout2) *reinterpret_cast<double2*>(gout2 + o + (size_t)i * W) = xn;
                    if (LAST) row_update(u, o + (size_t)i * W, (size_t)(3 * (s0 + lr) + i), xn);     // edge rows: W = x_K
                }
            }
        }
    }
    return accr;
}

// ---- phase C: interior rows  w_I = T_I + R x_K  and their recurrence update
template <int W>
KEXP_PHASE_FN void narrow_phase_c(int slot, const StepCtx sx, const int* __restrict__ rptr, int r0, int r1) {
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
                    x[j] = ldcg2(a.X + (size_t)__ldg(a.r_col + qbase + q + j) * W + 2 * lr);
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
    const int ell_ld = a.cap_s_rows;                                           // leading dimension of the transposed matrix arrays
    double* sv = reinterpret_cast<double*>(smem2);                             // S values, transposed: [9 * bpr][ell_ld]
    double* sP = sv + (size_t)9 * a.cap_bpr * ell_ld;                          // block-Jacobi inverses of the computed rows
    double* sred = sP + 9 * (size_t)a.cap_s_rows;                              // reduction scratch: 2 x warps x W, then 2 x W x 32
    int* ss = reinterpret_cast<int*>(sred + 2 * kNWarpsMax * W + 2 * W * 32 + 4 * W);   // column slots, transposed: [bpr][ell_ld]
    int* scnt = ss + (size_t)a.cap_bpr * ell_ld;                               // blocks per computed row
    int* kptr = scnt + a.cap_s_rows;
    int* rptr = kptr + a.cap_kt_rows + 1;
    int* hcol = rptr + a.cap_r_rows + 1;
    // mass-solve state: two iterate arrays over [own | H1 | H2] slots and b on the computed rows (16-byte aligned)
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
    const int n_own = s1 - s0;
    const int h0 = __ldg(a.halo_ptr + blockIdx.x), n_halo = __ldg(a.halo_ptr + blockIdx.x + 1) - h0;
    const int n_h1 = __ldg(a.halo1_cnt + blockIdx.x);
    const int n_rows = (a.levels == 2) ? n_own + n_h1 : n_own;                 // rows this CTA computes

    // ---- stage the constant data of this CTA
    {
        for (int i = threadIdx.x; i < n_halo; i += kNT) hcol[i] = __ldg(a.halo_cols + h0 + i);
        const int b0 = __ldg(a.cta_blk_ptr + blockIdx.x);
        const double* gsrc = a.lblocks + (size_t)10 * b0;
        const int* lrp = a.lrowptr + __ldg(a.cta_row_ptr + blockIdx.x);
        for (int lr = threadIdx.x; lr < n_rows; lr += kNT) {                   // one row per thread: transpose its blocks
            const int q0 = __ldg(lrp + lr), q1 = __ldg(lrp + lr + 1);
            scnt[lr] = q1 - q0;
            for (int q = q0; q < q1; ++q) {
                const double* bp = gsrc + (size_t)10 * q;
#pragma unroll
                for (int e = 0; e < 9; ++e) sv[(size_t)(9 * (q - q0) + e) * ell_ld + lr] = __ldg(bp + e);
                ss[(q - q0) * ell_ld + lr] = __double2loint(__ldg(bp + 9));
            }
        }
        __syncthreads();                                                       // hcol
        for (int i = threadIdx.x; i < 9 * n_rows; i += kNT) {
            const int lr = i / 9, e = i % 9;
            const int gr = lr < n_own ? s0 + lr : hcol[lr - n_own];
            sP[i] = __ldg(a.Pinv + (size_t)9 * gr + e);
        }
        for (int i = threadIdx.x; i <= k1r - k0r; i += kNT) kptr[i] = __ldg(a.kt_browptr + k0r + i) - kt_b0;
        if (a.n_b > 0) {
            const int q0 = __ldg(a.r_rowptr + r0);
            for (int i = threadIdx.x; i <= r1 - r0; i += kNT) rptr[i] = __ldg(a.r_rowptr + r0 + i) - q0;
        }
    }
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

        narrow_phase_a<W>(slot, sx, kptr, k0r, k1r, kt_b0);
        narrow_barrier(grid, a.barrier, bar_target);

        // ---- mass solve on the own rows of S: Chebyshev semi-iteration, K - 1 iterations behind x_1, one or two per barrier.
        // Published iterates rotate through four global buffers: gc holds x_k, gp holds x_{k-1}, the other two are written.
        const int K = a.kterm[step];
        int gc = 0, gp = 1;
        double* cur = xs0;                                   // shared: x_k on all slots the computed rows reference
        double* prev = xs1;                                  // shared: x_{k-1} on the computed rows -> x_{k+1}
        double2 accb = narrow_solve_start<W>(slot, a.xg[gc], a.xg[gp], cur, prev, bsm, sP, hcol, s0, n_own, n_rows);
        double2 accr = make_double2(0.0, 0.0);
        double omega = 1.0;
        int it = 1;                                          // next iteration to run: x_it -> x_{it+1}, it = 1 .. K-1
        const int iters = K - 1;
        auto next_omega = [&](int j) { omega = (j == 1) ? 1.0 / (1.0 - 0.5 * a.mu * a.mu) : 1.0 / (1.0 - 0.25 * a.mu * a.mu * omega); };
        auto free_buf = [&](int skip) { for (int b = 0; b < 4; ++b) if (b != gc && b != gp && b != skip) return b; return 0; };
        if (a.levels != 2 || (iters & 1)) {
            // single iterations: all of them (one-level plan) or one to make the rest an even count
            const int singles = (a.levels != 2) ? iters : 1;
            for (int j = 0; j < singles; ++j, ++it) {
                narrow_barrier(grid, a.barrier, bar_target);     // x_k of every CTA is published
                narrow_load_halo<W>(a.xg[gc], cur, hcol, n_own, 0, n_h1);
                __syncthreads();
                next_omega(it);
                const int f = free_buf(-1);
                if (it == K - 1) accr = narrow_solve_iter<W, true>(slot, sx, cur, prev, bsm, nullptr, a.X, omega, sv, ss, sP, scnt, ell_ld, s0, n_own, n_own);
                else narrow_solve_iter<W, false>(slot, sx, cur, prev, bsm, a.xg[f], nullptr, omega, sv, ss, sP, scnt, ell_ld, s0, n_own, n_own);
                gp = gc; gc = f;
                double* tsw = cur; cur = prev; prev = tsw;
            }
        }
        for (; it < K; it += 2) {
            // two iterations per barrier: x_{k+1} on own + H1 (needs x_k on own + H1 + H2, x_{k-1} on own + H1), then x_{k+2} on own
            narrow_barrier(grid, a.barrier, bar_target);
            narrow_load_halo<W>(a.xg[gc], cur, hcol, n_own, 0, n_halo);
            narrow_load_halo<W>(a.xg[gp], prev, hcol, n_own, 0, n_h1);
            __syncthreads();
            const int f0 = free_buf(-1), f1 = free_buf(f0);
            next_omega(it);
            narrow_solve_iter<W, false>(slot, sx, cur, prev, bsm, a.xg[f0], nullptr, omega, sv, ss, sP, scnt, ell_ld, s0, n_own, n_rows);
            __syncthreads();                                     // x_{k+1} on own + H1 complete in `prev`
            next_omega(it + 1);
            if (it + 1 == K - 1) accr = narrow_solve_iter<W, true>(slot, sx, prev, cur, bsm, nullptr, a.X, omega, sv, ss, sP, scnt, ell_ld, s0, n_own, n_own);
            else narrow_solve_iter<W, false>(slot, sx, prev, cur, bsm, a.xg[f1], nullptr, omega, sv, ss, sP, scnt, ell_ld, s0, n_own, n_own);
            gp = f0; gc = f1;                                    // cur holds x_{k+2} (own), prev x_{k+1} (own + H1): roles unchanged
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
        narrow_barrier(grid, a.barrier, bar_target);           // x_K (in a.X), the edge rows of S_n and the block sums are visible
        if (blockIdx.x == 0) narrow_check<W>(slot, sred + 2 * kNWarpsMax * W, K, a.rtol2[step], &s_bad);

        if (a.n_b > 0) {
            narrow_phase_c<W>(slot, sx, rptr, r0, r1);
            narrow_barrier(grid, a.barrier, bar_target);       // S_n complete
        }
        double* tsw = sx.Sin; sx.Sin = sx.Snew; sx.Snew = tsw;
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct NarrowCaps { int bpr, s_rows, kt_rows, r_rows, slots, halo; };

static void range_of(int nrows, int c, int G, int& start, int& end) {
    const int per = nrows / G, rem = nrows % G;
    start = c * per + std::min(c, rem);
    end = start + per + (c < rem ? 1 : 0);
}

// bytes of the constant part (matrix, index arrays, scratch), rounded to 16: the iterate buffers start there
template <int W>
static size_t narrow_smem_fixed(const NarrowCaps& c) {
    size_t b = sizeof(double) * 9 * (size_t)c.bpr * c.s_rows + sizeof(double) * 9 * (size_t)c.s_rows;
    b += sizeof(double) * (2 * kNWarpsMax * W + 2 * W * 32 + 4 * W);
    b += sizeof(int) * ((size_t)c.bpr * c.s_rows + c.s_rows + c.kt_rows + 1 + c.r_rows + 1 + c.halo);
    return (b + 15) & ~(size_t)15;
}

template <int W>
static size_t narrow_smem(const NarrowCaps& c) {
    return narrow_smem_fixed<W>(c) + sizeof(double) * 3 * W * ((size_t)2 * c.slots + c.s_rows);
}

__global__ void extract_block_cols_kernel(const double* __restrict__ blocks, int64_t nblocks, int* __restrict__ cols) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += (int64_t)gridDim.x * blockDim.x)
        cols[i] = __double2loint(blocks[10 * i + 9]);
}

// Per-CTA local problems of S for a grid of G CTAs (see NarrowArgs).  levels = 1: the CTA computes its own rows, slots =
// own | H1 (block columns referenced outside the own range, sorted).  levels = 2: it also computes the rows of H1
// (redundantly), slots = own | H1 | H2 with H2 the columns referenced by H1 rows outside own + H1.  Built once per
// operator and level count.
static int build_narrow_plan(kexp_ctx* ctx, int G, int levels) {
    NarrowPlanDev& P = ctx->narrow_plan[levels - 1];
    if (P.G == G && P.lblocks) return 0;
    const std::vector<int>& sb = ctx->As.h_browptr;
    const std::vector<double>& blk = ctx->As.h_blocks;
    if (sb.empty() || blk.empty()) return kexp::fail(ctx, "narrow apply: host copy of S not available");
    const int nbr = (int)sb.size() - 1;
    auto colof = [&](int k) { int64_t bits; memcpy(&bits, &blk[(size_t)k * 10 + 9], 8); return (int)bits; };
    std::vector<double> lb;
    std::vector<int> lrow, cta_blk((size_t)G + 1, 0), cta_row((size_t)G + 1, 0), hptr((size_t)G + 1, 0), hcols, h1cnt((size_t)G, 0);
    std::vector<int> h1, h2, rows;
    NarrowCaps caps{0, 0, 0, 0, 0, 0};
    for (int c = 0; c < G; ++c) {
        int r0, r1;
        range_of(nbr, c, G, r0, r1);
        const int n_own = r1 - r0;
        h1.clear(); h2.clear();
        for (int k = sb[r0]; k < sb[r1]; ++k) { const int col = colof(k); if (col < r0 || col >= r1) h1.push_back(col); }
        std::sort(h1.begin(), h1.end());
        h1.erase(std::unique(h1.begin(), h1.end()), h1.end());
        if (levels == 2) {
            for (int r : h1)
                for (int k = sb[r]; k < sb[r + 1]; ++k) {
                    const int col = colof(k);
                    if ((col < r0 || col >= r1) && !std::binary_search(h1.begin(), h1.end(), col)) h2.push_back(col);
                }
            std::sort(h2.begin(), h2.end());
            h2.erase(std::unique(h2.begin(), h2.end()), h2.end());
        }
        auto slot_of = [&](int col) -> int64_t {
            if (col >= r0 && col < r1) return col - r0;
            auto i1 = std::lower_bound(h1.begin(), h1.end(), col);
            if (i1 != h1.end() && *i1 == col) return n_own + (i1 - h1.begin());
            auto i2 = std::lower_bound(h2.begin(), h2.end(), col);
            return n_own + (int64_t)h1.size() + (i2 - h2.begin());
        };
        rows.clear();
        for (int r = r0; r < r1; ++r) rows.push_back(r);
        if (levels == 2) rows.insert(rows.end(), h1.begin(), h1.end());
        const size_t blk0 = lb.size() / 10;
        lrow.push_back(0);
        for (int r : rows) {
            for (int k = sb[r]; k < sb[r + 1]; ++k) {
                const size_t o = lb.size();
                lb.insert(lb.end(), blk.begin() + (size_t)k * 10, blk.begin() + (size_t)k * 10 + 10);
                const int64_t slot = slot_of(colof(k));
                memcpy(&lb[o + 9], &slot, 8);
            }
            lrow.push_back((int)(lb.size() / 10 - blk0));
        }
        cta_blk[c + 1] = (int)(lb.size() / 10);
        cta_row[c + 1] = (int)lrow.size();
        hcols.insert(hcols.end(), h1.begin(), h1.end());
        hcols.insert(hcols.end(), h2.begin(), h2.end());
        hptr[c + 1] = (int)hcols.size();
        h1cnt[c] = (int)h1.size();
        for (int r : rows) caps.bpr = std::max(caps.bpr, sb[r + 1] - sb[r]);
        caps.s_rows = std::max(caps.s_rows, (int)rows.size());
        caps.slots = std::max(caps.slots, n_own + (int)h1.size() + (int)h2.size());
        caps.halo = std::max(caps.halo, (int)(h1.size() + h2.size()));
    }
    cudaFree(P.lblocks); cudaFree(P.lrowptr); cudaFree(P.cta_blk_ptr); cudaFree(P.cta_row_ptr);
    cudaFree(P.halo_ptr); cudaFree(P.halo_cols); cudaFree(P.halo1_cnt);
    P = NarrowPlanDev();
    auto up = [&](auto** dst, const auto& v) -> bool {
        using T = std::remove_reference_t<decltype(**dst)>;
        if (cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1)) != cudaSuccess) return true;
        return cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess;
    };
    if (up(&P.lblocks, lb) || up(&P.lrowptr, lrow) || up(&P.cta_blk_ptr, cta_blk) || up(&P.cta_row_ptr, cta_row) ||
        up(&P.halo_ptr, hptr) || up(&P.halo_cols, hcols) || up(&P.halo1_cnt, h1cnt))
        return kexp::fail(ctx, "out of device memory for the narrow-apply plan");
    P.G = G; P.cap_blocks = caps.bpr; P.cap_rows = caps.s_rows; P.cap_slots = caps.slots; P.cap_halo = caps.halo;
    return 0;
}

static bool narrow_caps(kexp_ctx* ctx, int G, int levels, NarrowCaps& c) {
    const std::vector<int>& kb = ctx->Kt.h_browptr;
    const std::vector<int>& rp = ctx->Rt.h_rowptr;
    if (ctx->As.h_browptr.empty() || kb.empty()) return false;
    if (build_narrow_plan(ctx, G, levels)) return false;
    const NarrowPlanDev& P = ctx->narrow_plan[levels - 1];
    c = NarrowCaps{P.cap_blocks, P.cap_rows, 0, 0, P.cap_slots, P.cap_halo};
    for (int i = 0; i < G; ++i) {
        int a, b;
        range_of((int)kb.size() - 1, i, G, a, b);
        c.kt_rows = std::max(c.kt_rows, b - a);
        if (!rp.empty()) {
            range_of((int)rp.size() - 1, i, G, a, b);
            c.r_rows = std::max(c.r_rows, b - a);
        }
    }
    return true;
}

template <int W>
static int launch_narrow_w(kexp_ctx* ctx, NarrowArgs& a, const NarrowCaps& caps) {
    auto k = cheb_apply_narrow_kernel<W>;
    const size_t smem = narrow_smem<W>(caps);
    a.xs_offset = (int)(narrow_smem_fixed<W>(caps) / sizeof(double));
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

static size_t narrow_smem_for(int W, const NarrowCaps& caps) {
    switch (W) {
        case 2: return narrow_smem<2>(caps);
        case 4: return narrow_smem<4>(caps);
        case 8: return narrow_smem<8>(caps);
        case 16: return narrow_smem<16>(caps);
        default: return (size_t)1 << 40;
    }
}

// How many Chebyshev iterations per barrier a panel of width W can run (0: the persistent kernel is not applicable):
// Chebyshev mass solve, both matrices in 3x3 blocks, and the CTA's local problem + iterate buffers within the shared memory
// an SM offers.  Two iterations per barrier when the two-level problem fits (single-column panels at ref 5), else one.
static int narrow_levels(kexp_ctx* ctx, int W) {
    if (!ctx->use_narrow || W > ctx->narrow_max_w || W > 16) return 0;
    if (ctx->solver != 1 || !ctx->Kt.bsr() || !ctx->As.bsr()) return 0;
    if (ctx->n_b > 0 && (ctx->Rt.rowptr == nullptr)) return 0;
    if (ctx->n % 3 || ctx->n_s % 3 || ctx->n_b % 3) return 0;
    const size_t limit = 227u * 1024u - 1024u;
    for (int levels = std::min(2, ctx->narrow_sstep); levels >= 1; --levels) {
        NarrowCaps caps;
        if (!narrow_caps(ctx, ctx->num_sms, levels, caps)) return 0;
        if (narrow_smem_for(W, caps) <= limit) return levels;
    }
    return 0;
}

bool narrow_eligible(kexp_ctx* ctx, int W) { return narrow_levels(ctx, W) > 0; }

// Enqueue one whole apply of the panel held in ctx->S0 (mode 0: result in ctx->Y; mode 1: Clenshaw, result in the buffer
// returned through *result).  kterm / rtol2: per-step iteration counts and squared tolerances (host, M+1 entries).
int launch_narrow_apply(kexp_ctx* ctx, int W, int M, int mode, const int* kterm, const double* rtol2, bool with_basis, const double** result) {
    const int levels = narrow_levels(ctx, W);
    NarrowCaps caps;
    if (levels < 1 || !narrow_caps(ctx, ctx->num_sms, levels, caps)) return kexp::fail(ctx, "narrow apply: not applicable to this operator / width");
    if (M + 1 > kMaxTerms) return kexp::fail(ctx, "series too long");
    if (!ctx->d_nk) {
        KEXP_CUDA(cudaMalloc(&ctx->d_nk, sizeof(int) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_nrtol2, sizeof(double) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_npartials, sizeof(double) * 2 * kMaxW * (size_t)ctx->num_sms + 64));
    }
    if (!ctx->d_nktcols) {
        KEXP_CUDA(cudaMalloc(&ctx->d_nktcols, sizeof(int) * (size_t)std::max<int64_t>(ctx->Kt.nblocks, 1)));
        extract_block_cols_kernel<<<4 * ctx->num_sms, 256, 0, ctx->stream>>>(ctx->Kt.blocks, ctx->Kt.nblocks, ctx->d_nktcols);
    }
    // rotating publication buffers of the mass-solve iterates: n_s x 16 doubles each (the widest narrow panel)
    if (!ctx->d_nxg) KEXP_CUDA(cudaMalloc(&ctx->d_nxg, sizeof(double) * 4 * 16 * (size_t)ctx->n_s));
    // pageable sources: the driver stages them before the call returns, so the host vectors may be reused at once
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nk, kterm, sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nrtol2, rtol2, sizeof(double) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    const NarrowPlanDev& P = ctx->narrow_plan[levels - 1];
    NarrowArgs a;
    a.kt_browptr = ctx->Kt.browptr; a.kt_blocks = ctx->Kt.blocks; a.kt_rowmap = ctx->Kt.rowmap; a.kt_nbr = ctx->Kt.nbr;
    a.kt_cols = ctx->d_nktcols;
    a.lblocks = P.lblocks; a.lrowptr = P.lrowptr; a.cta_blk_ptr = P.cta_blk_ptr; a.cta_row_ptr = P.cta_row_ptr;
    a.halo_ptr = P.halo_ptr; a.halo_cols = P.halo_cols; a.halo1_cnt = P.halo1_cnt;
    a.Pinv = ctx->d_Pinv; a.s_nbr = ctx->As.nbr; a.levels = levels;
    a.r_rowptr = ctx->Rt.rowptr; a.r_col = ctx->Rt.col; a.r_val = ctx->Rt.val; a.Dblk = ctx->d_Dblk;
    a.n = (int)ctx->n; a.n_s = (int)ctx->n_s; a.n_b = (int)ctx->n_b; a.n_d = (int)ctx->n_d;
    a.SA = ctx->S0; a.SB = ctx->S1; a.Y = ctx->Y; a.T = ctx->T; a.X = ctx->X;
    for (int b = 0; b < 4; ++b) a.xg[b] = ctx->d_nxg + (size_t)b * 16 * ctx->n_s;
    a.tscale = ctx->d_tscale; a.coef = ctx->d_tscale + kMaxW; a.kterm = ctx->d_nk; a.rtol2 = ctx->d_nrtol2;
    a.M = M; a.mode = mode;
    a.basis = (mode == 1 || with_basis) ? ctx->d_basis : nullptr; a.basis_ld = ctx->basis_ld;
    a.alpha = 2.0 / (ctx->cheb_hi + ctx->cheb_lo);
    a.mu = (ctx->cheb_hi - ctx->cheb_lo) / (ctx->cheb_hi + ctx->cheb_lo);
    a.sc = ctx->d_sc; a.partials = ctx->d_npartials;
    a.barrier = reinterpret_cast<unsigned*>(ctx->d_npartials + 2 * kMaxW * (size_t)ctx->num_sms);
    KEXP_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), ctx->stream));
    a.cap_bpr = caps.bpr; a.cap_s_rows = caps.s_rows; a.cap_kt_rows = caps.kt_rows; a.cap_r_rows = caps.r_rows;
    a.cap_slots = caps.slots; a.cap_halo = caps.halo; a.xs_offset = 0;
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
    ctx->narrow_levels_used = levels;
    return 0;
}

}  // namespace kexp
