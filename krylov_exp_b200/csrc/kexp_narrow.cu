// kexp_narrow.cu -- the whole cheby_exp.apply (cheby_exp.py:110-165) of a NARROW panel (2, 4 or 8 sample columns) as ONE
// persistent cooperative kernel.
//
// Why: with few columns per GPU -- the per-GPU share of an averaging window split over many GPUs, the single-column
// recurrences of the shared-input forward pass and of the fused Clenshaw average, and the reference's own un-batched
// apply(x, y, t) -- an apply is a chain of ~25 terms x ~22 sparse products with a dependency on the neighbours after each
// one, and its cost is LATENCY, not bandwidth.  Five kernel launches per term (K~ product, mass solve, three row ranges of
// the recurrence update) each pay a launch boundary, re-stage the mass matrix in shared memory and serialise the gathers of
// a block row.  Here:
//   * one launch per apply: 1 CTA per SM, 256 threads with up to 255 registers; the CTA's rows of S (transposed), P^-1,
//     the index arrays of K~ and R and its send lists stay in shared memory for ALL terms and iterations;
//   * per term three phases separated by grid-wide barriers:  A: T = K~ S_{n-1} (depth rows updated in the epilogue)
//     | B: the mass solve, K-1 Chebyshev iterations, the edge-row recurrence update fused into the last one
//     | C: interior rows  w_I = T_I + R w_E  and their update;
//   * INSIDE the mass solve there is no grid barrier: the iterate lives in shared memory (own rows, two generations, plus
//     the halo rows the own rows reference), b and -- for up to 4 columns -- the thread's row of S and P^-1 in registers,
//     one thread per block row; the boundary rows of every new iterate travel to the neighbours' inboxes as flagged values
//     (flag-in-data, see the comment in front of ll_store), and a CTA waits only for the values it needs;
//   * data written by other CTAs is read with ld.global.cg / relaxed gpu-scope loads (L2, the coherence point), everything
//     else normally;
//   * the convergence record of every mass solve (||b - S x||/||b|| per column against the relaxed tolerance of its term) is
//     reduced in fixed order by CTA 0 behind the barrier that follows the solve;
//   * the kernel contains no CALL and no conditionally defined local array: both cost registers everywhere (comments at
//     narrow_check and in phase A; measurements in profiles/r02_summary.md section 3).
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
    // 1 CTA per SM, 256 threads, up to 255 registers each -- the mass solve keeps a whole block row (3 rows x W columns) of
    // b, S x and the new iterate in one thread's registers, for W <= 4 also the row of the matrix
    static constexpr int NT = 256;
    static constexpr int GK = (W == 2) ? 4 : 8;      // gathered blocks of K~ in flight per row
    static constexpr int SS = 3 * W + (W >= 4 ? 2 : 0);   // doubles per block-row slot of the shared-memory iterates (padded: the
                                                     // lanes of a warp read different slots; 3W doubles apart they would share banks)
};
constexpr int kNTmax = 256;
constexpr int kNWarpsMax = kNTmax / 32;
constexpr int kMaxOmega = 96;            // most Chebyshev iterations of one mass solve in the persistent kernel
constexpr int kRChunk = 9;               // nonzeros of a row of R in flight (interior dof x the 9 edge dofs of its cell)

struct NarrowArgs {
    // K~ (BSR 3x3, stored rows in locality order, rowmap = output block row)
    const int* kt_browptr; const double* kt_blocks; const int* kt_rowmap; int kt_nbr;
    // S as per-CTA LOCAL problems (build_narrow_plan): the blocks of a CTA's own rows, packed per CTA with the SLOT of each
    // block's column in the 10th double (slot s >= 0: own row s; s < 0: halo slot -s-1); lrowptr: local row pointers;
    // halo_cols: global block rows of the halo (sorted); srowptr / send_dst: the CTA's send entries [cta_send_ptr[c],
    // cta_send_ptr[c+1]): local source row and inbox slot (global numbering: halo_ptr[consumer] + its halo slot), sorted by slot
    const double* lblocks; const int* lrowptr; const int* cta_blk_ptr; const int* cta_row_ptr;
    const int* halo_ptr; const int* halo_cols; const int* srowptr; const int* cta_send_ptr; const int* send_dst;
    // local row order of a CTA: the rows that reference a halo slot first (cta_nb[c] of them), then the rows that need only
    // the CTA's own x; grow[s0 + lr] = global block row of local row lr (a permutation inside the CTA's range)
    const int* grow; const int* cta_nb;
    uint4* inbox; int inbox_slots;                           // two buffers of inbox_slots x 3W flagged doubles (zeroed before the launch)
    const double* Pinv; int s_nbr;
    // K~ re-laid out for coalesced loads (build_narrow_kt): the CTA's rows sorted by block count; kt2_cnt / kt2_orow: blocks and
    // output block row of every row in that order; groups of 32 rows, group g of CTA c starts at double2 index
    // kt2_gbase[kt2_grp_ptr[c] + g]; part p (0..4) of block slot q of row r of the group at  base + (5 q + p) 32 + r
    const double2* kt2; const int* kt2_cnt; const int* kt2_orow; const int* kt2_gbase; const int* kt2_grp_ptr;
    const int* kt2_cta_ptr;                                  // the CTA's rows of K~: [kt2_cta_ptr[c], kt2_cta_ptr[c+1]), cut by BLOCK count
    // R (CSR, n_b x n_s) or null; Dblk (n_d/3 blocks) or null
    const int* r_rowptr; const int* r_col; const double* r_val; const double* Dblk;
    int n, n_s, n_b, n_d;
    // panels [rows][W]
    double *SA, *SB, *Y, *T, *X;
    // per-panel tables (device): tscale[W], coef[(M+1)][kMaxW]; per-term iteration counts and squared tolerances
    const double* tscale; const double* coef; const int* kterm; const double* rtol2;
    int M;                               // highest term index
    int mode;                            // 0: forward recurrence y = sum a_n S_n;  1: Clenshaw (fused weighted average)
    double* basis; int basis_ld;         // mode 0: store column 0 of S_n (or null);  mode 1: z_m (required)
    double alpha, mu;                    // Chebyshev acceleration of the mass solve
    double omega[kMaxOmega];             // omega[k]: weight of iteration k (x_k -> x_{k+1}), k = 1 .. kMaxOmega-1
    CgScalars* sc; double* partials;     // [gridDim][2][W] block sums of ||r||^2, ||b||^2
    double* pairs;                       // convergence record handed to narrow_finish_kernel (4 doubles)
    unsigned* barrier;                   // grid barrier counter (zeroed before the launch)
    // shared-memory layout (element counts, maxima over the CTAs)
    int cap_bpr, cap_s_rows, cap_kt_rows, cap_r_rows, cap_halo, cap_send;     // cap_bpr: most blocks in a row of S
    int cap_kt_groups;                   // most 32-row groups of K~ in a CTA
    int xs_offset;                       // first double of the iterate buffers in dynamic shared memory
};

// The arguments live in constant memory (one slot per context, round-robin): every phase derives what it needs -- row
// ranges, shared-memory pointers, the constants of the step -- from blockIdx, the step number and this bank, so that
// nothing but three counters is live between the phases and each phase has the register file to itself.  (Passing the
// ADDRESS of a by-value kernel parameter around makes the compiler keep a per-thread copy in local memory: every pointer
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
#define KEXP_NARROW_INLINE 1       // 1: phases inlined into one kernel body (no subroutine calls at all: see narrow_check); 0: separate functions
#endif
#if KEXP_NARROW_INLINE
#define KEXP_PHASE_FN __device__ __forceinline__
#else
#define KEXP_PHASE_FN __device__ __noinline__
#endif
#ifndef KEXP_NARROW_TIMING
#define KEXP_NARROW_TIMING 0       // 1: thread 0 of a few CTAs accumulates clock64() per section and prints them (profiling builds only)
#endif
#if KEXP_NARROW_TIMING
__device__ long long g_ntim_dummy;
#define NTIM_DECL __shared__ long long s_tim[16]; if (threadIdx.x == 0) for (int i_ = 0; i_ < 16; ++i_) s_tim[i_] = 0;
#define NTIM_T0 long long t0_ = clock64();
#define NTIM_ADD(i) if (threadIdx.x == 0) { const long long t1_ = clock64(); s_tim[i] += t1_ - t0_; t0_ = t1_; }
#else
#define NTIM_DECL
#define NTIM_T0
#define NTIM_ADD(i)
#endif
// release side of the grid barrier: the CTA's writes (ordered before by bar.sync) must be visible before the arrive.  An
// acquire-release fence is enough; __threadfence() is the sequentially consistent one (MEMBAR.SC.GPU).
#ifndef KEXP_NARROW_FENCE_SC
#define KEXP_NARROW_FENCE_SC 0
#endif
#if KEXP_NARROW_FENCE_SC
#define KEXP_NARROW_FENCE() __threadfence()
#else
#define KEXP_NARROW_FENCE() asm volatile("fence.acq_rel.gpu;" ::: "memory")
#endif
__device__ __forceinline__ void narrow_barrier(cg::grid_group& grid, unsigned* counter, unsigned& target, long long* tim = nullptr) {
#if KEXP_NARROW_SYNC
#if KEXP_NARROW_TIMING
    long long c0 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) {
        long long c1 = clock64();
        target += gridDim.x;
        KEXP_NARROW_FENCE();
        long long c2 = clock64();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
        long long c3 = clock64();
        if (tim) { tim[0] += c1 - c0; tim[1] += c2 - c1; tim[2] += c3 - c2; tim[15] += 1; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && tim) tim[3] += clock64() - c0;
#else
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        KEXP_NARROW_FENCE();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
    }
    __syncthreads();
#endif
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

// S_{n-1} and S_{n-2} alternate between the two work panels; Clenshaw (mode 1) runs the terms m = M-1 ... 0
__device__ __forceinline__ StepCtx make_step(const NarrowArgs& a, int step) {
    StepCtx sx;
    sx.Sin = (step & 1) ? a.SA : a.SB;              // S_{n-1}
    sx.Snew = (step & 1) ? a.SB : a.SA;             // holds S_{n-2}, becomes S_n
    sx.step = step;
    sx.term = (a.mode == 0) ? step : a.M - step;
    sx.cs = (a.mode == 0) ? (step == 1 ? 1.0 : 2.0) : (sx.term == 0 ? 1.0 : 2.0);
    sx.has_old = (a.mode == 0) ? (step >= 2) : 1;
    sx.crow = a.coef + (size_t)((a.mode == 0) ? sx.term : 0) * kMaxW;
    return sx;
}

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

// Carve-up of the dynamic shared memory (every phase derives its pointers from the extern array itself, so that the
// compiler knows they are shared-memory addresses: LDS / STS instead of generic accesses).
struct NarrowSmem {
    double *sv, *sP, *sred;                       // S values (transposed: [9 * bpr][ell_ld]), block-Jacobi inverses (transposed: [9][ell_ld]), reduction scratch
    int *ss, *scnt, *sptr, *sdst, *kptr, *rptr;   // column slots ([bpr][ell_ld]), blocks per row, send entries (source row, inbox slot), K~: blocks per row (sorted order), row pointers of R
    int *korow, *kgb;                             // K~: output block row of every row (sorted order), base of every 32-row group
    int *sgrow;                                   // global block row of every local row
    double *xo0, *xo1, *xh;                       // iterates on the own rows (two), halo of the current iterate
    int ell_ld;
};
template <int W>
__device__ __forceinline__ NarrowSmem narrow_smem_carve(const NarrowArgs& a) {
    extern __shared__ double2 smem2[];
    NarrowSmem m;
    m.ell_ld = a.cap_s_rows;
    m.sv = reinterpret_cast<double*>(smem2);
    m.sP = m.sv + (size_t)9 * a.cap_bpr * m.ell_ld;
    m.sred = m.sP + 9 * (size_t)a.cap_s_rows;                                  // 2 x warps x W, then 2 x W x 32 (+ 4 W)
    m.ss = reinterpret_cast<int*>(m.sred + 2 * kNWarpsMax * W + 2 * W * 32 + 4 * W);
    m.scnt = m.ss + (size_t)a.cap_bpr * m.ell_ld;
    m.sptr = m.scnt + a.cap_s_rows;
    m.sdst = m.sptr + a.cap_send;
    m.kptr = m.sdst + a.cap_send;
    m.rptr = m.kptr + a.cap_kt_rows + 1;
    m.sgrow = m.rptr + a.cap_r_rows + 1;
    m.korow = m.sgrow + a.cap_s_rows;
    m.kgb = m.korow + a.cap_kt_rows;
    m.xo0 = reinterpret_cast<double*>(smem2) + a.xs_offset;                    // 16-byte aligned
    m.xo1 = m.xo0 + (size_t)NarrowCfg<W>::SS * a.cap_s_rows;
    m.xh = m.xo1 + (size_t)NarrowCfg<W>::SS * a.cap_s_rows;
    return m;
}

// ---- phase A: T = K~ S_{n-1}; depth rows (no solve) take their recurrence update in the epilogue.
// K~ is streamed from L2 / HBM once per term (37 MB at ref 5) and the rows of a warp are different rows: with the plain
// 80-byte blocks every 16-byte load of a warp touched 32 different lines and the phase was bound by the LSU request rate
// (30 k cycles per term at W = 2).  The copy used here stores, per group of 32 rows of similar length, part p of block
// slot q of all 32 rows side by side: the loads of a warp are contiguous.
template <int W>
KEXP_PHASE_FN void narrow_phase_a(int slot, int step) {
    const NarrowArgs& a = c_nargs[slot];
    const StepCtx sx = make_step(a, step);
    const int k0r = __ldg(a.kt2_cta_ptr + blockIdx.x), k1r = __ldg(a.kt2_cta_ptr + blockIdx.x + 1);
    const NarrowSmem smA = narrow_smem_carve<W>(a);
    const int* __restrict__ kcnt = smA.kptr;
    const int* __restrict__ korow = smA.korow;
    const int* __restrict__ kgb = smA.kgb;
    using C = PanelCfg<W, 2>;
    constexpr int LPB = C::LPB, BRW = C::BRW;
    constexpr int kNT = NarrowCfg<W>::NT, kNWarps = kNT / 32, kGatherK = NarrowCfg<W>::GK;
    constexpr int BPC = kNWarps * BRW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane % LPB, g = lane / LPB;
    const UpdCtx u = make_upd<W>(a, sx, l);
    const double* Sin = sx.Sin;
    const int row_eta = a.n_s + a.n_b;
    const int nrows = k1r - k0r;
    for (int base = 0; base < nrows; base += BPC) {
        const int lr = base + warp * BRW + g;
        if (lr < nrows) {
            const int nq = kcnt[lr];
            const int orow = korow[lr];
            const double2* gp = a.kt2 + kgb[lr >> 5] + (lr & 31);
            double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
            for (int q = 0; q < nq; q += kGatherK) {
                // (every element is defined on every path: a conditionally defined array looks live from the top of the kernel
                // to ptxas -- around the whole step loop -- and costs its registers everywhere, e.g. in the mass solve)
                double2 x[kGatherK][3], v8c[kGatherK];
#pragma unroll
                for (int j = 0; j < kGatherK; ++j) {
                    v8c[j] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int r = 0; r < 3; ++r) x[j][r] = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int j = 0; j < kGatherK; ++j)
                    if (q + j < nq) v8c[j] = __ldg(gp + (5 * (q + j) + 4) * 32);          // m22 and the block column
#pragma unroll
                for (int j = 0; j < kGatherK; ++j)
                    if (q + j < nq) {
                        const double* xr = Sin + (size_t)(3 * __double2loint(v8c[j].y)) * W + 2 * l;
#pragma unroll
                        for (int r = 0; r < 3; ++r) x[j][r] = ldcg2(xr + (size_t)r * W);
                    }
#pragma unroll
                for (int j = 0; j < kGatherK; ++j)
                    if (q + j < nq) {
                        const double2* bp = gp + (5 * (q + j)) * 32;
                        const double2 v01 = __ldg(bp), v23 = __ldg(bp + 32), v45 = __ldg(bp + 64), v67 = __ldg(bp + 96);
                        const double m[3][3] = {{v01.x, v01.y, v23.x}, {v23.y, v45.x, v45.y}, {v67.x, v67.y, v8c[j].x}};
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

// ---- phase B: the mass solve of one term.
// Shared-memory state (per CTA), in block-row SLOTS of 3 rows x W doubles: two iterate arrays over the OWN rows (xo0 / xo1:
// x_k and x_{k-1}, the latter overwritten by x_{k+1}) and ONE halo array (x_k on the rows of other CTAs that the own rows
// reference).  b (own rows) lives in REGISTERS: the row <-> thread assignment is fixed for the whole solve (NP passes of
// NT / LPB block rows).
//
// Halo exchange WITHOUT a grid barrier ("LL" protocol: flag-in-data).  Every double travels as 16 bytes
// {lo32, flag, hi32, flag}; 8-byte aligned words are single-copy atomic, so a reader that sees both flags equal to the
// number of the exchange it waits for has the value -- no fence, no counter, no separate "data is there" round trip.  The
// thread that computes x_{k+1} of a boundary row stores it straight into the inbox slot of every CTA that has that row in its
// halo (send lists precomputed on the host); the consumer's threads poll their own inbox elements and copy them to shared
// memory.  Two inbox buffers alternate (exchange number parity): a CTA can only be ONE exchange ahead of a neighbour,
// because its next iterate needs that neighbour's current one (the sparsity of S is structurally symmetric; checked on the
// host).  Exchange numbers increase monotonically through the whole launch; the inbox is zeroed before the launch.
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned flag) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// wait for exchange `flag` on the CTA's inbox elements [0, ne) and copy them to the halo array (shared memory)
template <int NT, int RW, int SS>    // RW = 3W doubles per inbox slot, SS doubles per shared-memory slot
__device__ __forceinline__ void ll_recv(const uint4* __restrict__ inbox, double* __restrict__ xh, int ne, unsigned flag) {
    constexpr int U = 4;                               // elements in flight per thread
    for (int e0 = threadIdx.x; e0 < ne; e0 += NT * U) {
        uint4 v[U];
        bool pending;
        do {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (e0 + u * NT < ne) v[u] = ll_load(inbox + e0 + u * NT);
            pending = false;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (e0 + u * NT < ne && (v[u].y != flag || v[u].w != flag)) pending = true;
        } while (pending);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (e0 + u * NT < ne) { const int e = e0 + u * NT; xh[(e / RW) * SS + (e % RW)] = __hiloint2double((int)v[u].z, (int)v[u].x); }
    }
}

// S x on one block row, all W columns in one thread: everything from shared memory.  The matrix is stored TRANSPOSED
// (element e of block q of local row lr at sv[(9 q + e) ell_ld + lr], its column slot at ss[q ell_ld + lr]: the rows of a warp
// read consecutive words -- no bank conflicts).  Slot s >= 0: own row s; s < 0: halo slot -s-1.
// the block row of x in slot s (s >= 0: own row s; s < 0: halo slot -s-1)
template <int W>
__device__ __forceinline__ void narrow_load_x(int s, const double* __restrict__ xo, const double* __restrict__ xh, double2 (&x)[3][W / 2]) {
    constexpr int NV = W / 2, SS = NarrowCfg<W>::SS;
    const double* xr = (s >= 0 ? xo + s * SS : xh + (-s - 1) * SS);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int v = 0; v < NV; ++v) x[r][v] = *reinterpret_cast<const double2*>(xr + r * W + 2 * v);
}
// one 3x3 block (row-major m[9]) times a block row of x
template <int W>
__device__ __forceinline__ void narrow_block_mul(const double (&m)[9], const double2 (&x)[3][W / 2], double2 (&acc)[3][W / 2]) {
    constexpr int NV = W / 2;
    // same order as the wide-panel kernels: for each block, columns r = 0, 1, 2 of the 3x3 block in turn
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                acc[i][v].x = fma(m[3 * i + r], x[r][v].x, acc[i][v].x);
                acc[i][v].y = fma(m[3 * i + r], x[r][v].y, acc[i][v].y);
            }
}
template <int W>
__device__ __forceinline__ void narrow_block_apply(const double (&m)[9], int s, const double* __restrict__ xo, const double* __restrict__ xh,
                                                   double2 (&acc)[3][W / 2]) {
    constexpr int NV = W / 2;
    double2 x[3][NV];
    narrow_load_x<W>(s, xo, xh, x);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                acc[i][v].x = fma(m[3 * i + r], x[r][v].x, acc[i][v].x);
                acc[i][v].y = fma(m[3 * i + r], x[r][v].y, acc[i][v].y);
            }
}
template <int W>
__device__ __forceinline__ void narrow_block_fma(const double* __restrict__ sv, const int* __restrict__ ss, int ell_ld, int lr, int q,
                                                 const double* __restrict__ xo, const double* __restrict__ xh, double2 (&acc)[3][W / 2]) {
    const double* vp = sv + (size_t)(9 * q) * ell_ld + lr;
    double m[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) m[e] = vp[e * ell_ld];
    narrow_block_apply<W>(m, ss[q * ell_ld + lr], xo, xh, acc);
}

// publish the boundary rows of an iterate (in shared memory) to the CTAs that have them in their halo: the CTA's send entries
// (source row, inbox slot of the consumer) are walked by all threads, consecutive threads -> consecutive doubles of a row,
// so that the flagged 16-byte stores of a warp are contiguous
template <int NT, int RW, int SS>
__device__ __forceinline__ void narrow_send(uint4* __restrict__ inbox_par, const int* __restrict__ srow, const int* __restrict__ sdst, int nsend,
                                            const double* __restrict__ xo, unsigned flag) {
    for (int e = threadIdx.x; e < nsend * RW; e += NT) {
        const int ent = e / RW, j = e % RW;
        ll_store(inbox_par + (size_t)sdst[ent] * RW + j, xo[srow[ent] * SS + j], flag);
    }
}

// The whole Chebyshev semi-iteration of one term on the CTA's own rows:  x_1 = alpha P^-1 b, then K - 1 iterations
//   x_{k+1} = x_{k-1} + omega_k (x_k + alpha P^-1 (b - S x_k) - x_{k-1});
// the last one also forms ||b - S x_{K-1}||^2, writes x_K to a.X and applies the recurrence update of the edge rows.
// One thread per block row (NP passes of NT rows), all W columns of the row in its registers.  MREG (one pass, at most 5
// blocks per row, W <= 4): the thread's matrix row -- 5 blocks of S and its block of P^-1, 54 doubles -- stays in registers
// for the whole solve as well: at W = 2 the matrix is 60 % of the shared-memory wavefronts of an iteration (75 KB per CTA
// at ref 5 against 50 KB of x), and the iteration is bound by exactly those.
// Leaves the block sums of ||r||^2 and ||b||^2 per column in a.partials.  Returns the next exchange number.
template <int W, int NP, bool MREG>
KEXP_PHASE_FN unsigned narrow_solve(int slot, int step, unsigned xid
#if KEXP_NARROW_TIMING
                                    , long long* s_tim
#endif
) {
    const NarrowArgs& a = c_nargs[slot];
    const NarrowSmem sm = narrow_smem_carve<W>(a);
    const StepCtx sx = make_step(a, step);
    const int K = a.kterm[step];
    int s0, s1;
    cta_range(a.s_nbr, s0, s1);
    const int n_own = s1 - s0;
    const int h0 = __ldg(a.halo_ptr + blockIdx.x), n_h1 = __ldg(a.halo_ptr + blockIdx.x + 1) - h0;
    const int nsend = __ldg(a.cta_send_ptr + blockIdx.x + 1) - __ldg(a.cta_send_ptr + blockIdx.x);
    constexpr int NT = NarrowCfg<W>::NT, kNWarps = NT / 32, NV = W / 2, SS = NarrowCfg<W>::SS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ell_ld = sm.ell_ld;
    const size_t inbox_half = (size_t)a.inbox_slots * (3 * W);
    const uint4* my_inbox = a.inbox + (size_t)h0 * (3 * W);
    const int ne = n_h1 * 3 * W;
    double2 bv[NP][3][NV];
    double* cur = sm.xo0;                               // x_k on the own rows
    double* prev = sm.xo1;                              // x_{k-1} on the own rows -> x_{k+1}
    constexpr int kRegBlocks = MREG ? 5 : 1;
    double mreg[kRegBlocks][9], preg[9];
    int sreg[kRegBlocks];
#pragma unroll
    for (int q = 0; q < kRegBlocks; ++q) {              // (defined on every path)
        sreg[q] = 0;
#pragma unroll
        for (int e = 0; e < 9; ++e) mreg[q][e] = 0.0;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) preg[e] = 0.0;
    if (MREG && (int)threadIdx.x < n_own) {
        const int lr = threadIdx.x;
        const int nq = sm.scnt[lr];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const bool have = q < nq;                    // shorter rows: zero blocks on the row's own slot
            sreg[q] = have ? sm.ss[q * ell_ld + lr] : lr;
#pragma unroll
            for (int e = 0; e < 9; ++e) mreg[q][e] = have ? sm.sv[(size_t)(9 * q + e) * ell_ld + lr] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 9; ++e) preg[e] = sm.sP[e * ell_ld + lr];
    }
    NTIM_T0
    // ---- start: b <- T (global), x_1 = alpha P^-1 b, x_0 = 0
    double accb[W];
#pragma unroll
    for (int c = 0; c < W; ++c) accb[c] = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int lr = p * NT + threadIdx.x;
        if (lr < n_own) {
            const size_t o = (size_t)(3 * sm.sgrow[lr]) * W;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    bv[p][i][v] = ldcg2(a.T + o + (size_t)i * W + 2 * v);
                    accb[2 * v] = fma(bv[p][i][v].x, bv[p][i][v].x, accb[2 * v]);
                    accb[2 * v + 1] = fma(bv[p][i][v].y, bv[p][i][v].y, accb[2 * v + 1]);
                }
            double P[3][3];
#pragma unroll
            for (int e = 0; e < 9; ++e) P[e / 3][e % 3] = MREG ? preg[e] : sm.sP[e * ell_ld + lr];
            double2 z[3][NV];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    z[i][v].x = a.alpha * fma(P[i][0], bv[p][0][v].x, fma(P[i][1], bv[p][1][v].x, P[i][2] * bv[p][2][v].x));
                    z[i][v].y = a.alpha * fma(P[i][0], bv[p][0][v].y, fma(P[i][1], bv[p][1][v].y, P[i][2] * bv[p][2][v].y));
                    *reinterpret_cast<double2*>(cur + lr * SS + i * W + 2 * v) = z[i][v];
                    *reinterpret_cast<double2*>(prev + lr * SS + i * W + 2 * v) = make_double2(0.0, 0.0);
                }
        }
    }
    __syncthreads();                                    // x_1 on the own rows complete
    narrow_send<NT, 3 * W, SS>(a.inbox + (size_t)(xid & 1u) * inbox_half, sm.sptr, sm.sdst, nsend, cur, xid);
    // warp sums of ||b||^2 per column (fixed order) -> reduction scratch
#pragma unroll
    for (int c = 0; c < W; ++c) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) accb[c] += __shfl_xor_sync(0xffffffffu, accb[c], off);
        if (lane == 0) sm.sred[(1 * kNWarps + warp) * W + c] = accb[c];
    }
    NTIM_ADD(6)
    // The rows of the CTA are ordered: the nb rows that reference a halo slot first, then the rows that need only the CTA's
    // own x_k.  For W = 2 the latter are computed BEFORE the halo of x_k is waited for: their work (1.2 k cycles) hides in the
    // latency of the exchange (the neighbours' values were sent one __syncthreads ago and need ~1.9 k cycles to be seen).
    // Wider panels keep the plain order: their own-row work is longer than the latency and would only delay the threads'
    // share of the polling (measured at W = 4: 2.78 against 2.60 ms per apply).
    constexpr bool kOverlap = (W == 2);
    const int nb = __ldg(a.cta_nb + blockIdx.x);
    for (int it = 1; it < K; ++it) {
        const double omega = a.omega[it];
        const bool last = (it == K - 1);
        double accr[W];                                     // ||b - S x_k||^2 per column: formed by the last iteration only
#pragma unroll
        for (int c = 0; c < W; ++c) accr[c] = 0.0;
        // x_{it+1} on the rows of one class (halo_rows: the rows that reference halo slots / the others)
        auto rows = [&](const bool halo_rows) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int lr = p * NT + threadIdx.x;
                if (lr < n_own && (lr < nb) == halo_rows) {
                    const int nq = MREG ? 5 : sm.scnt[lr];
                    double2 acc[3][NV];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v) acc[i][v] = make_double2(0.0, 0.0);
                    if (MREG) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) narrow_block_apply<W>(mreg[q], sreg[q], cur, sm.xh, acc);
                    } else if (W <= 4 && nq == 5) {          // BDM2 edge rows: the edge and the four other edges of its two cells
#pragma unroll
                        for (int q = 0; q < 5; ++q) narrow_block_fma<W>(sm.sv, sm.ss, ell_ld, lr, q, cur, sm.xh, acc);
                    } else {                                 // (8 columns: one block's 72 products at a time fill the registers)
#pragma unroll 1
                        for (int q = 0; q < nq; ++q) narrow_block_fma<W>(sm.sv, sm.ss, ell_ld, lr, q, cur, sm.xh, acc);
                    }
                    double P[3][3];
#pragma unroll
                    for (int e = 0; e < 9; ++e) P[e / 3][e % 3] = MREG ? preg[e] : sm.sP[e * ell_ld + lr];
                    // acc <- r = b - S x_k
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            acc[i][v].x = bv[p][i][v].x - acc[i][v].x; acc[i][v].y = bv[p][i][v].y - acc[i][v].y;
                            if (last) {
                                accr[2 * v] = fma(acc[i][v].x, acc[i][v].x, accr[2 * v]);
                                accr[2 * v + 1] = fma(acc[i][v].y, acc[i][v].y, accr[2 * v + 1]);
                            }
                        }
                    double2 xn[3][NV];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const double2 xo = *reinterpret_cast<const double2*>(cur + lr * SS + i * W + 2 * v);
                            const double2 xmv = *reinterpret_cast<const double2*>(prev + lr * SS + i * W + 2 * v);
                            double2 z;
                            z.x = fma(P[i][0], acc[0][v].x, fma(P[i][1], acc[1][v].x, P[i][2] * acc[2][v].x));
                            z.y = fma(P[i][0], acc[0][v].y, fma(P[i][1], acc[1][v].y, P[i][2] * acc[2][v].y));
                            xn[i][v].x = fma(omega, fma(a.alpha, z.x, xo.x) - xmv.x, xmv.x);
                            xn[i][v].y = fma(omega, fma(a.alpha, z.y, xo.y) - xmv.y, xmv.y);
                        }
                    if (!last) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int v = 0; v < NV; ++v) *reinterpret_cast<double2*>(prev + lr * SS + i * W + 2 * v) = xn[i][v];
                    } else {
                        const int gr = sm.sgrow[lr];
                        const size_t o = (size_t)(3 * gr) * W;
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const UpdCtx u = make_upd<W>(a, sx, v);
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                *reinterpret_cast<double2*>(a.X + o + (size_t)i * W + 2 * v) = xn[i][v];
                                row_update(u, o + (size_t)i * W + 2 * v, (size_t)(3 * gr + i), xn[i][v]);     // edge rows: W = x_K
                            }
                        }
                    }
                }
            }
        };
        if (kOverlap) rows(false);                      // needs the own rows of x_it only (complete: barrier of the last send)
        NTIM_ADD(1)
        // halo of x_it (exchange xid); the halo array is free: every thread has finished the halo rows of the previous
        // iteration (the barrier in front of its sends)
        ll_recv<NT, 3 * W, SS>(my_inbox + (size_t)(xid & 1u) * inbox_half, sm.xh, ne, xid);
        __syncthreads();
        NTIM_ADD(7)
        ++xid;
        rows(true);
        if (!kOverlap) rows(false);
        NTIM_ADD(2)
        if (last) {
#pragma unroll
            for (int c = 0; c < W; ++c) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) accr[c] += __shfl_xor_sync(0xffffffffu, accr[c], off);
                if (lane == 0) sm.sred[(0 * kNWarps + warp) * W + c] = accr[c];
            }
        }
        NTIM_ADD(9)
        if (!last) {
            __syncthreads();                            // x_{it+1} on the own rows complete (and nobody reads the halo of x_it any more)
            narrow_send<NT, 3 * W, SS>(a.inbox + (size_t)(xid & 1u) * inbox_half, sm.sptr, sm.sdst, nsend, prev, xid);
        }
        NTIM_ADD(3)
        double* tsw = cur; cur = prev; prev = tsw;
    }
    // block sums of ||r||^2 and ||b||^2 per column (fixed order): the warp sums in turn
    {
        __syncthreads();
        if (threadIdx.x < 2 * W) {
            const int q = threadIdx.x / W, c = threadIdx.x % W;
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kNWarps; ++w) t += sm.sred[(q * kNWarps + w) * W + c];
            a.partials[((size_t)blockIdx.x * 2 + q) * W + c] = t;
        }
    }
    NTIM_ADD(10)
    return xid;
}

// ---- phase C: interior rows  w_I = T_I + R x_K  and their recurrence update
template <int W>
KEXP_PHASE_FN void narrow_phase_c(int slot, int step) {
    const NarrowArgs& a = c_nargs[slot];
    const StepCtx sx = make_step(a, step);
    int r0, r1;
    cta_range(a.n_b, r0, r1);
    const int* __restrict__ rptr = narrow_smem_carve<W>(a).rptr;
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
            for (int j = 0; j < kRChunk; ++j) { x[j] = make_double2(0.0, 0.0); v[j] = 0.0; }      // defined on every path (see phase A)
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

// ---- convergence record of one solve (CTA 0, fixed summation order over the per-CTA block sums).  No division and no
// square root in here: their slow paths are subroutine CALLs, and a kernel that contains calls gets a far worse register
// allocation everywhere (ptxas then keeps call-free regions in the caller-saved registers only: the matrix rows of
// narrow_solve3 were spilled with 100 registers idle).  The worst column of a solve is found by cross-multiplication;
// pairs[0..1]: (||r||^2, ||b||^2) of the worst column of the last solve, pairs[2..3]: of the worst solve that missed its
// tolerance -- narrow_finish_kernel turns them into relative residuals behind the launch.
template <int W>
KEXP_PHASE_FN void narrow_check(int slot, int step, double* pairs, int* s_bad) {
    const NarrowArgs& a = c_nargs[slot];
    double* fin = narrow_smem_carve<W>(a).sred + 2 * kNWarpsMax * W;
    const int K = a.kterm[step];
    const double rtol2 = a.rtol2[step];
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
    if (threadIdx.x < W) {
        const double rr = tot[threadIdx.x], bb = tot[W + threadIdx.x];
        if (bb > 0.0 && rr > rtol2 * bb) atomicOr(s_bad, 1);
        a.sc->rr[threadIdx.x] = rr; a.sc->bb[threadIdx.x] = bb;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double wr = 0.0, wb = 1.0;                        // worst column: largest rr / bb
        for (int c = 0; c < W; ++c) {
            const double rr = tot[c], bb = tot[W + c];
            if (bb > 0.0 && rr * wb > wr * bb) { wr = rr; wb = bb; }
        }
        pairs[0] = wr; pairs[1] = wb;
        CgScalars* sc = a.sc;
        sc->iters = K;
        sc->total_iters += (unsigned long long)K;
        sc->total_solves += 1ull;
        if (K > sc->max_iters) sc->max_iters = K;
        sc->done = *s_bad ? 0 : 1;
        if (*s_bad) {
            sc->unconverged += 1;
            if (wr * pairs[3] > pairs[2] * wb) { pairs[2] = wr; pairs[3] = wb; }
        }
    }
}

// relative residuals of the convergence record (one thread, behind the persistent kernel)
__global__ void narrow_finish_kernel(const double* __restrict__ pairs, CgScalars* __restrict__ sc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        sc->max_rel_res = sqrt(pairs[0] / pairs[1]);
        if (pairs[2] > 0.0) sc->worst_rel_res = fmax(sc->worst_rel_res, sqrt(pairs[2] / pairs[3]));
    }
}

// ---- staging of the CTA's constant data (its rows of S transposed, P^-1, send lists, row pointers of K~ and R) and the
// prologue of the recurrence
template <int W>
KEXP_PHASE_FN void narrow_prologue(int slot) {
    const NarrowArgs& a = c_nargs[slot];
    constexpr int kNT = NarrowCfg<W>::NT;
    constexpr int LPR = W / 2;                         // lanes per scalar row
    constexpr int RPC = kNT / LPR;                     // scalar rows per CTA pass
    const NarrowSmem sm = narrow_smem_carve<W>(a);
    const int ell_ld = sm.ell_ld;
    double* sv = sm.sv; double* sP = sm.sP;
    int* ss = sm.ss; int* scnt = sm.scnt; int* sptr = sm.sptr; int* sdst = sm.sdst; int* kptr = sm.kptr; int* rptr = sm.rptr;

    int s0, s1, r0, r1;
    cta_range(a.s_nbr, s0, s1);
    const int k0r = __ldg(a.kt2_cta_ptr + blockIdx.x), k1r = __ldg(a.kt2_cta_ptr + blockIdx.x + 1);
    cta_range(a.n_b, r0, r1);
    const int n_own = s1 - s0;
    const int h0 = __ldg(a.halo_ptr + blockIdx.x), n_h1 = __ldg(a.halo_ptr + blockIdx.x + 1) - h0;

    // ---- stage the constant data of this CTA
    {
        const int b0 = __ldg(a.cta_blk_ptr + blockIdx.x);
        const double* gsrc = a.lblocks + (size_t)10 * b0;
        const int row_base = __ldg(a.cta_row_ptr + blockIdx.x);
        const int* lrp = a.lrowptr + row_base;
        for (int lr = threadIdx.x; lr < n_own; lr += kNT) {                    // one row per thread: transpose its blocks
            const int q0 = __ldg(lrp + lr), q1 = __ldg(lrp + lr + 1);
            scnt[lr] = q1 - q0;
            for (int q = q0; q < q1; ++q) {
                const double* bp = gsrc + (size_t)10 * q;
#pragma unroll
                for (int e = 0; e < 9; ++e) sv[(size_t)(9 * (q - q0) + e) * ell_ld + lr] = __ldg(bp + e);
                ss[(q - q0) * ell_ld + lr] = __double2loint(__ldg(bp + 9));
            }
        }
        for (int i = threadIdx.x; i < n_own; i += kNT) sm.sgrow[i] = __ldg(a.grow + s0 + i);
        for (int i = threadIdx.x; i < 9 * n_own; i += kNT)
            sP[(i % 9) * ell_ld + i / 9] = __ldg(a.Pinv + (size_t)9 * __ldg(a.grow + s0 + i / 9) + i % 9);
        const int sd0 = __ldg(a.cta_send_ptr + blockIdx.x), nsd = __ldg(a.cta_send_ptr + blockIdx.x + 1) - sd0;
        for (int i = threadIdx.x; i < nsd; i += kNT) { sptr[i] = __ldg(a.srowptr + sd0 + i); sdst[i] = __ldg(a.send_dst + sd0 + i); }
        for (int i = threadIdx.x; i < k1r - k0r; i += kNT) { kptr[i] = __ldg(a.kt2_cnt + k0r + i); sm.korow[i] = __ldg(a.kt2_orow + k0r + i); }
        {
            const int g0 = __ldg(a.kt2_grp_ptr + blockIdx.x), ng = __ldg(a.kt2_grp_ptr + blockIdx.x + 1) - g0;
            for (int i = threadIdx.x; i < ng; i += kNT) sm.kgb[i] = __ldg(a.kt2_gbase + g0 + i);
        }
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
}

template <int W, int NP, bool MREG>
__global__ void __launch_bounds__(NarrowCfg<W>::NT, 1)
cheb_apply_narrow_kernel(const int slot) {
    const NarrowArgs& a = c_nargs[slot];
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_bad;
    __shared__ double s_pairs[4];
    if (threadIdx.x == 0) { s_pairs[0] = 0.0; s_pairs[1] = 1.0; s_pairs[2] = 0.0; s_pairs[3] = 1.0; }
    narrow_prologue<W>(slot);
    unsigned bar_target = 0;
    NTIM_DECL
    narrow_barrier(grid, a.barrier, bar_target);

    unsigned xid = 1;                                    // number of the next halo exchange (the inbox holds zeros)
    // Every phase derives what it needs (row ranges, shared-memory pointers, the constants of the step) from blockIdx, the
    // step number and the constant bank: nothing but the three counters below is live between them.
    for (int step = 1; step <= a.M; ++step) {
        NTIM_T0
        narrow_phase_a<W>(slot, step);
        NTIM_ADD(4)
        narrow_barrier(grid, a.barrier, bar_target);
        NTIM_ADD(5)
        // ---- mass solve on the own rows of S: Chebyshev semi-iteration, K - 1 iterations behind x_1; halos travel through the
        // flagged inboxes, no grid barrier inside
#if KEXP_NARROW_TIMING
        xid = narrow_solve<W, NP, MREG>(slot, step, xid, s_tim);
        if (threadIdx.x == 0) s_tim[15] += a.kterm[step] - 1;
        t0_ = clock64();
#else
        xid = narrow_solve<W, NP, MREG>(slot, step, xid);
#endif
        narrow_barrier(grid, a.barrier, bar_target);           // x_K (in a.X), the edge rows of S_n and the block sums are visible
        NTIM_ADD(11)
        if (blockIdx.x == 0) narrow_check<W>(slot, step, s_pairs, &s_bad);
        if (a.n_b > 0) {
            NTIM_ADD(12)
            narrow_phase_c<W>(slot, step);
            NTIM_ADD(13)
            narrow_barrier(grid, a.barrier, bar_target);       // S_n complete
            NTIM_ADD(14)
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {           // (CTA 0 wrote them itself: program order)
        for (int k = 0; k < 4; ++k) a.pairs[k] = s_pairs[k];
    }
#if KEXP_NARROW_TIMING == 2
    if (threadIdx.x == 0) printf("NTIMA cta %3d A %lld barA %lld rows %d\n", (int)blockIdx.x, s_tim[4], s_tim[5],
                                 (int)(a.kt2_cta_ptr[blockIdx.x + 1] - a.kt2_cta_ptr[blockIdx.x]));
#endif
#if KEXP_NARROW_TIMING
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1))
        printf("NTIM cta %3d W %d iters %lld | rows(own) %lld rows(halo) %lld | A %lld barA %lld start %lld | recv %lld iter %lld | red %lld barK %lld chk %lld C %lld barC %lld | n_own %d n_h1 %d\n",
               (int)blockIdx.x, W, s_tim[15], s_tim[1], s_tim[2], s_tim[4], s_tim[5], s_tim[6], s_tim[7], s_tim[9],
               s_tim[10], s_tim[11], s_tim[12], s_tim[13], s_tim[14], (int)a.cap_s_rows, (int)a.cap_halo);
#endif
}

// ------------------------------------------------------------------------------------------------ host side
struct NarrowCaps { int bpr, s_rows, kt_rows, r_rows, halo, send, kt_groups; };

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
    b += sizeof(int) * ((size_t)c.bpr * c.s_rows + 2 * (size_t)c.s_rows + 2 * (size_t)c.send + 2 * (size_t)c.kt_rows + 1 + c.r_rows + 1 + c.kt_groups);
    return (b + 15) & ~(size_t)15;
}

template <int W>
static size_t narrow_smem(const NarrowCaps& c) {
    return narrow_smem_fixed<W>(c) + sizeof(double) * NarrowCfg<W>::SS * ((size_t)2 * c.s_rows + c.halo);
}

// gather kernel of build_narrow_kt: one thread per (row in the sorted order, block slot, 16-byte part)
__global__ void kt_transpose_kernel(const double2* __restrict__ blocks, const int* __restrict__ browptr, const int* __restrict__ srow,
                                    const int* __restrict__ cnt, const int* __restrict__ rbase, int nbr, int max_cnt, double2* __restrict__ out) {
    const int64_t total = (int64_t)nbr * max_cnt * 5;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / (max_cnt * 5)), rem = (int)(i % (max_cnt * 5)), q = rem / 5, p = rem % 5;
        if (q < cnt[row]) out[(size_t)rbase[row] + (size_t)(5 * q + p) * 32] = blocks[(size_t)5 * (browptr[srow[row]] + q) + p];
    }
}

// K~ for phase A of the persistent kernel (see NarrowArgs): per CTA the rows of its range sorted by block count (a warp then
// holds rows of equal length), cut into groups of 32 rows, every group stored transposed and padded to its longest row.
// The index arrays are built on the host from the row pointers, the blocks are moved by kt_transpose_kernel.
static int build_narrow_kt(kexp_ctx* ctx, int G) {
    NarrowPlanDev& P = ctx->narrow_plan;
    if (P.kt2) return 0;
    const std::vector<int>& kb = ctx->Kt.h_browptr;
    if (kb.empty()) return kexp::fail(ctx, "narrow apply: host copy of the row pointers of K~ not available");
    const int nbr = (int)kb.size() - 1;
    std::vector<int> srow((size_t)nbr), cnt((size_t)nbr), orow((size_t)nbr), rbase((size_t)nbr), gbase, grp_ptr((size_t)G + 1, 0);
    // the CTAs' row ranges: cut by block count (a warp holds rows of equal length, so a CTA's time follows its blocks)
    std::vector<int> cta_ptr((size_t)G + 1, 0);
    for (int c = 1; c < G; ++c) {
        const int64_t want = (int64_t)kb[nbr] * c / G;
        cta_ptr[c] = (int)(std::lower_bound(kb.begin(), kb.end(), (int)want) - kb.begin());
        cta_ptr[c] = std::min(std::max(cta_ptr[c], cta_ptr[c - 1]), nbr);
    }
    cta_ptr[G] = nbr;
    size_t total = 0;                                  // double2 elements of the transposed copy
    int max_cnt = 1, cap_groups = 0, cap_rows = 0;
    for (int c = 0; c < G; ++c) {
        const int r0 = cta_ptr[c], r1 = cta_ptr[c + 1];
        cap_rows = std::max(cap_rows, r1 - r0);
        std::vector<int> idx;
        for (int r = r0; r < r1; ++r) idx.push_back(r);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return kb[x + 1] - kb[x] > kb[y + 1] - kb[y]; });
        for (int i = 0; i < r1 - r0; ++i) {
            const int r = idx[i];
            srow[(size_t)r0 + i] = r;
            cnt[(size_t)r0 + i] = kb[r + 1] - kb[r];
            orow[(size_t)r0 + i] = ctx->Kt.h_rowmap.empty() ? r : ctx->Kt.h_rowmap[r];
            max_cnt = std::max(max_cnt, cnt[(size_t)r0 + i]);
        }
        const int ng = (r1 - r0 + 31) / 32;
        for (int g = 0; g < ng; ++g) {
            const int gmax = cnt[(size_t)r0 + 32 * g];  // sorted: the first row of the group is its longest
            gbase.push_back((int)total);
            for (int i = 32 * g; i < std::min(32 * g + 32, r1 - r0); ++i) rbase[(size_t)r0 + i] = (int)total + (i - 32 * g);
            total += (size_t)gmax * 5 * 32;
        }
        grp_ptr[c + 1] = (int)gbase.size();
        cap_groups = std::max(cap_groups, ng);
    }
    if (total >= ((size_t)1 << 31)) return kexp::fail(ctx, "narrow apply: K~ too large for the persistent kernel");
    int* d_srow = nullptr; int* d_rbase = nullptr;
    auto up = [&](int** dst, const std::vector<int>& v) -> bool {
        if (cudaMalloc(dst, sizeof(int) * std::max<size_t>(v.size(), 1)) != cudaSuccess) return true;
        return cudaMemcpy(*dst, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess;
    };
    if (up(&P.kt2_cnt, cnt) || up(&P.kt2_orow, orow) || up(&P.kt2_gbase, gbase) || up(&P.kt2_grp_ptr, grp_ptr) || up(&P.kt2_cta_ptr, cta_ptr) ||
        up(&d_srow, srow) || up(&d_rbase, rbase) ||
        cudaMalloc(&P.kt2, sizeof(double2) * std::max<size_t>(total, 1)) != cudaSuccess) {
        cudaFree(d_srow); cudaFree(d_rbase);
        return kexp::fail(ctx, "out of device memory for the narrow-apply copy of K~");
    }
    cudaMemsetAsync(P.kt2, 0, sizeof(double2) * std::max<size_t>(total, 1), ctx->stream);
    kt_transpose_kernel<<<8 * ctx->num_sms, 256, 0, ctx->stream>>>(reinterpret_cast<const double2*>(ctx->Kt.blocks), ctx->Kt.browptr, d_srow, P.kt2_cnt,
                                                                   d_rbase, nbr, max_cnt, reinterpret_cast<double2*>(P.kt2));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_srow); cudaFree(d_rbase);
    if (e != cudaSuccess) return kexp::fail(ctx, std::string("narrow apply: transposing K~ failed: ") + cudaGetErrorString(e));
    P.cap_kt_groups = cap_groups; P.cap_kt_rows = cap_rows;
    return 0;
}

// Per-CTA local problems of S for a grid of G CTAs (see NarrowArgs): the CTA's own rows with shared-memory slot numbers,
// its halo (block columns referenced outside the own range, sorted) and its send lists (which inbox slots of which CTAs
// receive each own row).  Built once per operator.  Fails (-> multi-launch path) if the sparsity of S is not structurally
// symmetric: the flag-in-data exchange relies on "I send to whoever sends to me".
static int build_narrow_plan(kexp_ctx* ctx, int G) {
    NarrowPlanDev& P = ctx->narrow_plan;
    if (P.G == G && P.lblocks) return 0;
    if (P.G == -G) return 1;                                   // known not to be applicable
    const std::vector<int>& sb = ctx->As.h_browptr;
    const std::vector<double>& blk = ctx->As.h_blocks;
    if (sb.empty() || blk.empty()) return kexp::fail(ctx, "narrow apply: host copy of S not available");
    const int nbr = (int)sb.size() - 1;
    auto colof = [&](int k) { int64_t bits; memcpy(&bits, &blk[(size_t)k * 10 + 9], 8); return (int)bits; };
    auto owner_of = [&](int row) {
        const int per = nbr / G, rem = nbr % G;
        const int cut = rem * (per + 1);
        return row < cut ? row / (per + 1) : rem + (row - cut) / std::max(per, 1);
    };
    std::vector<double> lb;
    std::vector<int> lrow, cta_blk((size_t)G + 1, 0), cta_row((size_t)G + 1, 0), hptr((size_t)G + 1, 0), hcols;
    std::vector<std::vector<int>> halos((size_t)G);
    std::vector<int> grow((size_t)nbr, 0), lpos((size_t)nbr, 0), cta_nb((size_t)G, 0);     // local row order (see NarrowArgs)
    NarrowCaps caps{0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < G; ++c) {
        int r0, r1;
        range_of(nbr, c, G, r0, r1);
        std::vector<int>& h1 = halos[c];
        for (int k = sb[r0]; k < sb[r1]; ++k) { const int col = colof(k); if (col < r0 || col >= r1) h1.push_back(col); }
        std::sort(h1.begin(), h1.end());
        h1.erase(std::unique(h1.begin(), h1.end()), h1.end());
        // local order: the rows that reference a halo column first, then the rows that need only the CTA's own rows
        {
            int nbc = 0;
            for (int pass = 0; pass < 2; ++pass)
                for (int r = r0; r < r1; ++r) {
                    bool needs = false;
                    for (int k = sb[r]; k < sb[r + 1] && !needs; ++k) { const int col = colof(k); needs = (col < r0 || col >= r1); }
                    if (needs == (pass == 0)) { lpos[r] = nbc; grow[(size_t)r0 + nbc] = r; ++nbc; }
                    if (pass == 0 && r == r1 - 1) cta_nb[c] = nbc;
                }
        }
        const size_t blk0 = lb.size() / 10;
        lrow.push_back(0);
        for (int lr = 0; lr < r1 - r0; ++lr) {
            const int r = grow[(size_t)r0 + lr];
            for (int k = sb[r]; k < sb[r + 1]; ++k) {
                const size_t o = lb.size();
                lb.insert(lb.end(), blk.begin() + (size_t)k * 10, blk.begin() + (size_t)k * 10 + 10);
                const int col = colof(k);
                int64_t slot;
                if (col >= r0 && col < r1) slot = lpos[col];
                else slot = -(int64_t)(std::lower_bound(h1.begin(), h1.end(), col) - h1.begin()) - 1;
                memcpy(&lb[o + 9], &slot, 8);
            }
            lrow.push_back((int)(lb.size() / 10 - blk0));
            caps.bpr = std::max(caps.bpr, sb[r + 1] - sb[r]);
        }
        cta_blk[c + 1] = (int)(lb.size() / 10);
        cta_row[c + 1] = (int)lrow.size();
        hcols.insert(hcols.end(), h1.begin(), h1.end());
        hptr[c + 1] = (int)hcols.size();
        caps.s_rows = std::max(caps.s_rows, r1 - r0);
        caps.halo = std::max(caps.halo, (int)h1.size());
    }
    // send lists: row `row` of CTA p goes to inbox slot hptr[c] + h of every CTA c with halos[c][h] == row
    std::vector<std::vector<std::pair<int, int>>> sends((size_t)G);           // per producer: (local row, inbox slot)
    std::vector<std::vector<int>> talks_to((size_t)G);
    for (int c = 0; c < G; ++c)
        for (size_t h = 0; h < halos[c].size(); ++h) {
            const int row = halos[c][h], p = owner_of(row);
            int p0, p1;
            range_of(nbr, p, G, p0, p1);
            if (row < p0 || row >= p1) return kexp::fail(ctx, "narrow apply: internal error (row owner)");
            sends[p].push_back({lpos[row], hptr[c] + (int)h});
            talks_to[p].push_back(c);
        }
    bool symmetric = true;
    for (int p = 0; p < G && symmetric; ++p) {
        std::sort(talks_to[p].begin(), talks_to[p].end());
        talks_to[p].erase(std::unique(talks_to[p].begin(), talks_to[p].end()), talks_to[p].end());
    }
    for (int p = 0; p < G && symmetric; ++p)
        for (int c : talks_to[p])
            if (!std::binary_search(talks_to[c].begin(), talks_to[c].end(), p)) { symmetric = false; break; }
    if (!symmetric) { P.G = -G; return 1; }
    std::vector<int> srow, cta_send((size_t)G + 1, 0), sdst;                  // per send entry: source row (local), inbox slot
    for (int p = 0; p < G; ++p) {
        std::sort(sends[p].begin(), sends[p].end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.second < y.second; });
        for (const auto& e : sends[p]) { srow.push_back(e.first); sdst.push_back(e.second); }
        cta_send[p + 1] = (int)sdst.size();
        caps.send = std::max(caps.send, (int)sends[p].size());
    }
    free_narrow_plan(P);
    auto up = [&](auto** dst, const auto& v) -> bool {
        using T = std::remove_reference_t<decltype(**dst)>;
        if (cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1)) != cudaSuccess) return true;
        return cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess;
    };
    if (up(&P.lblocks, lb) || up(&P.lrowptr, lrow) || up(&P.cta_blk_ptr, cta_blk) || up(&P.cta_row_ptr, cta_row) ||
        up(&P.halo_ptr, hptr) || up(&P.halo_cols, hcols) || up(&P.srowptr, srow) || up(&P.cta_send_ptr, cta_send) || up(&P.send_dst, sdst) ||
        up(&P.grow, grow) || up(&P.cta_nb, cta_nb))
        return kexp::fail(ctx, "out of device memory for the narrow-apply plan");
    P.G = G; P.cap_blocks = caps.bpr; P.cap_rows = caps.s_rows; P.cap_halo = caps.halo; P.cap_send = caps.send;
    P.inbox_slots = (int)hcols.size();
    return 0;
}

void free_narrow_plan(NarrowPlanDev& P) {
    cudaFree(P.lblocks); cudaFree(P.lrowptr); cudaFree(P.cta_blk_ptr); cudaFree(P.cta_row_ptr);
    cudaFree(P.halo_ptr); cudaFree(P.halo_cols); cudaFree(P.srowptr); cudaFree(P.cta_send_ptr); cudaFree(P.send_dst);
    cudaFree(P.grow); cudaFree(P.cta_nb);
    cudaFree(P.kt2); cudaFree(P.kt2_cnt); cudaFree(P.kt2_orow); cudaFree(P.kt2_gbase); cudaFree(P.kt2_grp_ptr); cudaFree(P.kt2_cta_ptr);
    cudaFree(P.inbox);
    P = NarrowPlanDev();
}

static bool narrow_caps(kexp_ctx* ctx, int G, NarrowCaps& c) {
    const std::vector<int>& kb = ctx->Kt.h_browptr;
    const std::vector<int>& rp = ctx->Rt.h_rowptr;
    if (ctx->As.h_browptr.empty() || kb.empty()) return false;
    if (build_narrow_plan(ctx, G) || build_narrow_kt(ctx, G)) return false;
    const NarrowPlanDev& P = ctx->narrow_plan;
    c = NarrowCaps{P.cap_blocks, P.cap_rows, P.cap_kt_rows, 0, P.cap_halo, P.cap_send, P.cap_kt_groups};
    for (int i = 0; i < G; ++i) {
        int a, b;
        if (!rp.empty()) {
            range_of((int)rp.size() - 1, i, G, a, b);
            c.r_rows = std::max(c.r_rows, b - a);
        }
    }
    return true;
}

static int narrow_threads(int W) { return W == 2 ? NarrowCfg<2>::NT : NarrowCfg<4>::NT; }
// passes of NT block rows a CTA needs for its own rows (one thread per block row; b of every pass lives in registers: 1 or 2)
static int narrow_passes(int W, const NarrowCaps& c) {
    const int nt = narrow_threads(W);
    return (c.s_rows + nt - 1) / nt;
}

template <int W, int NP, bool MREG>
static int launch_narrow_w(kexp_ctx* ctx, NarrowArgs& a, const NarrowCaps& caps) {
    auto k = cheb_apply_narrow_kernel<W, NP, MREG>;
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
        default: return (size_t)1 << 40;
    }
}

// Is the persistent kernel applicable to a panel of width W?  Chebyshev mass solve, both matrices in 3x3 blocks, S
// structurally symmetric, the CTA's local problem + iterate buffers within the shared memory an SM offers, at most two
// passes over the own rows.
bool narrow_eligible(kexp_ctx* ctx, int W) {
    if (!ctx->use_narrow || W > ctx->narrow_max_w || W > 8) return false;
    if (ctx->solver != 1 || !ctx->Kt.bsr() || !ctx->As.bsr()) return false;
    if (ctx->n_b > 0 && (ctx->Rt.rowptr == nullptr)) return false;
    if (ctx->n % 3 || ctx->n_s % 3 || ctx->n_b % 3) return false;
    const size_t limit = 227u * 1024u - 1024u;
    NarrowCaps caps;
    if (!narrow_caps(ctx, ctx->num_sms, caps)) return false;
    return narrow_smem_for(W, caps) <= limit && narrow_passes(W, caps) <= 2;
}

// Enqueue one whole apply of the panel held in ctx->S0 (mode 0: result in ctx->Y; mode 1: Clenshaw, result in the buffer
// returned through *result).  kterm / rtol2: per-step iteration counts and squared tolerances (host, M+1 entries).
int launch_narrow_apply(kexp_ctx* ctx, int W, int M, int mode, const int* kterm, const double* rtol2, bool with_basis, const double** result) {
    NarrowCaps caps;
    if (!narrow_eligible(ctx, W) || !narrow_caps(ctx, ctx->num_sms, caps)) return kexp::fail(ctx, "narrow apply: not applicable to this operator / width");
    if (M + 1 > kMaxTerms) return kexp::fail(ctx, "series too long");
    if (!ctx->d_nk) {
        KEXP_CUDA(cudaMalloc(&ctx->d_nk, sizeof(int) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_nrtol2, sizeof(double) * kMaxTerms));
        KEXP_CUDA(cudaMalloc(&ctx->d_npartials, sizeof(double) * 2 * kMaxW * (size_t)ctx->num_sms + 64));
    }
    NarrowPlanDev& P = ctx->narrow_plan;
    // halo inboxes: two buffers of (all halo slots) x 3 x 8 flagged doubles (sized for the widest narrow panel), zeroed per launch
    const size_t inbox_bytes = sizeof(uint4) * 2 * (size_t)std::max(P.inbox_slots, 1) * 3 * 8;
    if (!P.inbox) KEXP_CUDA(cudaMalloc(&P.inbox, inbox_bytes));
    KEXP_CUDA(cudaMemsetAsync(P.inbox, 0, sizeof(uint4) * 2 * (size_t)std::max(P.inbox_slots, 1) * 3 * W, ctx->stream));
    // pageable sources: the driver stages them before the call returns, so the host vectors may be reused at once
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nk, kterm, sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_nrtol2, rtol2, sizeof(double) * (size_t)(M + 1), cudaMemcpyHostToDevice, ctx->stream));
    NarrowArgs a;
    a.kt_browptr = ctx->Kt.browptr; a.kt_blocks = ctx->Kt.blocks; a.kt_rowmap = ctx->Kt.rowmap; a.kt_nbr = ctx->Kt.nbr;
    a.kt2 = reinterpret_cast<const double2*>(P.kt2); a.kt2_cnt = P.kt2_cnt; a.kt2_orow = P.kt2_orow; a.kt2_gbase = P.kt2_gbase;
    a.kt2_grp_ptr = P.kt2_grp_ptr; a.kt2_cta_ptr = P.kt2_cta_ptr;
    a.lblocks = P.lblocks; a.lrowptr = P.lrowptr; a.cta_blk_ptr = P.cta_blk_ptr; a.cta_row_ptr = P.cta_row_ptr;
    a.halo_ptr = P.halo_ptr; a.halo_cols = P.halo_cols; a.srowptr = P.srowptr; a.cta_send_ptr = P.cta_send_ptr; a.send_dst = P.send_dst;
    a.grow = P.grow; a.cta_nb = P.cta_nb;
    a.inbox = reinterpret_cast<uint4*>(P.inbox); a.inbox_slots = P.inbox_slots;
    a.Pinv = ctx->d_Pinv; a.s_nbr = ctx->As.nbr;
    a.r_rowptr = ctx->Rt.rowptr; a.r_col = ctx->Rt.col; a.r_val = ctx->Rt.val; a.Dblk = ctx->d_Dblk;
    a.n = (int)ctx->n; a.n_s = (int)ctx->n_s; a.n_b = (int)ctx->n_b; a.n_d = (int)ctx->n_d;
    a.SA = ctx->S0; a.SB = ctx->S1; a.Y = ctx->Y; a.T = ctx->T; a.X = ctx->X;
    a.tscale = ctx->d_tscale; a.coef = ctx->d_tscale + kMaxW; a.kterm = ctx->d_nk; a.rtol2 = ctx->d_nrtol2;
    a.M = M; a.mode = mode;
    a.basis = (mode == 1 || with_basis) ? ctx->d_basis : nullptr; a.basis_ld = ctx->basis_ld;
    a.alpha = 2.0 / (ctx->cheb_hi + ctx->cheb_lo);
    a.mu = (ctx->cheb_hi - ctx->cheb_lo) / (ctx->cheb_hi + ctx->cheb_lo);
    a.omega[0] = 1.0;
    for (int k = 1; k < kMaxOmega; ++k) a.omega[k] = (k == 1) ? 1.0 / (1.0 - 0.5 * a.mu * a.mu) : 1.0 / (1.0 - 0.25 * a.mu * a.mu * a.omega[k - 1]);
    for (int n = 1; n <= M; ++n)
        if (kterm[n] > kMaxOmega) return kexp::fail(ctx, "narrow apply: too many mass-solve iterations");
    a.sc = ctx->d_sc; a.partials = ctx->d_npartials;
    a.barrier = reinterpret_cast<unsigned*>(ctx->d_npartials + 2 * kMaxW * (size_t)ctx->num_sms);
    a.pairs = ctx->d_npartials + 2 * kMaxW * (size_t)ctx->num_sms + 2;
    KEXP_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), ctx->stream));
    a.cap_bpr = caps.bpr; a.cap_s_rows = caps.s_rows; a.cap_kt_rows = caps.kt_rows; a.cap_r_rows = caps.r_rows;
    a.cap_halo = caps.halo; a.cap_send = caps.send; a.cap_kt_groups = caps.kt_groups; a.xs_offset = 0;
    const int np = narrow_passes(W, caps);
    const bool mreg = ctx->narrow_mreg && np == 1 && caps.bpr <= 5 && W <= 4;      // matrix rows in registers
    int rc = 1;
    switch (W * 100 + np * 10 + (mreg ? 1 : 0)) {
        case 211: rc = launch_narrow_w<2, 1, true>(ctx, a, caps); break;
        case 210: rc = launch_narrow_w<2, 1, false>(ctx, a, caps); break;
        case 220: rc = launch_narrow_w<2, 2, false>(ctx, a, caps); break;
        case 411: rc = launch_narrow_w<4, 1, true>(ctx, a, caps); break;
        case 410: rc = launch_narrow_w<4, 1, false>(ctx, a, caps); break;
        case 420: rc = launch_narrow_w<4, 2, false>(ctx, a, caps); break;
        case 810: rc = launch_narrow_w<8, 1, false>(ctx, a, caps); break;
        default: return kexp::fail(ctx, "narrow apply: unsupported panel width");
    }
    if (rc) return rc;
    narrow_finish_kernel<<<1, 32, 0, ctx->stream>>>(a.pairs, ctx->d_sc);
    if (result) *result = (M % 2 == 0) ? ctx->S0 : ctx->S1;       // the buffer written by the last step
    ctx->solve_path = 3;
    return 0;
}

}  // namespace kexp
