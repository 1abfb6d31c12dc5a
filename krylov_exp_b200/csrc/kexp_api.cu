// kexp_api.cu -- host side of libkexp.so: C ABI (include/kexp.h), operator upload,
// batched PCG mass solve driver, Chebyshev term loop captured as one CUDA graph,
// device-resident batches, statistics.  Kernels: kexp_kernels.cuh.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include "kexp_internal.h"

using namespace kexp;

namespace kexp {

int fail(kexp_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return 1;
}

int pick_W(int B) {
    int W = 2;
    while (W < B && W < kMaxW) W <<= 1;
    return W;
}

void free_csr(DevCsr& d) {
    cudaFree(d.rowptr); cudaFree(d.col); cudaFree(d.val);
    cudaFree(d.browptr); cudaFree(d.blocks); cudaFree(d.rowmap);
    d = DevCsr();
}

// CSR -> BSR(3x3) with 80-byte blocks {9 values row-major, block column in the low word of
// the 10th double}.  Only kept when padding zeros cost < 40 % (BDM2/DG1 matrices: 0 %).
static int build_bsr3(kexp_ctx* ctx, DevCsr& d, int nrows, int ncols, const int32_t* rp, const int32_t* ci, const double* v, bool force) {
    if (nrows % 3 || ncols % 3 || nrows == 0) return 0;
    const char* env = getenv("KEXP_NO_BSR");
    if (!force && env && env[0] == '1') return 0;
    const int nbr = nrows / 3;
    std::vector<int> browptr((size_t)nbr + 1, 0);
    std::vector<double> blocks;
    blocks.reserve((size_t)(rp[nrows] / 9 + nbr) * 10);
    std::vector<int> cols;
    int max_bpr = 1;
    bool all_diag = true;
    for (int b = 0; b < nbr; ++b) {
        cols.clear();
        for (int k = rp[3 * b]; k < rp[3 * b + 3]; ++k) cols.push_back(ci[k] / 3);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        const size_t base = blocks.size();
        blocks.resize(base + cols.size() * 10, 0.0);
        for (size_t j = 0; j < cols.size(); ++j) {
            int64_t bits = (int64_t)cols[j];
            memcpy(&blocks[base + j * 10 + 9], &bits, sizeof(double));
        }
        for (int i = 0; i < 3; ++i)
            for (int k = rp[3 * b + i]; k < rp[3 * b + i + 1]; ++k) {
                const size_t j = std::lower_bound(cols.begin(), cols.end(), ci[k] / 3) - cols.begin();
                blocks[base + j * 10 + 3 * i + (ci[k] % 3)] += v[k];
            }
        if (force && nrows == ncols) {
            // diagonal block first: the solve kernels take the row's own x from the gathered diagonal block
            const size_t jd = std::lower_bound(cols.begin(), cols.end(), b) - cols.begin();
            if (jd < cols.size() && cols[jd] == b) {
                if (jd != 0)
                    for (int e = 0; e < 10; ++e) std::swap(blocks[base + jd * 10 + e], blocks[base + e]);
            } else {
                all_diag = false;
            }
        }
        browptr[b + 1] = browptr[b] + (int)cols.size();
        max_bpr = std::max(max_bpr, (int)cols.size());
    }
    const int64_t nblocks = browptr[nbr];
    if (!force && (double)nblocks * 9.0 > 1.4 * (double)rp[nrows]) return 0;
    // Rectangular-coupled matrices (K~: edge / interior / depth rows of one cell gather the same column blocks but
    // live far apart): store the block rows sorted by their smallest block column, so that a CTA's contiguous range
    // re-uses gathered rows out of L1; rowmap gives the output block row of each stored row.
    std::vector<int> rowmap;
    const char* ro = getenv("KEXP_NO_ROW_ORDER");
    if (!force && !(ro && ro[0] == '1')) {
        std::vector<int> key((size_t)nbr), order((size_t)nbr);
        for (int b = 0; b < nbr; ++b) {
            int mn = ncols;
            for (int k = browptr[b]; k < browptr[b + 1]; ++k) { int64_t bits; memcpy(&bits, &blocks[(size_t)k * 10 + 9], 8); mn = std::min(mn, (int)bits); }
            key[b] = mn; order[b] = b;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return key[a] < key[b2]; });
        std::vector<int> bp2((size_t)nbr + 1, 0);
        std::vector<double> bl2(blocks.size());
        size_t w = 0;
        for (int i = 0; i < nbr; ++i) {
            const int b = order[i];
            const size_t cnt = (size_t)(browptr[b + 1] - browptr[b]);
            memcpy(&bl2[w * 10], &blocks[(size_t)browptr[b] * 10], cnt * 10 * sizeof(double));
            w += cnt;
            bp2[i + 1] = (int)w;
        }
        browptr.swap(bp2); blocks.swap(bl2); rowmap.swap(order);
    }
    KEXP_CUDA(cudaMalloc(&d.browptr, sizeof(int) * (size_t)(nbr + 1)));
    KEXP_CUDA(cudaMalloc(&d.blocks, sizeof(double) * 10 * (size_t)std::max<int64_t>(nblocks, 1)));
    KEXP_CUDA(cudaMemcpy(d.browptr, browptr.data(), sizeof(int) * (size_t)(nbr + 1), cudaMemcpyHostToDevice));
    KEXP_CUDA(cudaMemcpy(d.blocks, blocks.data(), sizeof(double) * 10 * (size_t)nblocks, cudaMemcpyHostToDevice));
    if (!rowmap.empty()) {
        KEXP_CUDA(cudaMalloc(&d.rowmap, sizeof(int) * (size_t)nbr));
        KEXP_CUDA(cudaMemcpy(d.rowmap, rowmap.data(), sizeof(int) * (size_t)nbr, cudaMemcpyHostToDevice));
    }
    d.nbr = nbr; d.nblocks = nblocks; d.max_bpr = max_bpr;
    d.h_browptr = browptr;
    d.h_rowmap = rowmap;
    if (force && nrows == ncols) d.h_blocks = blocks;
    d.diag_first = force && nrows == ncols && all_diag;
    if (max_bpr > kSpmmSmemBlocks) { cudaFree(d.browptr); cudaFree(d.blocks); d.browptr = nullptr; d.blocks = nullptr; d.h_browptr.clear(); }
    return 0;
}

int upload_csr(kexp_ctx* ctx, DevCsr& d, int nrows, int ncols, const int32_t* rp, const int32_t* ci, const double* v, bool force_bsr) {
    free_csr(d);
    d.nrows = nrows; d.ncols = ncols;
    if (nrows == 0 || rp == nullptr) { d.nrows = 0; return 0; }
    d.nnz = rp[nrows];
    for (int64_t k = 0; k < d.nnz; ++k)
        if (ci[k] < 0 || ci[k] >= ncols) return fail(ctx, "CSR column index out of range");
    KEXP_CUDA(cudaMalloc(&d.rowptr, sizeof(int) * (size_t)(nrows + 1)));
    KEXP_CUDA(cudaMalloc(&d.col, sizeof(int) * (size_t)std::max<int64_t>(d.nnz, 1)));
    KEXP_CUDA(cudaMalloc(&d.val, sizeof(double) * (size_t)std::max<int64_t>(d.nnz, 1)));
    KEXP_CUDA(cudaMemcpy(d.rowptr, rp, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyHostToDevice));
    KEXP_CUDA(cudaMemcpy(d.col, ci, sizeof(int) * (size_t)d.nnz, cudaMemcpyHostToDevice));
    KEXP_CUDA(cudaMemcpy(d.val, v, sizeof(double) * (size_t)d.nnz, cudaMemcpyHostToDevice));
    d.h_rowptr.assign(rp, rp + nrows + 1);
    return build_bsr3(ctx, d, nrows, ncols, rp, ci, v, force_bsr);
}

}  // namespace kexp

// ------------------------------------------------------------------ launch plumbing
namespace {

template <typename K>
int max_grid(kexp_ctx* ctx, K kernel, int dyn_smem = 0) {
    const void* key = (const void*)kernel;
    auto it = ctx->occ_cache.find(key);
    if (it != ctx->occ_cache.end()) return it->second;
    if (dyn_smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, dyn_smem);
    occ = std::max(1, std::min(occ, 8));
    if (dyn_smem > 0 && ctx->spmm_blocks_per_sm > 0) occ = std::min(occ, ctx->spmm_blocks_per_sm);   // experiment knob
    int g = occ * ctx->num_sms;
    ctx->occ_cache[key] = g;
    return g;
}

// Launch on the active lane's stream; with PDL on (and outside timing mode) the kernel may begin its
// predecessor-independent prologue while the previous kernel drains (see pdl_wait in the kernels).
template <typename... KArgs, typename... Args>
void launch_k(kexp_ctx* ctx, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (ctx->use_pdl && !ctx->timing) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline int blocks_for(int64_t groups, int lpr) {   // blocks needed so every group has one thread set
    int64_t per_block = kThreads / lpr;
    return (int)std::max<int64_t>(1, (groups + per_block - 1) / per_block);
}

cudaEvent_t next_event(kexp_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

struct TimedLaunch {   // RAII: event pair around one launch when kernel timing is on
    kexp_ctx* ctx; int cls;
    TimedLaunch(kexp_ctx* c, int k) : ctx(c), cls(k) {
        ctx->st_launches++;
        if (ctx->timing) {
            int idx = (int)ctx->ev_used;
            cudaEventRecord(next_event(ctx), ctx->stream);
            ctx->ev_pending.push_back({cls, idx});
        }
    }
    ~TimedLaunch() {
        if (ctx->timing) cudaEventRecord(next_event(ctx), ctx->stream);
    }
};

void drain_events(kexp_ctx* ctx) {
    if (!ctx->timing) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& pr : ctx->ev_pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev_pool[pr.second], ctx->ev_pool[pr.second + 1]);
        ctx->cls[pr.first].ms += ms;
        ctx->cls[pr.first].launches++;
    }
    ctx->ev_pending.clear();
    ctx->ev_used = 0;
}

void lane_store_cur(kexp_ctx* ctx, LaneState& L) {
    L.stream = ctx->stream; L.curW = ctx->curW;
    L.S0 = ctx->S0; L.S1 = ctx->S1; L.Y = ctx->Y; L.T = ctx->T; L.X = ctx->X; L.P = ctx->P; L.Q = ctx->Q;
    L.d_sc = ctx->d_sc; L.d_partials = ctx->d_partials; L.d_counter = ctx->d_counter; L.d_tscale = ctx->d_tscale;
    L.h_pin = ctx->h_pin; L.graphs.swap(ctx->graphs);
    L.base_iters = ctx->base_iters; L.base_solves = ctx->base_solves;
}

void lane_load(kexp_ctx* ctx, LaneState& L) {
    ctx->stream = L.stream; ctx->curW = L.curW;
    ctx->S0 = L.S0; ctx->S1 = L.S1; ctx->Y = L.Y; ctx->T = L.T; ctx->X = L.X; ctx->P = L.P; ctx->Q = L.Q;
    ctx->d_sc = L.d_sc; ctx->d_partials = L.d_partials; ctx->d_counter = L.d_counter; ctx->d_tscale = L.d_tscale;
    ctx->h_pin = L.h_pin; ctx->graphs.swap(L.graphs);
    ctx->base_iters = L.base_iters; ctx->base_solves = L.base_solves;
}

int alloc_lane_scalars(kexp_ctx* ctx);      // below (uses the active-lane fields)

// Make lane i the active one (allocating its stream / scalars on first use).
int swap_lane(kexp_ctx* ctx, int i) {
    if (i == ctx->cur_lane) return 0;
    lane_store_cur(ctx, ctx->lane_store[ctx->cur_lane]);
    LaneState& L = ctx->lane_store[i];
    const bool fresh = !L.allocated;
    lane_load(ctx, L);
    ctx->cur_lane = i;
    if (fresh) {
        L.allocated = true;
        if (alloc_lane_scalars(ctx)) return 1;
    }
    return 0;
}

int ensure_panels(kexp_ctx* ctx, int W) {
    if (W <= ctx->curW) return 0;
    cudaStreamSynchronize(ctx->stream);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.second);
    ctx->graphs.clear();
    double** bufs[] = {&ctx->S0, &ctx->S1, &ctx->Y, &ctx->T, &ctx->X, &ctx->P, &ctx->Q};
    for (auto b : bufs) { cudaFree(*b); *b = nullptr; }
    const size_t full = sizeof(double) * (size_t)ctx->n * W, part = sizeof(double) * (size_t)std::max<int64_t>(ctx->n_s, 1) * W;
    KEXP_CUDA(cudaMalloc(&ctx->S0, full));
    KEXP_CUDA(cudaMalloc(&ctx->S1, full));
    KEXP_CUDA(cudaMalloc(&ctx->Y, full));
    KEXP_CUDA(cudaMalloc(&ctx->T, full));
    KEXP_CUDA(cudaMalloc(&ctx->X, part));
    KEXP_CUDA(cudaMalloc(&ctx->P, part));
    KEXP_CUDA(cudaMalloc(&ctx->Q, part));
    ctx->curW = W;
    return 0;
}

int ensure_stage(kexp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return 0;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr; ctx->stage_bytes = 0;
    KEXP_CUDA(cudaMalloc(&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
    return 0;
}

// staging buffer of the host uploads: used on the copy stream only, so that an upload never waits for work of the compute stream
static int ensure_stage_up(kexp_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stage_up_bytes) return 0;
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->d_stage_up);
    ctx->d_stage_up = nullptr; ctx->stage_up_bytes = 0;
    KEXP_CUDA(cudaMalloc(&ctx->d_stage_up, bytes));
    ctx->stage_up_bytes = bytes;
    return 0;
}

// The scratch input batch of the host-pointer entry points is read by the apply / fused-average implementations only: they
// record this event behind their last read, and the next upload into the batch waits for it instead of for everything the
// compute stream holds (e.g. the forward half of an averaged evaluation, under which the backward inputs can then cross PCIe).
static int mark_consumed(kexp_ctx* ctx, kexp_batch* in) {
    if (in != ctx->tmp_in) return 0;
    if (!in->ev_consumed) KEXP_CUDA(cudaEventCreateWithFlags(&in->ev_consumed, cudaEventDisableTiming));
    KEXP_CUDA(cudaEventRecord(in->ev_consumed, ctx->stream));
    return 0;
}

// Y = A X (MODE 0) or the CG product q = A p with the fused p.q reduction (MODE 1); BSR(3x3)
// kernel when the matrix has that structure, generic CSR kernel otherwise.
// Measured on the bench step (profiles/r01e_summary.md): 1 -> 5990, 2 -> 5810, 3 -> 5930, 4 -> 5590, 5 -> 5380 applications/s.
// Explicit gather-ahead costs registers (spills at 5) and buys nothing: the compiler's own schedule of the
// five-block row with one block per group is the fastest.
#ifndef KEXP_GB
#define KEXP_GB 1
#endif
constexpr int kSpmmGB = KEXP_GB;                 // gathered blocks in flight per block row
// Gathered blocks per group for panels of up to 8 columns.  Measured at ref 5 x 8 columns (profiles/r02_summary.md): 1 -> 131 us,
// 3 -> 131 us, 4 -> 149 us per mass solve -- more gathers in flight buy nothing there either (5 spill at 128 registers).
#ifndef KEXP_GB_NARROW
#define KEXP_GB_NARROW 1
#endif
#ifndef KEXP_GB_NARROW_MAXW
#define KEXP_GB_NARROW_MAXW 8
#endif
constexpr int solve_gb(int W) { return W <= KEXP_GB_NARROW_MAXW ? KEXP_GB_NARROW : KEXP_GB; }
#ifndef KEXP_GB0
#define KEXP_GB0 3
#endif
constexpr int kSpmm0GB = KEXP_GB0;               // the same for the plain product T = K~ S
constexpr int kCptVec = 2;                  // sample columns per thread in the streaming vector kernels

// fu (MODE 0 only): recurrence update to apply in the epilogue to the rows from fused_row0 on (those that need no solve)
template <int W, int MODE>
void launch_spmm(kexp_ctx* ctx, const DevCsr& A, const double* X, double* Y, FusedUpd fu = FusedUpd{nullptr, nullptr, nullptr, nullptr, 0.0, 0},
                 int fused_row0 = 0) {
    constexpr int LPR = W / 2, RPW = 32 / LPR;
    if (A.bsr()) {
        auto k = bsr3_spmm_kernel<W, MODE, (MODE == 0 ? kSpmm0GB : kSpmmGB)>;
        constexpr int per_cta = PanelCfg<W, kSpmmCpt>::BPC;
        int grid = std::min(max_grid(ctx, k, kSpmmSmemBytes), std::max(1, (A.nbr + per_cta - 1) / per_cta));
        launch_k(ctx, k, grid, kThreads, kSpmmSmemBytes, A.browptr, A.blocks, X, Y, A.nbr, A.max_bpr, ctx->T, ctx->d_Pinv,
                 (const int*)A.rowmap, ctx->d_sc, ctx->d_partials, ctx->d_counter, fu, fused_row0, (const double*)ctx->d_Dblk);
    } else {
        // MODE 1 (the CG product) always has a BSR copy: n_s is a multiple of 3 by contract
        auto k = csr_spmm_kernel<W>;
        int grid = std::min(max_grid(ctx, k), std::max(1, (A.nrows + RPW * 8 - 1) / (RPW * 8)));
        k<<<grid, kThreads, 0, ctx->stream>>>(A.rowptr, A.col, A.val, X, Y, A.nrows);
    }
}

// ---- one batched application of M^-1 K on the work panels: input `in` (n x W),
// leaves T (rows needing no solve), X (solved rows). W is a template parameter.
// fu: recurrence update of the solved rows to fuse into the last iteration of the Chebyshev mass solve (fu.S != nullptr);
// returns true when it was fused (the caller then skips the separate update launch for these rows).
template <int W>
bool enqueue_operator(kexp_ctx* ctx, const double* in, double relax = 1.0, FusedUpd fu = FusedUpd{nullptr, nullptr, nullptr, nullptr, 0.0, 0}) {
    if (!ctx->use_fused_update) fu.S = nullptr;
    constexpr int LPB = PanelCfg<W, kCptVec>::LPB;
    const int ns = (int)ctx->n_s;
    const int red_grid = ctx->red_blocks_per_sm * ctx->num_sms;   // reduction kernels: few, fat blocks keep the last-block sum short
    // rows that need no solve take their recurrence update in the epilogue of the K~ product (BSR kernel only)
    const bool depth_fused = fu.S != nullptr && ctx->use_fused_update && ctx->Kt.bsr() && ctx->n_d > 0 && (ctx->n_s + ctx->n_b) % 3 == 0;
    ctx->depth_rows_fused = depth_fused;
    {   // T = Kt * in
        TimedLaunch tl(ctx, KC_SPMM_K);
        if (depth_fused) launch_spmm<W, 0>(ctx, ctx->Kt, in, ctx->T, fu, (int)(ctx->n_s + ctx->n_b));
        else launch_spmm<W, 0>(ctx, ctx->Kt, in, ctx->T);
    }
    const int ntri = ns / 3;
    if (ctx->solver == 1) {
        // Chebyshev semi-iteration: one fused kernel per iteration, no inner products
        const double alpha = 2.0 / (ctx->cheb_hi + ctx->cheb_lo);
        const double mu = (ctx->cheb_hi - ctx->cheb_lo) / (ctx->cheb_hi + ctx->cheb_lo);
        // relaxed tolerance of this Chebyshev term (see kexp_set_cheby): rtol*relax needs ln(relax)/acosh(sigma) fewer iterations
        const double sigma1 = (ctx->cheb_hi + ctx->cheb_lo) / (ctx->cheb_hi - ctx->cheb_lo);
        const int K = std::max(2, ctx->graph_its - (int)std::floor(std::log(relax) / std::acosh(sigma1)));
        const double rtol2 = (ctx->rtol * relax) * (ctx->rtol * relax);
        // x_1 in xk, x_0 = 0 in xm; iteration k overwrites x_{k-1} with x_{k+1}; x_K must end up in ctx->X
        double* xk = (K % 2 == 1) ? ctx->X : ctx->P;
        double* xm = (xk == ctx->X) ? ctx->P : ctx->X;
        {   // one cooperative kernel for the whole solve when every CTA's row range fits its shared-memory stage
            const DevCsr& A = ctx->As;
            auto k = cheb_solve_coop_kernel<W, solve_gb(W)>;
            // as many CTAs as can be co-resident (narrow panels would otherwise get few, long row ranges)
            const int grid = std::max(1, std::min(max_grid(ctx, k, kSpmmSmemBytes), A.nbr));
            const int rows_cta = (A.nbr + grid - 1) / grid;
            if (ctx->use_coop && K >= 2 && rows_cta <= kSpmmSmemRows && (int64_t)rows_cta * A.max_bpr <= kSpmmSmemBlocks) {
                TimedLaunch tl(ctx, KC_SPMM_CG);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads);
                cfg.dynamicSmemBytes = kSpmmSmemBytes; cfg.stream = ctx->stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeCooperative;
                attr[0].val.cooperative = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, k, (const int*)A.browptr, (const double*)A.blocks, (const double*)ctx->T, xk, xm,
                                   (const double*)ctx->d_Pinv, A.nbr, K, alpha, mu, rtol2, A.diag_first ? 1 : 0, ctx->d_sc, ctx->d_partials, ctx->d_counter, fu);
                ctx->st_op_apps++;
                ctx->solve_path = 2;
                return fu.S != nullptr;
            }
        }
        ctx->solve_path = 1;
        {
            auto k = cheb_solve_init_kernel<W, kCptVec>;
            int grid = std::min(std::min(max_grid(ctx, k), red_grid), blocks_for(ntri, LPB));
            TimedLaunch tl(ctx, KC_CG_INIT);
            launch_k(ctx, k, grid, kThreads, 0, ctx->T, xk, xm, ctx->d_Pinv, ntri, alpha, ctx->d_sc, ctx->d_partials, ctx->d_counter);
        }
        double omega = 1.0;
        const DevCsr& A = ctx->As;
        constexpr int per_cta = PanelCfg<W, kSpmmCpt>::BPC;
        for (int it = 1; it < K; ++it) {
            omega = (it == 1) ? 1.0 / (1.0 - 0.5 * mu * mu) : 1.0 / (1.0 - 0.25 * mu * mu * omega);
            TimedLaunch tl(ctx, KC_SPMM_CG);
            if (it == K - 1) {
                auto k = bsr3_cheb_iter_kernel<W, 1, solve_gb(W)>;
                int grid = std::min(max_grid(ctx, k, kSpmmSmemBytes), std::max(1, (A.nbr + per_cta - 1) / per_cta));
                launch_k(ctx, k, grid, kThreads, kSpmmSmemBytes, A.browptr, A.blocks, xk, xm, ctx->T, ctx->d_Pinv, A.nbr, A.max_bpr,
                         omega, alpha, K, rtol2, ctx->d_sc, ctx->d_partials, ctx->d_counter, fu);
            } else {
                auto k = bsr3_cheb_iter_kernel<W, 0, solve_gb(W)>;
                int grid = std::min(max_grid(ctx, k, kSpmmSmemBytes), std::max(1, (A.nbr + per_cta - 1) / per_cta));
                launch_k(ctx, k, grid, kThreads, kSpmmSmemBytes, A.browptr, A.blocks, xk, xm, ctx->T, ctx->d_Pinv, A.nbr, A.max_bpr,
                         omega, alpha, K, rtol2, ctx->d_sc, ctx->d_partials, ctx->d_counter, FusedUpd{nullptr, nullptr, nullptr, nullptr, 0.0, 0});
            }
            std::swap(xk, xm);
        }
        ctx->st_op_apps++;
        return fu.S != nullptr;
    }
    fu.S = nullptr;                          // PCG: no last iteration known in advance
    ctx->solve_path = 0;
    {
        auto k = cg_init_kernel<W, kCptVec>;
        int grid = std::min(std::min(max_grid(ctx, k), red_grid), blocks_for(ntri, LPB));
        TimedLaunch tl(ctx, KC_CG_INIT);
        launch_k(ctx, k, grid, kThreads, 0, ctx->T, ctx->X, ctx->P, ctx->d_Pinv, ntri, ctx->d_sc, ctx->d_partials, ctx->d_counter);
    }
    for (int it = 0; it < ctx->graph_its; ++it) {
        {
            TimedLaunch tl(ctx, KC_SPMM_CG);
            launch_spmm<W, 1>(ctx, ctx->As, ctx->P, ctx->Q);
        }
        {
            auto k = cg_update_kernel<W, kCptVec>;
            int grid = std::min(std::min(max_grid(ctx, k), red_grid), blocks_for(ntri, LPB));
            TimedLaunch tl(ctx, KC_CG_XR);
            launch_k(ctx, k, grid, kThreads, 0, ctx->X, ctx->T, ctx->P, ctx->Q, ctx->d_Pinv, ntri, ctx->d_sc, ctx->d_partials, ctx->d_counter);
        }
    }
    ctx->st_op_apps++;
    return false;
}

template <int W>
void enqueue_cheb_update(kexp_ctx* ctx, double* S, const double* coef, ChebStep st, bool solved_rows_done = false) {
    constexpr int LPR = W / 2;
    const int ns = (int)ctx->n_s, nb = (int)ctx->n_b, nd = (int)ctx->n_d;
    ChebStep st_b = st, st_d = st;             // back-substitution rows, remaining (depth) rows
    st_b.check_cg = st_d.check_cg = 0;
    if (solved_rows_done) {
        // the solved rows took their update inside the mass solve; the convergence record moves to the next launch
        if (nb > 0) st_b.check_cg = st.check_cg; else st_d.check_cg = st.check_cg;
    } else {
        auto k = cheb_update_kernel<W, 0>;
        int grid = std::min(max_grid(ctx, k), blocks_for(ns, LPR));
        TimedLaunch tl(ctx, KC_CHEB);
        launch_k(ctx, k, grid, kThreads, 0, ctx->X, nullptr, nullptr, nullptr, nullptr, nullptr, S, ctx->Y, ns, coef, ctx->d_tscale, st, ctx->d_sc);
    }
    if (nb > 0) {
        auto k = cheb_update_kernel<W, 1>;
        int grid = std::min(max_grid(ctx, k), blocks_for(nb, LPR));
        const size_t o = (size_t)ns * W;
        TimedLaunch tl(ctx, KC_CHEB);
        launch_k(ctx, k, grid, kThreads, 0, ctx->T + o, ctx->X, ctx->Rt.rowptr, ctx->Rt.col, ctx->Rt.val, nullptr, S + o, ctx->Y + o, nb, coef, ctx->d_tscale, st_b, ctx->d_sc);
    }
    if (nd > 0 && !ctx->depth_rows_fused) {
        const size_t o = (size_t)(ns + nb) * W;
        TimedLaunch tl(ctx, KC_CHEB);
        if (ctx->d_Dblk) {
            auto k = cheb_update_kernel<W, 2>;
            int grid = std::min(max_grid(ctx, k), blocks_for(nd / 3, LPR));
            launch_k(ctx, k, grid, kThreads, 0, ctx->T + o, nullptr, nullptr, nullptr, nullptr, ctx->d_Dblk, S + o, ctx->Y + o, nd, coef, ctx->d_tscale, st_d, ctx->d_sc);
        } else {
            auto k = cheb_update_kernel<W, 0>;
            int grid = std::min(max_grid(ctx, k), blocks_for(nd, LPR));
            launch_k(ctx, k, grid, kThreads, 0, ctx->T + o, nullptr, nullptr, nullptr, nullptr, nullptr, S + o, ctx->Y + o, nd, coef, ctx->d_tscale, st_d, ctx->d_sc);
        }
    }
}

template <int W>
void enqueue_op_finish(kexp_ctx* ctx, double* out) {
    constexpr int LPR = W / 2;
    cudaStream_t s = ctx->stream;
    const int ns = (int)ctx->n_s, nb = (int)ctx->n_b, nd = (int)ctx->n_d;
    {
        TimedLaunch tl(ctx, KC_MISC);
        op_finish_kernel<W, 0><<<std::min(4 * ctx->num_sms, blocks_for(ns, LPR)), kThreads, 0, s>>>(ctx->X, nullptr, nullptr, nullptr, nullptr, nullptr, out, ns);
    }
    if (nb > 0) {
        const size_t o = (size_t)ns * W;
        TimedLaunch tl(ctx, KC_MISC);
        op_finish_kernel<W, 1><<<std::min(4 * ctx->num_sms, blocks_for(nb, LPR)), kThreads, 0, s>>>(ctx->T + o, ctx->X, ctx->Rt.rowptr, ctx->Rt.col, ctx->Rt.val, nullptr, out + o, nb);
    }
    if (nd > 0) {
        const size_t o = (size_t)(ns + nb) * W;
        TimedLaunch tl(ctx, KC_MISC);
        if (ctx->d_Dblk)
            op_finish_kernel<W, 2><<<std::min(4 * ctx->num_sms, blocks_for(nd / 3, LPR)), kThreads, 0, s>>>(ctx->T + o, nullptr, nullptr, nullptr, nullptr, ctx->d_Dblk, out + o, nd);
        else
            op_finish_kernel<W, 0><<<std::min(4 * ctx->num_sms, blocks_for(nd, LPR)), kThreads, 0, s>>>(ctx->T + o, nullptr, nullptr, nullptr, nullptr, nullptr, out + o, nd);
    }
}

// The whole cheby_exp.apply for one panel: S0 holds the input, Y receives the result.
template <int W>
void enqueue_cheby(kexp_ctx* ctx, const PanelPlan& plan) {
    const size_t nelem2 = (size_t)ctx->n * W / 2;
    const double* coef = ctx->d_tscale + kMaxW;             // row m: coefficient of S_m per column
    {
        TimedLaunch tl(ctx, KC_CHEB);
        int grid = (int)std::min<size_t>((size_t)8 * ctx->num_sms, (nelem2 + kThreads - 1) / kThreads);
        launch_k(ctx, cheb_init_kernel<W>, std::max(grid, 1), kThreads, 0, ctx->S0, ctx->Y, nelem2, coef);
    }
    const int bgrid = std::max(1, std::min(4 * ctx->num_sms, (int)((ctx->n + kThreads - 1) / kThreads)));
    if (ctx->capture_basis) {
        TimedLaunch tl(ctx, KC_MISC);
        launch_k(ctx, basis_store_kernel<W>, bgrid, kThreads, 0, ctx->S0, (int)ctx->n, ctx->d_basis, ctx->basis_ld, 0);
    }
    for (int n = 1; n <= plan.M; ++n) {
        const double* in = (n & 1) ? ctx->S0 : ctx->S1;     // S_{n-1}
        double* Sn = (n & 1) ? ctx->S1 : ctx->S0;           // holds S_{n-2}, becomes S_n
        ChebStep st;
        st.n = n; st.has_old = (n >= 2) ? 1 : 0; st.cs = (n == 1) ? 1.0 : 2.0; st.check_cg = 1;
        const bool fused = enqueue_operator<W>(ctx, in, (ctx->use_relax && ctx->solver == 1) ? plan.relax[n] : 1.0,
                                               FusedUpd{Sn, ctx->Y, coef + (size_t)n * kMaxW, ctx->d_tscale, st.cs, st.has_old});
        enqueue_cheb_update<W>(ctx, Sn, coef, st, fused);
        if (ctx->capture_basis) {
            TimedLaunch tl(ctx, KC_MISC);
            launch_k(ctx, basis_store_kernel<W>, bgrid, kThreads, 0, Sn, (int)ctx->n, ctx->d_basis, ctx->basis_ld, n);
        }
    }
}

#define DISPATCH_W(W, CALL)                  \
    switch (W) {                             \
        case 2: { constexpr int kW = 2; CALL; } break;   \
        case 4: { constexpr int kW = 4; CALL; } break;   \
        case 8: { constexpr int kW = 8; CALL; } break;   \
        case 16: { constexpr int kW = 16; CALL; } break; \
        case 32: { constexpr int kW = 32; CALL; } break; \
        case 64: { constexpr int kW = 64; CALL; } break; \
        default: return kexp::fail(ctx, "unsupported panel width"); \
    }

// iterations and squared tolerance of the mass solve of a term whose tolerance is relaxed by `relax`
inline int cheb_iterations(const kexp_ctx* ctx, double relax) {
    const double sigma1 = (ctx->cheb_hi + ctx->cheb_lo) / (ctx->cheb_hi - ctx->cheb_lo);
    return std::max(2, ctx->graph_its - (int)std::floor(std::log(relax) / std::acosh(sigma1)));
}

int run_cheby_panel(kexp_ctx* ctx, int id, int W, const PanelPlan& plan) {
    if (narrow_eligible(ctx, W)) {
        // narrow panels are latency bound: the whole apply is ONE persistent cooperative kernel (kexp_narrow.cu)
        std::vector<int> kterm((size_t)plan.M + 1, 2);
        std::vector<double> r2((size_t)plan.M + 1, 0.0);
        for (int n = 1; n <= plan.M; ++n) {
            const double relax = ctx->use_relax ? plan.relax[n] : 1.0;
            kterm[n] = cheb_iterations(ctx, relax);
            r2[n] = (ctx->rtol * relax) * (ctx->rtol * relax);
        }
        TimedLaunch tl(ctx, KC_SPMM_CG);
        if (launch_narrow_apply(ctx, W, plan.M, 0, kterm.data(), r2.data(), ctx->capture_basis, nullptr)) return 1;
        ctx->st_op_apps += plan.M;
        return 0;
    }
    if (ctx->timing || !ctx->use_graph) {
        DISPATCH_W(W, enqueue_cheby<kW>(ctx, plan));
        KEXP_CUDA(cudaGetLastError());
        return 0;
    }
    const GraphKey key{id + (ctx->capture_basis ? 1000 : 0), W, plan.sig};
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        if (ctx->graphs.size() >= 48) {                  // sample times that change every call: do not hoard graphs
            cudaStreamSynchronize(ctx->stream);
            for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.second);
            ctx->graphs.clear();
        }
        cudaGraph_t graph;
        int64_t l0 = ctx->st_launches, a0 = ctx->st_op_apps;
        KEXP_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        DISPATCH_W(W, enqueue_cheby<kW>(ctx, plan));
        KEXP_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
        ctx->st_launches = l0; ctx->st_op_apps = a0;     // counted per replay below
        cudaGraphExec_t exec;
        KEXP_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = ctx->graphs.emplace(key, exec).first;
    }
    KEXP_CUDA(cudaGraphLaunch(it->second, ctx->stream));
    ctx->st_op_apps += plan.M;
    ctx->st_launches += 1 + (int64_t)plan.M * (2 + (ctx->solver == 1 ? (ctx->use_coop ? 0 : ctx->graph_its - 1) : 2 * ctx->graph_its) + ((ctx->solver == 1 && ctx->use_fused_update) ? 0 : 1) + (ctx->n_b > 0) + ((ctx->n_d > 0 && !ctx->depth_rows_fused) ? 1 : 0))
                        + (ctx->capture_basis ? plan.M + 1 : 0);
    return 0;
}

void invalidate_graphs(kexp_ctx* ctx) {
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.second);
    ctx->graphs.clear();
    LaneState& other = ctx->lane_store[1 - ctx->cur_lane];
    if (other.allocated) {
        if (other.stream) cudaStreamSynchronize(other.stream);
        for (auto& g : other.graphs) cudaGraphExecDestroy(g.second);
        other.graphs.clear();
    }
}

// occupancy queries must not happen during capture: touch every kernel once
template <int W>
void warm_occupancy(kexp_ctx* ctx) {
    max_grid(ctx, csr_spmm_kernel<W>);
    max_grid(ctx, bsr3_spmm_kernel<W, 0, kSpmm0GB>, kSpmmSmemBytes); max_grid(ctx, bsr3_spmm_kernel<W, 1, kSpmmGB>, kSpmmSmemBytes);
    max_grid(ctx, cg_init_kernel<W, kCptVec>); max_grid(ctx, cg_update_kernel<W, kCptVec>);
    max_grid(ctx, cheb_solve_init_kernel<W, kCptVec>);
    max_grid(ctx, cheb_solve_coop_kernel<W, solve_gb(W)>, kSpmmSmemBytes);
    max_grid(ctx, bsr3_cheb_iter_kernel<W, 0, solve_gb(W)>, kSpmmSmemBytes); max_grid(ctx, bsr3_cheb_iter_kernel<W, 1, solve_gb(W)>, kSpmmSmemBytes);
    max_grid(ctx, cheb_update_kernel<W, 0>); max_grid(ctx, cheb_update_kernel<W, 1>); max_grid(ctx, cheb_update_kernel<W, 2>);
}

// Upload one panel's operator scales and series coefficients (graphs read them from d_tscale).
int set_plan(kexp_ctx* ctx, const PanelPlan& plan) {
    cudaStreamSynchronize(ctx->stream);        // h_pin reuse
    memcpy(ctx->h_pin, plan.tscale, sizeof(double) * kMaxW);
    memcpy(ctx->h_pin + kMaxW, plan.coef.data(), sizeof(double) * plan.coef.size());
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_tscale, ctx->h_pin, sizeof(double) * (kMaxW + plan.coef.size()), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int check_ready(kexp_ctx* ctx, int id) {
    if (!ctx) return 1;
    if (!ctx->have_op) return kexp::fail(ctx, "operator not set (kexp_set_operator)");
    if (id < 0 || id >= KEXP_MAX_CHEBY || ctx->cheb[id].ncheb < 0) return kexp::fail(ctx, "Chebyshev table not set (kexp_set_cheby)");
    return 0;
}

template <int W>
void launch_pack(kexp_ctx* ctx, const double* d_src, int64_t ld, int shared_x, int B, int col0, double* dst) {
    dim3 blk(32, 8);
    int grid = (int)((ctx->n + 31) / 32);
    pack_kernel<W><<<grid, blk, 0, ctx->stream>>>(d_src, ld, shared_x, B, col0, ctx->d_perm, dst, (int)ctx->n);
}

template <int W>
void launch_unpack(kexp_ctx* ctx, const double* src, double* d_dst, int64_t ld, int B, int col0) {
    dim3 blk(32, 8);
    int grid = (int)((ctx->n + 31) / 32);
    unpack_kernel<W><<<grid, blk, 0, ctx->stream>>>(src, ctx->d_perm, (int)ctx->n, d_dst, ld, B, col0);
}

template <int W>
void launch_colsum(kexp_ctx* ctx, const double* Ypanel, double* v_dev, int accumulate) {
    int grid = std::min(8 * ctx->num_sms, blocks_for(ctx->n, W / 2));
    weighted_colsum_kernel<W><<<grid, kThreads, 0, ctx->stream>>>(Ypanel, ctx->d_w, ctx->d_perm, (int)ctx->n, v_dev, accumulate);
}

template <int W>
void launch_spmm_K(kexp_ctx* ctx, const double* in, double* out) {
    launch_spmm<W, 0>(ctx, ctx->Kt, in, out);
}

template <int W>
void launch_operator_full(kexp_ctx* ctx, const double* in, double* out) {
    enqueue_operator<W>(ctx, in);
    enqueue_op_finish<W>(ctx, out);
}

int pack_panel(kexp_ctx* ctx, const double* d_src, int64_t ld, int shared_x, int B, int col0, int W, double* dst) {
    TimedLaunch tl(ctx, KC_MISC);
    DISPATCH_W(W, launch_pack<kW>(ctx, d_src, ld, shared_x, B, col0, dst));
    return 0;
}

int unpack_panel(kexp_ctx* ctx, const double* src, double* d_dst, int64_t ld, int B, int col0, int W) {
    TimedLaunch tl(ctx, KC_MISC);
    DISPATCH_W(W, launch_unpack<kW>(ctx, src, d_dst, ld, B, col0));
    return 0;
}

int colsum_panel(kexp_ctx* ctx, const double* Ypanel, const double* w, int col0, int B, int W, double* v_dev, int accumulate) {
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < kMaxW; ++k) ctx->h_pin[kPinWeights + k] = (k < W && col0 + k < B) ? w[col0 + k] : 0.0;
    KEXP_CUDA(cudaMemcpyAsync(ctx->d_w, ctx->h_pin + kPinWeights, sizeof(double) * kMaxW, cudaMemcpyHostToDevice, ctx->stream));
    TimedLaunch tl(ctx, KC_MISC);
    DISPATCH_W(W, launch_colsum<kW>(ctx, Ypanel, v_dev, accumulate));
    return 0;
}

}  // namespace

namespace {
template <int W>
void launch_combine(kexp_ctx* ctx, int nb, const double* D, double* Y) {
    auto k = basis_combine_kernel<W>;
    const size_t smem_max = sizeof(double) * (size_t)std::min(nb, kCombineChunk) * W;
    if (smem_max > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    const int grid = std::min(8 * ctx->num_sms, blocks_for(ctx->n, W / 2));
    for (int m0 = 0; m0 < nb; m0 += kCombineChunk) {          // long series: one pass per chunk of terms
        const int cnt = std::min(kCombineChunk, nb - m0);
        k<<<grid, kThreads, sizeof(double) * (size_t)cnt * W, ctx->stream>>>(ctx->d_basis + m0, ctx->basis_ld, cnt, D + (size_t)m0 * W, Y,
                                                                             (int)ctx->n, m0 > 0 ? 1 : 0);
    }
}

template <int W>
void enqueue_clenshaw(kexp_ctx* ctx, const PanelPlan& plan) {
    const int n = (int)ctx->n;
    const size_t bytes = sizeof(double) * (size_t)n * W;
    const int bgrid = std::max(1, std::min(4 * ctx->num_sms, (n + kThreads - 1) / kThreads));
    const double* coef = ctx->d_tscale + kMaxW;             // one row of zeros: the Y accumulation is not used
    cudaMemsetAsync(ctx->S0, 0, bytes, ctx->stream);
    cudaMemsetAsync(ctx->S1, 0, bytes, ctx->stream);
    cudaMemsetAsync(ctx->Y, 0, bytes, ctx->stream);
    auto add_z = [&](double* S, int m) {
        TimedLaunch tl(ctx, KC_MISC);
        basis_axpy_kernel<W><<<bgrid, kThreads, 0, ctx->stream>>>(S, n, ctx->d_basis, ctx->basis_ld, m);
    };
    double* A = ctx->S0;                                    // u_{m+1}
    double* Bf = ctx->S1;                                   // u_{m+2}
    add_z(A, plan.M);                                       // u_M = z_M
    for (int m = plan.M - 1; m >= 0; --m) {
        add_z(Bf, m);                                       // u_{m+2} + z_m
        ChebStep st;
        st.n = 0; st.has_old = 1; st.cs = (m == 0) ? 1.0 : 2.0; st.check_cg = 1;
        const bool fused = enqueue_operator<W>(ctx, A, (ctx->use_relax && ctx->solver == 1) ? plan.relax[m] : 1.0,
                                               FusedUpd{Bf, ctx->Y, coef, ctx->d_tscale, st.cs, 1});
        enqueue_cheb_update<W>(ctx, Bf, coef, st, fused);   // + cs * c * M^-1 K u_{m+1}
        std::swap(A, Bf);
    }
}

template <int W>
void launch_clenshaw_combine(kexp_ctx* ctx, const double* X, const double* G, int nb, int accumulate) {
    auto k = clenshaw_combine_kernel<W>;
    const size_t smem_max = sizeof(double) * (size_t)std::min(nb, kCombineChunk) * W;
    if (smem_max > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    const int grid = std::max(1, std::min(8 * ctx->num_sms, (int)((ctx->n + kThreads - 1) / kThreads)));
    for (int m0 = 0; m0 < nb; m0 += kCombineChunk) {          // long series: one pass over the inputs per chunk of terms
        const int cnt = std::min(kCombineChunk, nb - m0);
        k<<<grid, kThreads, sizeof(double) * (size_t)cnt * W, ctx->stream>>>(X, (int)ctx->n, G + (size_t)m0 * W, cnt, ctx->d_basis + m0,
                                                                             ctx->basis_ld, accumulate);
    }
}

int alloc_lane_scalars(kexp_ctx* ctx) {
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_sc, sizeof(CgScalars)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_sc, 0, sizeof(CgScalars)) == cudaSuccess;
    // block partials (3 accumulators x up to 8 blocks/SM) + group partials, and the tickets
    ok = ok && cudaMalloc(&ctx->d_partials, sizeof(double) * 3 * 9 * (size_t)ctx->num_sms * kMaxW + 4096) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_counter, sizeof(unsigned) * (size_t)(2 + ctx->num_sms)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_counter, 0, sizeof(unsigned) * (size_t)(2 + ctx->num_sms)) == cudaSuccess;
    // [kMaxW operator scales][kMaxTerms rows of kMaxW series coefficients]; the pinned mirror has the
    // averaging weights behind it
    ok = ok && cudaMalloc(&ctx->d_tscale, sizeof(double) * kPinWeights) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_pin, sizeof(double) * (kPinWeights + kMaxW)) == cudaSuccess;
    const double r2 = ctx->rtol * ctx->rtol;
    ok = ok && cudaMemcpy(&ctx->d_sc->rtol2, &r2, sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    ctx->curW = 0;
    if (!ok) return kexp::fail(ctx, "out of memory allocating lane state");
    return 0;
}

template <typename F>
int for_each_lane(kexp_ctx* ctx, F f) {     // run f with every allocated lane active, end on lane 0
    int rc = 0;
    for (int i = 1; i >= 0; --i) {
        if (!ctx->lane_store[i].allocated && i != ctx->cur_lane) continue;
        if (swap_lane(ctx, i)) return 1;
        rc |= f();
    }
    return rc;
}
}  // namespace

// ====================================================================== C ABI
extern "C" {

static int poll_deferred(kexp_ctx* ctx, bool block);      // deferred convergence check of asynchronous applies (below)
static bool panels_unset(const kexp_batch* b);
static int unconverged_error(kexp_ctx* ctx, double worst, bool deferred);
static void raise_iterations(kexp_ctx* ctx);

const char* kexp_version(void) { return "kexp-b200 0.2 (sm_100a)"; }

const char* kexp_last_error(kexp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int kexp_create(kexp_ctx** out, int device) {
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "kexp_create: no CUDA device (this library has no CPU path)\n");
        return 2;
    }
    if (device < 0 || device >= ndev) return 3;
    if (cudaSetDevice(device) != cudaSuccess) return 4;
    kexp_ctx* ctx = new kexp_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        fprintf(stderr, "kexp_create: device %d is sm_%d%d; this build targets sm_100a (B200)\n", device, prop.major, prop.minor);
        delete ctx;
        return 5;
    }
    {
        static int next_slot = 0;                   // constant-memory argument slots of the persistent kernel (kexp_narrow.cu)
        ctx->narrow_slot = (next_slot++) % 8;
    }
    ctx->lane_store[0].allocated = true;
    bool ok = alloc_lane_scalars(ctx) == 0;
    ok = ok && cudaMalloc(&ctx->d_w, sizeof(double) * kMaxW) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev_a) == cudaSuccess && cudaEventCreate(&ctx->ev_b) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_chk, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost(&ctx->h_sc, sizeof(CgScalars)) == cudaSuccess;
    if (ok) memset(ctx->h_sc, 0, sizeof(CgScalars));
    if (!ok) { kexp_destroy(ctx); return 6; }
    const char* g = getenv("KEXP_NO_GRAPH");
    if (g && g[0] == '1') ctx->use_graph = false;
    const char* ln = getenv("KEXP_LANES");
    if (ln && ln[0] == '1') { ctx->nlanes = 1; ctx->panel_cap = 64; }
    if (ln && ln[0] == '2') { ctx->nlanes = 2; ctx->panel_cap = 32; }
    if (const char* fz = getenv("KEXP_FUSED_UPDATE")) ctx->use_fused_update = (fz[0] != '0');
    if (const char* nw = getenv("KEXP_NARROW")) ctx->use_narrow = (nw[0] != '0');
    if (const char* nr = getenv("KEXP_NARROW_MREG")) ctx->narrow_mreg = (nr[0] != '0');
    if (const char* nm = getenv("KEXP_NARROW_MAXW")) { const int v = atoi(nm); if (v >= 2 && v <= 8) ctx->narrow_max_w = v; }
    const char* rl = getenv("KEXP_RELAX");
    if (rl) ctx->use_relax = (rl[0] == '1');
    const char* co = getenv("KEXP_COOP");
    if (co) ctx->use_coop = (co[0] == '1');
    if (const char* ec = getenv("KEXP_ECONOMISE")) ctx->use_econ = (ec[0] != '0');
    if (const char* ee = getenv("KEXP_ECON_EPS")) { const double v = atof(ee); if (v > 0.0 && v < 1e-6) ctx->econ_eps = v; }
    const char* shb = getenv("KEXP_SHARED_BASIS");
    if (shb && shb[0] == '0') ctx->use_shared_basis = false;
    const char* sb = getenv("KEXP_SPMM_BLOCKS");
    if (sb && atoi(sb) >= 1) ctx->spmm_blocks_per_sm = atoi(sb);
    const char* pd = getenv("KEXP_PDL");
    if (pd && pd[0] == '0') ctx->use_pdl = false;
    const char* rb = getenv("KEXP_RED_BLOCKS");
    if (rb && atoi(rb) >= 1 && atoi(rb) <= 8) ctx->red_blocks_per_sm = atoi(rb);
    const char* ad = getenv("KEXP_ADAPTIVE_ITS");
    if (ad && ad[0] == '0') ctx->adaptive_its = false;
    if (kexp_set_mass_solver(ctx, 1e-12, 60, 3)) { kexp_destroy(ctx); return 7; }
    *out = ctx;
    return 0;
}

int kexp_destroy(kexp_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    swap_lane(ctx, 0);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    invalidate_graphs(ctx);
    if (ctx->tmp_in) { kexp_batch_destroy(ctx->tmp_in); kexp_batch_destroy(ctx->tmp_out); ctx->tmp_in = ctx->tmp_out = nullptr; }
    kexp::krylov_free(ctx);
    {
        LaneState& L = ctx->lane_store[1];
        if (L.allocated) {
            double* lb[] = {L.S0, L.S1, L.Y, L.T, L.X, L.P, L.Q, L.d_partials, L.d_tscale};
            for (double* b : lb) cudaFree(b);
            cudaFree(L.d_sc); cudaFree(L.d_counter);
            if (L.h_pin) cudaFreeHost(L.h_pin);
            if (L.stream) cudaStreamDestroy(L.stream);
        }
    }
    free_csr(ctx->Kt); free_csr(ctx->As); free_csr(ctx->Rt);
    double* bufs[] = {ctx->S0, ctx->S1, ctx->Y, ctx->T, ctx->X, ctx->P, ctx->Q, ctx->d_Dblk, ctx->d_Pinv,
                      ctx->d_partials, ctx->d_tscale, ctx->d_w, ctx->d_stage, ctx->d_stage_up, ctx->d_v, ctx->d_basis, ctx->d_D};
    for (double* b : bufs) cudaFree(b);
    cudaFree(ctx->d_perm); cudaFree(ctx->d_sc); cudaFree(ctx->d_counter);
    cudaFree(ctx->d_nk); cudaFree(ctx->d_nrtol2); cudaFree(ctx->d_npartials);
    cudaFree(ctx->d_nktcols);
    kexp::free_narrow_plan(ctx->narrow_plan);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->ev_prev) cudaEventDestroy(ctx->ev_prev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    if (ctx->ev_unpacked) cudaEventDestroy(ctx->ev_unpacked);
    if (ctx->ev_d2h_done) cudaEventDestroy(ctx->ev_d2h_done);
    cudaFree(ctx->d_stage_out);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_chk) cudaEventDestroy(ctx->ev_chk);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->nccl_comm && ctx->nccl_lib) {
        typedef int (*destroy_t)(void*);
        destroy_t f = (destroy_t)dlsym(ctx->nccl_lib, "ncclCommDestroy");
        if (f) f(ctx->nccl_comm);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

static int build_block_jacobi(kexp_ctx* ctx, int64_t n_s, const int32_t* rp, const int32_t* ci, const double* v,
                              int precond, std::vector<double>& Pinv) {
    if (n_s % 3 != 0) return kexp::fail(ctx, "mass system size must be a multiple of 3 (BDM2 entity blocks)");
    const int64_t nb = n_s / 3;
    Pinv.assign((size_t)nb * 9, 0.0);
    for (int64_t b = 0; b < nb; ++b) {
        double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < 3; ++i) {
            const int64_t r = 3 * b + i;
            for (int k = rp[r]; k < rp[r + 1]; ++k) {
                const int64_t c = ci[k];
                if (c >= 3 * b && c < 3 * b + 3 && (precond == 3 || c == r)) A[i][c - 3 * b] += v[k];
            }
        }
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        if (!(std::fabs(det) > 0.0)) return kexp::fail(ctx, "singular diagonal block in the mass matrix");
        double* P = &Pinv[(size_t)b * 9];
        P[0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det;
        P[1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det;
        P[2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
        P[3] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det;
        P[4] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det;
        P[5] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
        P[6] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det;
        P[7] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det;
        P[8] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
    }
    return 0;
}

int kexp_set_operator_plan(kexp_ctx* ctx, int64_t n, int64_t n_s, int64_t n_b,
                           const int32_t* Kt_rowptr, const int32_t* Kt_col, const double* Kt_val,
                           const int32_t* As_rowptr, const int32_t* As_col, const double* As_val,
                           const int32_t* Rt_rowptr, const int32_t* Rt_col, const double* Rt_val,
                           const double* Dblk, const double* Pinv_blocks, const int32_t* perm) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (n <= 0 || n_s <= 0 || n_b < 0 || n_s + n_b > n) return kexp::fail(ctx, "bad operator plan sizes");
    if (n_s % 3) return kexp::fail(ctx, "n_s must be a multiple of 3");
    if (n >= (int64_t)1 << 31) return kexp::fail(ctx, "n too large for 32-bit row indices");
    if (!Kt_rowptr || !As_rowptr) return kexp::fail(ctx, "null matrix");
    const int64_t n_d = n - n_s - n_b;
    if (Dblk && n_d % 3) return kexp::fail(ctx, "block rows must be a multiple of 3");
    if (n_b > 0 && !Rt_rowptr) return kexp::fail(ctx, "back-substitution matrix missing");
    invalidate_graphs(ctx);
    if (ctx->tmp_in) { kexp_batch_destroy(ctx->tmp_in); kexp_batch_destroy(ctx->tmp_out); ctx->tmp_in = ctx->tmp_out = nullptr; }
    ctx->have_op = false;
    kexp::free_narrow_plan(ctx->narrow_plan);      // the per-CTA local problems of S belong to the old operator
    cudaFree(ctx->d_nktcols); ctx->d_nktcols = nullptr;
    ctx->n = n; ctx->n_s = n_s; ctx->n_b = n_b; ctx->n_d = n_d;
    ctx->curW = 0;   // panels are re-allocated for the new size
    ctx->lane_store[1 - ctx->cur_lane].curW = 0;
    if (upload_csr(ctx, ctx->Kt, (int)n, (int)n, Kt_rowptr, Kt_col, Kt_val, false)) return 1;
    if (upload_csr(ctx, ctx->As, (int)n_s, (int)n_s, As_rowptr, As_col, As_val, true)) return 1;
    if (!ctx->As.bsr()) return kexp::fail(ctx, "mass matrix could not be stored as 3x3 blocks");
    if (upload_csr(ctx, ctx->Rt, (int)n_b, (int)n_s, n_b ? Rt_rowptr : nullptr, Rt_col, Rt_val, false)) return 1;
    ctx->h_As_rowptr.assign(As_rowptr, As_rowptr + n_s + 1);
    ctx->h_As_col.assign(As_col, As_col + As_rowptr[n_s]);
    ctx->h_As_val.assign(As_val, As_val + As_rowptr[n_s]);
    cudaFree(ctx->d_Dblk); ctx->d_Dblk = nullptr;
    if (Dblk && n_d > 0) {
        KEXP_CUDA(cudaMalloc(&ctx->d_Dblk, sizeof(double) * 3 * (size_t)n_d));
        KEXP_CUDA(cudaMemcpy(ctx->d_Dblk, Dblk, sizeof(double) * 3 * (size_t)n_d, cudaMemcpyHostToDevice));
    }
    std::vector<double> Pinv;
    const double* P = Pinv_blocks;
    if (!P) {
        if (build_block_jacobi(ctx, n_s, As_rowptr, As_col, As_val, ctx->precond, Pinv)) return 1;
        P = Pinv.data();
    }
    cudaFree(ctx->d_Pinv); ctx->d_Pinv = nullptr;
    KEXP_CUDA(cudaMalloc(&ctx->d_Pinv, sizeof(double) * 3 * (size_t)n_s + 64));   // + slack: bulk copies round up to 16 B
    KEXP_CUDA(cudaMemcpy(ctx->d_Pinv, P, sizeof(double) * 3 * (size_t)n_s, cudaMemcpyHostToDevice));
    cudaFree(ctx->d_perm); ctx->d_perm = nullptr;
    if (perm) {
        std::vector<char> seen((size_t)n, 0);
        for (int64_t i = 0; i < n; ++i) {
            if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]) return kexp::fail(ctx, "perm is not a permutation");
            seen[perm[i]] = 1;
        }
        KEXP_CUDA(cudaMalloc(&ctx->d_perm, sizeof(int) * (size_t)n));
        KEXP_CUDA(cudaMemcpy(ctx->d_perm, perm, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    }
    cudaFree(ctx->d_v); ctx->d_v = nullptr;
    KEXP_CUDA(cudaMalloc(&ctx->d_v, sizeof(double) * (size_t)n));
    {   // panels of 32 columns when that makes the mass-solve working set (3 panels of n_s rows) L2 resident
        int l2 = 0;
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device);
        const double ws32 = 3.0 * 8.0 * 32.0 * (double)n_s, ws64 = 2.0 * ws32;
        if (!getenv("KEXP_LANES") && !getenv("KEXP_PANEL_CAP")) ctx->panel_cap = (ws64 > 0.75 * l2 && ws32 <= 0.75 * l2) ? 32 : 64;
        const char* pc = getenv("KEXP_PANEL_CAP");
        if (pc && (atoi(pc) == 32 || atoi(pc) == 64 || atoi(pc) == 16)) ctx->panel_cap = atoi(pc);
    }
    ctx->have_op = true;
    return 0;
}

int kexp_set_operator(kexp_ctx* ctx, int64_t n_u, int64_t n_eta,
                      const int32_t* K_rowptr, const int32_t* K_col, const double* K_val,
                      const int32_t* Mu_rowptr, const int32_t* Mu_col, const double* Mu_val,
                      const double* Meta_inv_blocks, int64_t n_cells) {
    if (!ctx) return 1;
    if (n_eta != 3 * n_cells) return kexp::fail(ctx, "n_eta must equal 3*n_cells (DG1)");
    if (!Meta_inv_blocks) return kexp::fail(ctx, "Meta_inv_blocks is null");
    return kexp_set_operator_plan(ctx, n_u + n_eta, n_u, 0, K_rowptr, K_col, K_val, Mu_rowptr, Mu_col, Mu_val,
                                  nullptr, nullptr, nullptr, Meta_inv_blocks, nullptr, nullptr);
}

int kexp_set_mass_solver(kexp_ctx* ctx, double rtol, int max_it, int precond) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!(rtol > 0.0) || max_it < 1 || (precond != 1 && precond != 3)) return kexp::fail(ctx, "bad mass solver parameters");
    rtol = std::max(rtol, 2e-14);          // below the attainable accuracy of the fp64 residual (include/kexp.h)
    invalidate_graphs(ctx);
    const bool rebuild = ctx->have_op && precond != ctx->precond;
    ctx->rtol = rtol; ctx->max_it = max_it; ctx->precond = precond;
    ctx->graph_its = max_it; ctx->its_tuned = false; ctx->solver = 0;
    if (rebuild) {
        std::vector<double> Pinv;
        if (build_block_jacobi(ctx, ctx->n_s, ctx->h_As_rowptr.data(), ctx->h_As_col.data(), ctx->h_As_val.data(), precond, Pinv)) return 1;
        KEXP_CUDA(cudaMemcpy(ctx->d_Pinv, Pinv.data(), sizeof(double) * Pinv.size(), cudaMemcpyHostToDevice));
    }
    const double r2 = rtol * rtol;
    KEXP_CUDA(cudaMemcpy(&ctx->d_sc->rtol2, &r2, sizeof(double), cudaMemcpyHostToDevice));
    LaneState& other = ctx->lane_store[1 - ctx->cur_lane];
    if (other.allocated) KEXP_CUDA(cudaMemcpy(&other.d_sc->rtol2, &r2, sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

// Chebyshev semi-iteration for the mass solve.  [lmin, lmax] must enclose the spectrum of P^-1 S
// (P the 3x3 block-Jacobi); the iteration count follows from 1/T_k(sigma) <= rtol and is raised
// (up to max_it) if the measured residual of a solve misses rtol.
int kexp_set_mass_solver_chebyshev(kexp_ctx* ctx, double rtol, double lmin, double lmax, int max_it) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!(rtol > 0.0) || !(lmin > 0.0) || !(lmax > lmin) || max_it < 2) return kexp::fail(ctx, "bad Chebyshev mass solver parameters");
    if (kexp_set_mass_solver(ctx, rtol, max_it, 3)) return 1;
    const double sigma1 = (lmax + lmin) / (lmax - lmin);
    int K = (int)std::ceil(std::acosh(1.0 / rtol) / std::acosh(sigma1)) + 1;     // the measured residual is the arbiter (adaptive +2)
    K = std::max(2, std::min(K, max_it));
    ctx->solver = 1; ctx->cheb_lo = lmin; ctx->cheb_hi = lmax;
    ctx->graph_its = K; ctx->its_tuned = true;      // fixed length; only ever grows
    return 0;
}

// Tolerance relaxation per term: an error introduced in the mass solve of term n reaches y only through
// sum_{m>=n} a_m U_{m-n}(B) e with ||U_k|| <= k+1 on the spectrum, i.e. with weight w_n = sum_{m>=n} |a_m| (m-n+1).
// The a_m decay super-exponentially beyond m ~ L, so late terms tolerate rtol * w_max / w_n (capped).
static void relax_factors(const double* a, int M, std::vector<double>& relax) {
    relax.assign((size_t)M + 1, 1.0);
    std::vector<double> wgt((size_t)M + 1, 0.0);
    double wmax = 0.0;
    for (int n = 1; n <= M; ++n) {
        double w = 0.0;
        for (int m = n; m <= M; ++m) w += std::fabs(a[m]) * (double)(m - n + 1);
        wgt[n] = w; wmax = std::max(wmax, w);
    }
    for (int n = 1; n <= M; ++n)
        relax[n] = (wgt[n] > 0.0) ? std::min(1e6, std::max(1.0, wmax / wgt[n])) : 1e6;
}

int kexp_set_cheby_time_range(kexp_ctx* ctx, int id, double t_valid) {
    if (check_ready(ctx, id)) return 1;
    if (!(t_valid >= 0.0)) return kexp::fail(ctx, "t_valid must be >= 0 (0 disables the economised series)");
    ctx->cheb[id].t_valid = t_valid;
    ctx->plan_cache.clear();
    return 0;
}

int kexp_set_cheby(kexp_ctx* ctx, int id, const double* a, int ncheb, double L) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (id < 0 || id >= KEXP_MAX_CHEBY) return kexp::fail(ctx, "cheby id out of range");
    if (!a || ncheb < 0 || !(L > 0.0)) return kexp::fail(ctx, "bad Chebyshev table");
    if (ncheb >= kMaxTerms) return kexp::fail(ctx, "Chebyshev series too long (ncheb must be below 1024)");
    invalidate_graphs(ctx);
    ChebTable& tb = ctx->cheb[id];
    tb.ncheb = ncheb; tb.L = L; tb.t_valid = 0.0;
    ctx->plan_cache.clear();
    tb.h_coef.assign(a, a + ncheb + 1);
    // Tolerance relaxation per term: an error introduced in the mass solve of term n reaches y only through
    // sum_{m>=n} a_m U_{m-n}(B) e with ||U_k|| <= k+1 on the spectrum, i.e. with weight w_n = sum_{m>=n} |a_m| (m-n+1).
    // The a_m decay super-exponentially beyond m ~ L, so late terms tolerate rtol * w_max / w_n (capped).
    relax_factors(tb.h_coef.data(), ncheb, tb.h_relax);
    return 0;
}

// ------------------------------------------------------------------ batches
int kexp_batch_create(kexp_ctx* ctx, int B, kexp_batch** out) {
    if (!ctx || !out) return 1;
    cudaSetDevice(ctx->device);
    if (!ctx->have_op) return kexp::fail(ctx, "operator not set");
    if (B < 1) return kexp::fail(ctx, "B must be >= 1");
    kexp_batch* b = new kexp_batch();
    b->ctx = ctx; b->B = B;
    const int cap = (B > ctx->panel_cap) ? ctx->panel_cap : kMaxW;   // split only what does not fit one panel of the cap
    b->stride = cap;
    for (int c0 = 0; c0 < B; c0 += cap) {
        const int W = pick_W(std::min(B - c0, cap));
        double* p = nullptr;
        if (cudaMalloc(&p, sizeof(double) * (size_t)ctx->n * W) != cudaSuccess) {
            kexp_batch_destroy(b);
            return kexp::fail(ctx, "out of device memory for batch");
        }
        b->panelW.push_back(W); b->panel.push_back(p);
    }
    *out = b;
    return 0;
}

int kexp_batch_num_panels(kexp_batch* b) { return b ? (int)b->panel.size() : 0; }

int kexp_batch_destroy(kexp_batch* b) {
    if (!b) return 0;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->pending && b->ctx->copy_stream) cudaStreamSynchronize(b->ctx->copy_stream);
    for (double* p : b->panel) cudaFree(p);
    for (cudaEvent_t e : b->ready) cudaEventDestroy(e);
    if (b->ev_consumed) cudaEventDestroy(b->ev_consumed);
    cudaFree(b->shared_vec);
    delete b;
    return 0;
}

// Upload of a host batch.  Distinct-input batches are copied and packed panel by panel on a separate copy
// stream, each panel signalling an event; with wait == false the caller (the host-pointer entry points) starts
// computing on panel 0 while panel 1 is still crossing PCIe.
static int finish_upload(kexp_batch* b) {
    kexp_ctx* ctx = b->ctx;
    if (b->pending) {
        KEXP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        b->pending = false;
    }
    return 0;
}

static int batch_upload_impl(kexp_batch* b, const double* X, int64_t ldx, int x_is_shared, bool wait) {
    if (!b || !X) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (swap_lane(ctx, 0)) return 1;
    const int nvec = x_is_shared ? 1 : b->B;
    if (!x_is_shared && ldx < ctx->n) return kexp::fail(ctx, "ldx < n");
    ctx->st_h2d_bytes += (int64_t)sizeof(double) * ctx->n * nvec;
    if (x_is_shared ? ensure_stage(ctx, sizeof(double) * (size_t)ctx->n * nvec) : ensure_stage_up(ctx, sizeof(double) * (size_t)ctx->n * nvec)) return 1;
    b->is_shared = x_is_shared != 0;
    if (finish_upload(b)) return 1;
    if (b->is_shared) {
        KEXP_CUDA(cudaMemcpy2DAsync(ctx->d_stage, sizeof(double) * ctx->n, X, sizeof(double) * ctx->n, sizeof(double) * ctx->n, 1,
                                    cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->use_shared_basis) {
            // one vector only: the forward pass runs the recurrence for a single column (see basis_combine_kernel)
            if (!b->shared_vec) KEXP_CUDA(cudaMalloc(&b->shared_vec, sizeof(double) * 2 * (size_t)ctx->n));
            if (pack_panel(ctx, ctx->d_stage, ctx->n, 1, 1, 0, 2, b->shared_vec)) return 1;
        } else {
            for (size_t p = 0; p < b->panel.size(); ++p)
                if (pack_panel(ctx, ctx->d_stage, ctx->n, 1, b->B, (int)p * b->stride, b->panelW[p], b->panel[p])) return 1;
        }
        KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (!ctx->copy_stream) {
        KEXP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        KEXP_CUDA(cudaEventCreateWithFlags(&ctx->ev_prev, cudaEventDisableTiming));
    }
    while (b->ready.size() < b->panel.size()) {
        cudaEvent_t e;
        KEXP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->ready.push_back(e);
    }
    // the copy stream starts after whatever the compute stream still does with the panels: for the scratch batch of the
    // host-pointer entry points that is its last apply (mark_consumed), for any other batch everything queued so far
    if (b == ctx->tmp_in && b->ev_consumed) {
        KEXP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, b->ev_consumed, 0));
    } else {
        KEXP_CUDA(cudaEventRecord(ctx->ev_prev, ctx->stream));
        KEXP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_prev, 0));
    }
    cudaStream_t compute = ctx->stream;
    ctx->stream = ctx->copy_stream;                  // pack kernels of this upload run behind their copy
    int rc = 0;
    for (size_t p = 0; p < b->panel.size() && !rc; ++p) {
        const int col0 = (int)p * b->stride;
        const int ncols = std::min(b->panelW[p], b->B - col0);
        cudaError_t e = cudaSuccess;
        if (b->order.empty())
            e = cudaMemcpy2DAsync(ctx->d_stage_up + (size_t)col0 * ctx->n, sizeof(double) * ctx->n, X + (size_t)col0 * ldx,
                                  sizeof(double) * ldx, sizeof(double) * ctx->n, ncols, cudaMemcpyHostToDevice, ctx->copy_stream);
        else
            for (int c = 0; c < ncols && e == cudaSuccess; ++c)
                e = cudaMemcpyAsync(ctx->d_stage_up + (size_t)(col0 + c) * ctx->n, X + (size_t)b->order[col0 + c] * ldx,
                                    sizeof(double) * ctx->n, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e != cudaSuccess) { rc = kexp::fail(ctx, cudaGetErrorString(e)); break; }
        rc = pack_panel(ctx, ctx->d_stage_up, ctx->n, 0, b->B, col0, b->panelW[p], b->panel[p]);
        if (!rc && cudaEventRecord(b->ready[p], ctx->copy_stream) != cudaSuccess) rc = kexp::fail(ctx, "event record failed");
    }
    ctx->stream = compute;
    if (rc) return rc;
    b->pending = true;
    if (wait) return finish_upload(b);
    return 0;
}

int kexp_batch_upload(kexp_batch* b, const double* X, int64_t ldx, int x_is_shared) {
    return batch_upload_impl(b, X, ldx, x_is_shared, true);
}

int kexp_batch_download(kexp_batch* b, double* Y, int64_t ldy) {
    if (!b || !Y) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (ldy < ctx->n) return kexp::fail(ctx, "ldy < n");
    if (b->is_shared && ctx->use_shared_basis) return kexp::fail(ctx, "a shared-input batch holds one vector; nothing to download");
    if (finish_upload(b)) return 1;
    if (ensure_stage(ctx, sizeof(double) * (size_t)ctx->n * b->B)) return 1;
    ctx->st_d2h_bytes += (int64_t)sizeof(double) * ctx->n * b->B;
    for (size_t p = 0; p < b->panel.size(); ++p)
        if (unpack_panel(ctx, b->panel[p], ctx->d_stage, ctx->n, b->B, (int)p * b->stride, b->panelW[p])) return 1;
    if (b->order.empty()) {
        KEXP_CUDA(cudaMemcpy2DAsync(Y, sizeof(double) * ldy, ctx->d_stage, sizeof(double) * ctx->n, sizeof(double) * ctx->n, b->B,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        for (int j = 0; j < b->B; ++j)
            KEXP_CUDA(cudaMemcpyAsync(Y + (size_t)b->order[j] * ldy, ctx->d_stage + (size_t)j * ctx->n, sizeof(double) * ctx->n,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    }
    KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
    return poll_deferred(ctx, true);
}

// Device-resident vectors in the caller's layout (B vectors of n doubles in user dof order, vector k at d + k*ld; e.g. a
// torch tensor): device -> panels and panels -> device, enqueued on the library's stream, no host copy, no synchronisation.
int kexp_batch_upload_dev(kexp_batch* b, const double* d_X, int64_t ldx, int x_is_shared) {
    if (!b || !d_X) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (swap_lane(ctx, 0)) return 1;
    if (!x_is_shared && ldx < ctx->n) return kexp::fail(ctx, "ldx < n");
    if (finish_upload(b)) return 1;
    b->is_shared = x_is_shared != 0;
    if (b->is_shared) {
        if (ctx->use_shared_basis) {
            if (!b->shared_vec) KEXP_CUDA(cudaMalloc(&b->shared_vec, sizeof(double) * 2 * (size_t)ctx->n));
            return pack_panel(ctx, d_X, ctx->n, 1, 1, 0, 2, b->shared_vec);
        }
        for (size_t p = 0; p < b->panel.size(); ++p)
            if (pack_panel(ctx, d_X, ctx->n, 1, b->B, (int)p * b->stride, b->panelW[p], b->panel[p])) return 1;
        return 0;
    }
    const double* src = d_X;
    int64_t ld = ldx;
    if (!b->order.empty()) {                       // stored column j holds the caller's column order[j]
        if (ensure_stage(ctx, sizeof(double) * (size_t)ctx->n * b->B)) return 1;
        for (int j = 0; j < b->B; ++j)
            KEXP_CUDA(cudaMemcpyAsync(ctx->d_stage + (size_t)j * ctx->n, d_X + (size_t)b->order[j] * ldx, sizeof(double) * ctx->n,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        src = ctx->d_stage; ld = ctx->n;
    }
    for (size_t p = 0; p < b->panel.size(); ++p)
        if (pack_panel(ctx, src, ld, 0, b->B, (int)p * b->stride, b->panelW[p], b->panel[p])) return 1;
    KEXP_CUDA(cudaGetLastError());
    return 0;
}

int kexp_batch_export_dev(kexp_batch* b, double* d_Y, int64_t ldy) {
    if (!b || !d_Y) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (ldy < ctx->n) return kexp::fail(ctx, "ldy < n");
    if (panels_unset(b)) return kexp::fail(ctx, "a shared-input batch holds one vector; nothing to export");
    if (finish_upload(b)) return 1;
    if (b->order.empty()) {
        for (size_t p = 0; p < b->panel.size(); ++p)
            if (unpack_panel(ctx, b->panel[p], d_Y, ldy, b->B, (int)p * b->stride, b->panelW[p])) return 1;
    } else {
        if (ensure_stage(ctx, sizeof(double) * (size_t)ctx->n * b->B)) return 1;
        for (size_t p = 0; p < b->panel.size(); ++p)
            if (unpack_panel(ctx, b->panel[p], ctx->d_stage, ctx->n, b->B, (int)p * b->stride, b->panelW[p])) return 1;
        for (int j = 0; j < b->B; ++j)
            KEXP_CUDA(cudaMemcpyAsync(d_Y + (size_t)b->order[j] * ldy, ctx->d_stage + (size_t)j * ctx->n, sizeof(double) * ctx->n,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    }
    KEXP_CUDA(cudaGetLastError());
    return 0;
}

// The library's compute stream (cudaStream_t), so that a caller can enqueue its own device work -- the nonlinear slow
// right-hand side between the two exponentials, stage combinations -- in order with the library's, without host joins.
int kexp_get_stream(kexp_ctx* ctx, void** stream_out) {
    if (!ctx || !stream_out) return 1;
    if (swap_lane(ctx, 0)) return 1;
    *stream_out = (void*)ctx->stream;
    return 0;
}

// Asynchronous download: the panels are unpacked on the compute stream into a staging buffer of their own, the
// device-to-host copies run on a separate stream behind them, and the call returns at once -- the copy overlaps
// whatever is enqueued next (e.g. the backward half of an averaged evaluation and its uploads; PCIe is full duplex).
// Y must stay valid until kexp_wait_transfers (pinned memory for a truly asynchronous copy).
int kexp_batch_download_async(kexp_batch* b, double* Y, int64_t ldy) {
    if (!b || !Y) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (ldy < ctx->n) return kexp::fail(ctx, "ldy < n");
    if (b->is_shared && ctx->use_shared_basis) return kexp::fail(ctx, "a shared-input batch holds one vector; nothing to download");
    if (finish_upload(b)) return 1;
    const size_t bytes = sizeof(double) * (size_t)ctx->n * b->B;
    ctx->st_d2h_bytes += (int64_t)bytes;
    if (!ctx->d2h_stream) {
        KEXP_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        KEXP_CUDA(cudaEventCreateWithFlags(&ctx->ev_unpacked, cudaEventDisableTiming));
        KEXP_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h_done, cudaEventDisableTiming));
    }
    if (bytes > ctx->stage_out_bytes) {
        KEXP_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
        KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_stage_out); ctx->d_stage_out = nullptr; ctx->stage_out_bytes = 0;
        KEXP_CUDA(cudaMalloc(&ctx->d_stage_out, bytes));
        ctx->stage_out_bytes = bytes;
    }
    if (ctx->d2h_pending) KEXP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h_done, 0));   // staging still being read
    for (size_t p = 0; p < b->panel.size(); ++p)
        if (unpack_panel(ctx, b->panel[p], ctx->d_stage_out, ctx->n, b->B, (int)p * b->stride, b->panelW[p])) return 1;
    KEXP_CUDA(cudaEventRecord(ctx->ev_unpacked, ctx->stream));
    KEXP_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_unpacked, 0));
    if (b->order.empty()) {
        KEXP_CUDA(cudaMemcpy2DAsync(Y, sizeof(double) * ldy, ctx->d_stage_out, sizeof(double) * ctx->n, sizeof(double) * ctx->n, b->B,
                                    cudaMemcpyDeviceToHost, ctx->d2h_stream));
    } else {
        for (int j = 0; j < b->B; ++j)
            KEXP_CUDA(cudaMemcpyAsync(Y + (size_t)b->order[j] * ldy, ctx->d_stage_out + (size_t)j * ctx->n, sizeof(double) * ctx->n,
                                      cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    KEXP_CUDA(cudaEventRecord(ctx->ev_d2h_done, ctx->d2h_stream));
    ctx->d2h_pending = true;
    return 0;
}

int kexp_wait_transfers(kexp_ctx* ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (ctx->d2h_pending) {
        KEXP_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
        ctx->d2h_pending = false;
    }
    return poll_deferred(ctx, true);
}

int kexp_batch_set_order(kexp_batch* b, const int32_t* order) {
    if (!b) return 1;
    kexp_ctx* ctx = b->ctx;
    if (finish_upload(b)) return 1;
    if (!order) { b->order.clear(); return 0; }
    std::vector<char> seen((size_t)b->B, 0);
    for (int j = 0; j < b->B; ++j) {
        if (order[j] < 0 || order[j] >= b->B || seen[order[j]]) return kexp::fail(ctx, "order is not a permutation of 0..B-1");
        seen[order[j]] = 1;
    }
    b->order.assign(order, order + b->B);
    return 0;
}

// order[j] = caller's column stored at position j, sorted by |t| (stable)
static void order_by_abs_time(const double* t, int B, std::vector<int>& order) {
    order.resize((size_t)B);
    for (int j = 0; j < B; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(t[a]) < std::fabs(t[b]); });
}

// Combination coefficients of the shared-input path: D[m][k] = sum_n a_n E_nm(s_k), where
// S_n(s b) = sum_m E_nm(s) S_m(b).  Multiplication by b in the S basis: b S_0 = S_1,
// b S_m = (S_{m+1} - S_{m-1}) / 2; then E_0 = e_0, E_1 = s shift(E_0), E_n = 2 s shift(E_{n-1}) + E_{n-2}.
static void expansion_column(const std::vector<double>& a, double s, double* Dcol, int stride) {
    const int nb = (int)a.size();
    std::vector<double> e0((size_t)nb + 2, 0.0), e1((size_t)nb + 2, 0.0), e2((size_t)nb + 2, 0.0), sh((size_t)nb + 2, 0.0);
    auto shift = [&](const std::vector<double>& e, std::vector<double>& o) {
        std::fill(o.begin(), o.end(), 0.0);
        o[1] += e[0];
        for (int m = 1; m <= nb; ++m) { o[m + 1] += 0.5 * e[m]; o[m - 1] -= 0.5 * e[m]; }
    };
    e0[0] = 1.0;
    for (int m = 0; m < nb; ++m) Dcol[(size_t)m * stride] = a[0] * e0[m];
    if (nb == 1) return;
    shift(e0, sh);
    for (int m = 0; m <= nb; ++m) e1[m] = s * sh[m];
    for (int m = 0; m < nb; ++m) Dcol[(size_t)m * stride] += a[1] * e1[m];
    for (int n = 2; n < nb; ++n) {
        shift(e1, sh);
        for (int m = 0; m <= nb; ++m) e2[m] = 2.0 * s * sh[m] + e0[m];
        for (int m = 0; m < nb; ++m) Dcol[(size_t)m * stride] += a[n] * e2[m];
        e0.swap(e1); e1.swap(e2);
    }
}

static unsigned long long fnv1a(const void* data, size_t bytes, unsigned long long h = 1469598103934665603ull) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static void finish_plan(PanelPlan& P) {
    P.sig = fnv1a(&P.M, sizeof(P.M));
    P.sig = fnv1a(P.relax.data(), sizeof(double) * P.relax.size(), P.sig);
}

// Number of leading coefficients to keep so that the dropped tail is below eps * sum |b_m|
static int truncated_length(const double* b, int nb, int stride, double eps) {
    double total = 0.0;
    for (int m = 0; m < nb; ++m) total += std::fabs(b[(size_t)m * stride]);
    double tail = 0.0;
    int keep = nb;
    while (keep > 1 && tail + std::fabs(b[(size_t)(keep - 1) * stride]) <= eps * total) { tail += std::fabs(b[(size_t)(keep - 1) * stride]); --keep; }
    return keep;
}

// Can the per-column economised expansions be used for these sample times?
static bool economise(const kexp_ctx* ctx, const ChebTable& tb, const double* t, int B) {
    if (!ctx->use_econ || !(tb.t_valid > 0.0) || tb.ncheb < 2) return false;
    for (int k = 0; k < B; ++k)
        if (std::fabs(t[k]) * (1.0 + ctx->econ_margin) > tb.t_valid) return false;    // outside the stated validity range
    return true;
}

// Plans of all panels of a batch of independent columns.
//   reference series:   y_k = sum_n a_n S_n((t_k/L) A) x_k                                  (cheby_exp.py:110-165)
//   economised:         the spectrum of (t_k/L) A only fills the fraction s_k = |t_k|/t_valid of the Chebyshev
//                       interval, so with C = sign(t_k) (t_valid/L) A / (1+margin) (the same operator for every
//                       column, spectrum in i[-1,1]) the SAME polynomial is  sum_m b_m^(k) S_m(C) x_k,
//                       b^(k) = E(s_k)^T a  (exact re-expansion, expansion_column), whose coefficients fall below
//                       round-off long before m = ncheb when s_k < 1.  Dropped tail <= econ_eps * sum|b|.
static const std::vector<PanelPlan>& batch_plans(kexp_ctx* ctx, int id, const kexp_batch* bt, const double* t) {
    const ChebTable& tb = ctx->cheb[id];
    const int B = bt->B, np = (int)bt->panel.size();
    unsigned long long key = fnv1a(&id, sizeof(id));
    key = fnv1a(&B, sizeof(B), key);
    key = fnv1a(&bt->stride, sizeof(bt->stride), key);
    key = fnv1a(t, sizeof(double) * B, key);
    auto it = ctx->plan_cache.find(key);
    if (it != ctx->plan_cache.end()) return it->second;
    if (ctx->plan_cache.size() >= 64) ctx->plan_cache.clear();
    std::vector<PanelPlan> plans((size_t)np);
    const bool econ = economise(ctx, tb, t, B);
    const int nb = tb.ncheb + 1;
    std::vector<double> bcol((size_t)nb);
    for (int p = 0; p < np; ++p) {
        PanelPlan& P = plans[p];
        const int W = bt->panelW[p], col0 = p * bt->stride;
        for (int c = 0; c < kMaxW; ++c) P.tscale[c] = 0.0;
        if (!econ) {
            P.M = tb.ncheb;
            P.coef.assign((size_t)nb * kMaxW, 0.0);
            for (int m = 0; m < nb; ++m)
                for (int c = 0; c < W; ++c) P.coef[(size_t)m * kMaxW + c] = tb.h_coef[m];
            for (int c = 0; c < W && col0 + c < B; ++c) P.tscale[c] = t[col0 + c] / tb.L;
            P.relax = tb.h_relax;
        } else {
            std::vector<double> full((size_t)nb * kMaxW, 0.0);
            std::vector<int> len((size_t)W, 1);
            int keep = 1;
            const double cscale = tb.t_valid / ((1.0 + ctx->econ_margin) * tb.L);
            for (int c = 0; c < W && col0 + c < B; ++c) {
                const double tk = t[col0 + c];
                expansion_column(tb.h_coef, std::fabs(tk) * (1.0 + ctx->econ_margin) / tb.t_valid, &full[c], kMaxW);
                len[c] = truncated_length(&full[c], nb, kMaxW, ctx->econ_eps);
                keep = std::max(keep, len[c]);
                P.tscale[c] = (tk < 0.0) ? -cscale : cscale;
            }
            P.M = keep - 1;
            P.coef.assign(full.begin(), full.begin() + (size_t)keep * kMaxW);
            // relaxation: the tightest requirement of any column of the panel
            P.relax.assign((size_t)keep, 1e6);
            std::vector<double> rc;
            for (int c = 0; c < W && col0 + c < B; ++c) {
                for (int m = 0; m < keep; ++m) bcol[m] = full[(size_t)m * kMaxW + c];
                relax_factors(bcol.data(), P.M, rc);
                for (int n = 1; n <= P.M; ++n) P.relax[n] = std::min(P.relax[n], rc[n]);
            }
            P.relax[0] = 1.0;
        }
        finish_plan(P);
    }
    return ctx->plan_cache.emplace(key, std::move(plans)).first->second;
}

// A shared-input batch uploaded with the shared-basis path on holds ONE vector (shared_vec); its panels are
// uninitialised device memory and must never be read.
static bool panels_unset(const kexp_batch* b) { return b->is_shared && b->ctx->use_shared_basis; }

// Forward half of an averaged evaluation: one input for all sample times.
static int cheby_apply_shared(kexp_ctx* ctx, int id, kexp_batch* in, kexp_batch* out, const double* t, bool time_it) {
    const ChebTable& tb = ctx->cheb[id];
    const int B = in->B;
    const bool econ = economise(ctx, tb, t, B);
    // basis U_m = S_m(c A) x for ONE column: c = t*/L with t* = max|t_k| (reference series, 29 terms at M1), or
    // c = t_valid / ((1+margin) L) with the series truncated where every column's re-expansion has died out
    double tstar = 0.0;
    for (int k = 0; k < B; ++k) tstar = std::max(tstar, std::fabs(t[k]));
    if (tstar == 0.0) tstar = 1.0;
    const double tbasis = econ ? tb.t_valid / (1.0 + ctx->econ_margin) : tstar;
    int nb = tb.ncheb + 1;
    // coefficients for every output panel, panel p at offset nb * (sum of earlier widths)
    size_t totW = 0;
    for (int W : out->panelW) totW += (size_t)W;
    std::vector<double> D((size_t)nb * totW, 0.0);
    PanelPlan plan;
    {
        size_t off = 0;
        int keep = 1;
        for (size_t p = 0; p < out->panel.size(); ++p) {
            const int W = out->panelW[p];
            for (int c = 0; c < W; ++c) {
                const int k = (int)p * out->stride + c;
                if (k < B) {
                    expansion_column(tb.h_coef, t[k] / tbasis, &D[off + c], W);
                    keep = std::max(keep, econ ? truncated_length(&D[off + c], nb, W, ctx->econ_eps) : nb);
                }
            }
            off += (size_t)nb * W;
        }
        if (keep < nb) {                                   // repack with the shorter row count
            std::vector<double> D2((size_t)keep * totW, 0.0);
            size_t o1 = 0, o2 = 0;
            for (size_t p = 0; p < out->panel.size(); ++p) {
                const int W = out->panelW[p];
                std::copy(D.begin() + o1, D.begin() + o1 + (size_t)keep * W, D2.begin() + o2);
                o1 += (size_t)nb * W; o2 += (size_t)keep * W;
            }
            D.swap(D2);
            nb = keep;
        }
        plan.M = nb - 1;
        for (int c = 0; c < kMaxW; ++c) plan.tscale[c] = 0.0;
        plan.tscale[0] = tbasis / tb.L;
        plan.coef.assign((size_t)nb * kMaxW, 0.0);         // the 2-wide panel's own Y is not used
        // relaxation of the basis run: the tightest requirement of any output column
        plan.relax.assign((size_t)nb, 1e6);
        std::vector<double> bcol((size_t)nb), rc;
        off = 0;
        for (size_t p = 0; p < out->panel.size(); ++p) {
            const int W = out->panelW[p];
            for (int c = 0; c < W; ++c) {
                if ((int)p * out->stride + c >= B) continue;
                for (int m = 0; m < nb; ++m) bcol[m] = D[off + (size_t)m * W + c];
                relax_factors(bcol.data(), plan.M, rc);
                for (int n = 1; n <= plan.M; ++n) plan.relax[n] = std::min(plan.relax[n], rc[n]);
            }
            off += (size_t)nb * W;
        }
        plan.relax[0] = 1.0;
        finish_plan(plan);
    }
    const int ld = (tb.ncheb + 1 + 3) & ~3;
    if (ctx->basis_ld < ld) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_basis); ctx->d_basis = nullptr; ctx->basis_ld = 0;
        KEXP_CUDA(cudaMalloc(&ctx->d_basis, sizeof(double) * (size_t)ctx->n * ld));
        ctx->basis_ld = ld;
        invalidate_graphs(ctx);
    }
    if (swap_lane(ctx, 0)) return 1;
    cudaStreamSynchronize(ctx->stream);              // d_D reuse
    if (ctx->d_D_cap < D.size()) {
        cudaFree(ctx->d_D); ctx->d_D = nullptr; ctx->d_D_cap = 0;
        KEXP_CUDA(cudaMalloc(&ctx->d_D, sizeof(double) * D.size()));
        ctx->d_D_cap = D.size();
    }
    KEXP_CUDA(cudaMemcpy(ctx->d_D, D.data(), sizeof(double) * D.size(), cudaMemcpyHostToDevice));
    if (ensure_panels(ctx, 2)) return 1;
    warm_occupancy<2>(ctx);
    max_grid(ctx, basis_store_kernel<2>);
    if (time_it) KEXP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if (set_plan(ctx, plan)) return 1;
    KEXP_CUDA(cudaMemcpyAsync(ctx->S0, in->shared_vec, sizeof(double) * 2 * (size_t)ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->capture_basis = true;
    const int rc = run_cheby_panel(ctx, id, 2, plan);
    ctx->capture_basis = false;
    if (rc) return 1;
    size_t off = 0;
    for (size_t p = 0; p < out->panel.size(); ++p) {
        const int W = out->panelW[p];
        TimedLaunch tl(ctx, KC_MISC);
        DISPATCH_W(W, launch_combine<kW>(ctx, nb, ctx->d_D + off, out->panel[p]));
        off += (size_t)nb * W;
    }
    KEXP_CUDA(cudaGetLastError());
    if (time_it) KEXP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused weighted average  v = sum_k w_k exp(t_k L) X_k  as ONE Clenshaw recurrence (see
// clenshaw_combine_kernel).  t and w are in the stored column order of `in`.
// ---------------------------------------------------------------------------------------------
static int run_clenshaw(kexp_ctx* ctx, int id, const PanelPlan& plan) {
    constexpr int W = 2;
    if (narrow_eligible(ctx, W)) {
        // step j = 1..M of the persistent kernel runs Clenshaw index m = M - j (relaxation is stored by m)
        std::vector<int> kterm((size_t)plan.M + 1, 2);
        std::vector<double> r2((size_t)plan.M + 1, 0.0);
        for (int j = 1; j <= plan.M; ++j) {
            const double relax = ctx->use_relax ? plan.relax[plan.M - j] : 1.0;
            kterm[j] = cheb_iterations(ctx, relax);
            r2[j] = (ctx->rtol * relax) * (ctx->rtol * relax);
        }
        TimedLaunch tl(ctx, KC_SPMM_CG);
        if (launch_narrow_apply(ctx, W, plan.M, 1, kterm.data(), r2.data(), true, nullptr)) return 1;
        ctx->st_op_apps += plan.M;
        return 0;
    }
    if (ctx->timing || !ctx->use_graph) {
        enqueue_clenshaw<W>(ctx, plan);
        KEXP_CUDA(cudaGetLastError());
        return 0;
    }
    const GraphKey key{id + 2000, W, plan.sig};
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        if (ctx->graphs.size() >= 48) {
            cudaStreamSynchronize(ctx->stream);
            for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.second);
            ctx->graphs.clear();
        }
        cudaGraph_t graph;
        int64_t l0 = ctx->st_launches, a0 = ctx->st_op_apps;
        KEXP_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_clenshaw<W>(ctx, plan);
        KEXP_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
        ctx->st_launches = l0; ctx->st_op_apps = a0;
        cudaGraphExec_t exec;
        KEXP_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = ctx->graphs.emplace(key, exec).first;
    }
    KEXP_CUDA(cudaGraphLaunch(it->second, ctx->stream));
    ctx->st_op_apps += plan.M;
    ctx->st_launches += 1 + (int64_t)plan.M * (3 + (ctx->solver == 1 ? (ctx->use_coop ? 0 : ctx->graph_its - 1) : 2 * ctx->graph_its) +
                                               ((ctx->solver == 1 && ctx->use_fused_update) ? 0 : 1) + (ctx->n_b > 0) + ((ctx->n_d > 0 && !ctx->depth_rows_fused) ? 1 : 0));
    return 0;
}

static int cheby_average_fused_impl(kexp_ctx* ctx, int id, kexp_batch* in, const double* t, const double* w, double* v_dev) {
    if (check_ready(ctx, id)) return 1;
    if (!in || !t || !w || !v_dev) return kexp::fail(ctx, "null argument");
    cudaSetDevice(ctx->device);
    const ChebTable& tb = ctx->cheb[id];
    const int B = in->B;
    const bool shared = in->is_shared && in->shared_vec && ctx->use_shared_basis;
    if (swap_lane(ctx, 0)) return 1;
    if (finish_upload(in)) return 1;
    const bool econ = economise(ctx, tb, t, B);
    double tstar = 0.0;
    for (int k = 0; k < B; ++k) tstar = std::max(tstar, std::fabs(t[k]));
    if (tstar == 0.0) tstar = 1.0;
    const double tbasis = econ ? tb.t_valid / (1.0 + ctx->econ_margin) : tstar;
    const int nbf = tb.ncheb + 1;
    // b^(k) for every column, the common truncation, and the per-step tolerance relaxation
    std::vector<double> D((size_t)nbf * B);
    int keep = 1;
    for (int k = 0; k < B; ++k) {
        expansion_column(tb.h_coef, t[k] / tbasis, &D[k], B);
        if (w[k] != 0.0) keep = std::max(keep, truncated_length(&D[k], nbf, B, ctx->econ_eps));
    }
    PanelPlan plan;
    plan.M = keep - 1;
    for (int c = 0; c < kMaxW; ++c) plan.tscale[c] = 0.0;
    plan.tscale[0] = tbasis / tb.L;
    plan.coef.assign((size_t)kMaxW, 0.0);
    // step m applies the operator to u_{m+1}; an error there reaches v through a polynomial of degree m
    // (norm <= m+1), and u_{m+1} only holds the share sum_{j>m} (j-m) |b_j| of a column's series
    plan.relax.assign((size_t)keep, 1e6);
    for (int k = 0; k < B; ++k) {
        if (w[k] == 0.0) continue;
        std::vector<double> om((size_t)keep, 0.0);
        double omax = 0.0;
        for (int m = 0; m < plan.M; ++m) {
            double g = 0.0;
            for (int j = m + 1; j <= plan.M; ++j) g += (double)(j - m) * std::fabs(D[(size_t)j * B + k]);
            om[m] = (double)(m + 1) * g;
            omax = std::max(omax, om[m]);
        }
        for (int m = 0; m < plan.M; ++m)
            if (om[m] > 0.0) plan.relax[m] = std::min(plan.relax[m], std::max(1.0, omax / om[m]));
    }
    finish_plan(plan);
    // Z = inputs x (w_k b_m^(k)):   [n][ld]
    const int ld = (tb.ncheb + 1 + 3) & ~3;
    if (ctx->basis_ld < ld) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_basis); ctx->d_basis = nullptr; ctx->basis_ld = 0;
        KEXP_CUDA(cudaMalloc(&ctx->d_basis, sizeof(double) * (size_t)ctx->n * ld));
        ctx->basis_ld = ld;
        invalidate_graphs(ctx);
    }
    std::vector<double> G;
    std::vector<size_t> goff;
    if (shared) {
        G.assign((size_t)keep * 2, 0.0);
        for (int m = 0; m < keep; ++m)
            for (int k = 0; k < B; ++k) G[(size_t)m * 2] += w[k] * D[(size_t)m * B + k];
        goff.push_back(0);
    } else {
        for (size_t p = 0; p < in->panel.size(); ++p) {
            const int W = in->panelW[p], col0 = (int)p * in->stride;
            goff.push_back(G.size());
            G.resize(G.size() + (size_t)keep * W, 0.0);
            double* g = &G[goff.back()];
            for (int m = 0; m < keep; ++m)
                for (int c = 0; c < W && col0 + c < B; ++c) g[(size_t)m * W + c] = w[col0 + c] * D[(size_t)m * B + col0 + c];
        }
    }
    cudaStreamSynchronize(ctx->stream);              // d_D reuse
    if (ctx->d_D_cap < G.size()) {
        cudaFree(ctx->d_D); ctx->d_D = nullptr; ctx->d_D_cap = 0;
        KEXP_CUDA(cudaMalloc(&ctx->d_D, sizeof(double) * G.size()));
        ctx->d_D_cap = G.size();
    }
    KEXP_CUDA(cudaMemcpy(ctx->d_D, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice));
    if (ensure_panels(ctx, 2)) return 1;
    warm_occupancy<2>(ctx);
    KEXP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if (shared) {
        TimedLaunch tl(ctx, KC_MISC);
        launch_clenshaw_combine<2>(ctx, in->shared_vec, ctx->d_D, keep, 0);
    } else {
        for (size_t p = 0; p < in->panel.size(); ++p) {
            const int W = in->panelW[p];
            TimedLaunch tl(ctx, KC_MISC);
            DISPATCH_W(W, launch_clenshaw_combine<kW>(ctx, in->panel[p], ctx->d_D + goff[p], keep, p > 0));
        }
    }
    KEXP_CUDA(cudaGetLastError());
    if (set_plan(ctx, plan)) return 1;
    if (run_clenshaw(ctx, id, plan)) return 1;
    // the result is column 0 of the buffer written last (M swaps after starting in S0)
    const double* res = (plan.M % 2 == 0) ? ctx->S0 : ctx->S1;
    const double one[2] = {1.0, 0.0};
    if (colsum_panel(ctx, res, one, 0, 2, 2, v_dev, 0)) return 1;
    KEXP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    KEXP_CUDA(cudaGetLastError());
    return mark_consumed(ctx, in);
}

static int cheby_apply_dev_impl(kexp_ctx* ctx, int id, kexp_batch* in, kexp_batch* out, const double* t, bool time_it) {
    if (check_ready(ctx, id)) return 1;
    if (!in || !out || !t) return kexp::fail(ctx, "null argument");
    if (in == out) return kexp::fail(ctx, "x and y must not alias (cheby_exp.py:119,127)");
    if (in->B != out->B) return kexp::fail(ctx, "batch sizes differ");
    if (in->is_shared && in->shared_vec && ctx->use_shared_basis) return cheby_apply_shared(ctx, id, in, out, t, time_it);
    if (panels_unset(in)) return kexp::fail(ctx, "the input batch holds one shared vector, its panels were never filled");
    cudaSetDevice(ctx->device);
    if (in->stride != out->stride || in->panelW != out->panelW) return kexp::fail(ctx, "batches have different panel layouts");
    int maxW = 0;
    for (int W : in->panelW) maxW = std::max(maxW, W);
    for (int W : in->panelW) DISPATCH_W(W, warm_occupancy<kW>(ctx));
    const int np = (int)in->panel.size();
    const std::vector<PanelPlan>& plans = batch_plans(ctx, id, in, t);
    const bool two = ctx->nlanes == 2 && !ctx->timing && np >= 2;
    if (swap_lane(ctx, 0)) return 1;
    if (time_it) KEXP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    if (two) KEXP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    bool lane1_used = false;
    for (int p = 0; p < np; ++p) {
        const int lane = two ? (p & 1) : 0;
        if (swap_lane(ctx, lane)) return 1;
        if (lane == 1 && !lane1_used) {                 // lane 1 starts after everything already queued on lane 0
            KEXP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fork, 0));
            lane1_used = true;
        }
        const int W = in->panelW[p];
        if (ensure_panels(ctx, W)) return 1;
        if (in->pending) KEXP_CUDA(cudaStreamWaitEvent(ctx->stream, in->ready[p], 0));     // panel still being uploaded
        const size_t bytes = sizeof(double) * (size_t)ctx->n * W;
        if (set_plan(ctx, plans[p])) return 1;
        KEXP_CUDA(cudaMemcpyAsync(ctx->S0, in->panel[p], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (run_cheby_panel(ctx, id, W, plans[p])) return 1;
        KEXP_CUDA(cudaMemcpyAsync(out->panel[p], ctx->Y, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (lane1_used) {
        if (swap_lane(ctx, 1)) return 1;
        KEXP_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream));
        if (swap_lane(ctx, 0)) return 1;
        KEXP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    if (swap_lane(ctx, 0)) return 1;
    if (time_it) KEXP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    (void)maxW;
    return mark_consumed(ctx, in);
}

// The CG loop inside the graph has a fixed number of early-exit iterations.  After the
// first apply it is trimmed to the observed maximum + 2; every apply then checks (one small
// read-back) that no solve missed its tolerance and otherwise repeats itself with more
// iterations, up to max_it.  A solve that still misses rtol at max_it is an ERROR (the result
// is not delivered silently).  In asynchronous mode (kexp_set_async) the read-back of a tuned
// context is deferred: it is copied to pinned memory behind the apply and inspected by a later
// call, at the latest by kexp_synchronize / kexp_batch_download / kexp_stats.
static int read_cg_counters(kexp_ctx* ctx, int* max_iters, int* unconverged, double* worst) {
    *max_iters = 0; *unconverged = 0; *worst = 0.0;
    return for_each_lane(ctx, [&]() -> int {
        CgScalars* h = ctx->h_sc;
        const size_t off = offsetof(CgScalars, max_iters);
        KEXP_CUDA(cudaMemcpyAsync((char*)h + off, (const char*)ctx->d_sc + off, sizeof(CgScalars) - off, cudaMemcpyDeviceToHost, ctx->stream));
        KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
        *max_iters = std::max(*max_iters, h->max_iters); *unconverged += h->unconverged;
        *worst = std::max(*worst, h->worst_rel_res);
        return 0;
    });
}

static int unconverged_error(kexp_ctx* ctx, double worst, bool deferred) {
    char buf[320];
    snprintf(buf, sizeof(buf), "mass solve did not reach rtol %.3g within max_it = %d iterations: max rel residual %.3e%s",
             ctx->rtol, ctx->max_it, worst,
             deferred ? " (detected after an asynchronous apply: results computed since the last check are invalid)" : "");
    return kexp::fail(ctx, buf);
}

static void raise_iterations(kexp_ctx* ctx) {
    ctx->graph_its = (ctx->solver == 1) ? std::min(ctx->max_it, ctx->graph_its + 2) : ctx->max_it;
    ctx->its_tuned = true;              // do not trim again
    invalidate_graphs(ctx);
}

// Look at a deferred read-back.  block = false: only if the copy has already finished.
static int poll_deferred(kexp_ctx* ctx, bool block) {
    if (!ctx->chk_pending) return 0;
    if (block) KEXP_CUDA(cudaEventSynchronize(ctx->ev_chk));
    else if (cudaEventQuery(ctx->ev_chk) != cudaSuccess) { cudaGetLastError(); return 0; }
    ctx->chk_pending = false;
    const int unc = ctx->h_sc->unconverged;
    if (unc > ctx->unconverged_seen) {
        ctx->unconverged_seen = unc;
        const double worst = ctx->h_sc->worst_rel_res;
        if (ctx->graph_its < ctx->max_it) raise_iterations(ctx);       // later applies get more iterations
        return unconverged_error(ctx, worst, true);
    }
    return 0;
}

extern "C++" {
template <typename F>
static int run_checked(kexp_ctx* ctx, F enqueue) {
    if (poll_deferred(ctx, false)) return 1;
    for (;;) {
        if (enqueue()) return 1;
        if (ctx->timing) drain_events(ctx);
        if (!ctx->adaptive_its) return 0;
        if (ctx->async_checks && ctx->its_tuned && ctx->nlanes == 1 && !ctx->timing) {
            if (ctx->chk_pending) return 0;                            // the pending copy covers the counters (they only grow)
            const size_t off = offsetof(CgScalars, max_iters);
            KEXP_CUDA(cudaMemcpyAsync((char*)ctx->h_sc + off, (const char*)ctx->d_sc + off, sizeof(CgScalars) - off, cudaMemcpyDeviceToHost, ctx->stream));
            KEXP_CUDA(cudaEventRecord(ctx->ev_chk, ctx->stream));
            ctx->chk_pending = true;
            return 0;
        }
        int mx = 0, unc = 0;
        double worst = 0.0;
        if (read_cg_counters(ctx, &mx, &unc, &worst)) return 1;
        if (unc > ctx->unconverged_seen) {
            // a solve hit the trimmed cap (CG) or missed rtol with the fixed length (Chebyshev): raise it and redo
            ctx->unconverged_seen = unc;
            if (ctx->graph_its >= ctx->max_it) return unconverged_error(ctx, worst, false);
            raise_iterations(ctx);
            continue;
        }
        if (!ctx->its_tuned && mx > 0) {
            const int want = std::min(ctx->max_it, mx + 2);
            if (want < ctx->graph_its) { ctx->graph_its = want; invalidate_graphs(ctx); }
            ctx->its_tuned = true;
        }
        return 0;
    }
}
}  // extern "C++"

int kexp_cheby_apply_dev(kexp_ctx* ctx, int id, kexp_batch* in, kexp_batch* out, const double* t) {
    if (!ctx) return 1;
    std::vector<double> ts;
    if (in && out && t && in->B == out->B) {
        // t is given in the caller's column order; the output takes the stored order of a per-column input
        if (!(in->is_shared && ctx->use_shared_basis) && in != out) out->order = in->order;
        if (!out->order.empty()) {
            ts.resize((size_t)out->B);
            for (int j = 0; j < out->B; ++j) ts[j] = t[out->order[j]];
            t = ts.data();
        }
    }
    return run_checked(ctx, [&]() { return cheby_apply_dev_impl(ctx, id, in, out, t, true); });
}

int kexp_set_async(kexp_ctx* ctx, int on) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!on && poll_deferred(ctx, true)) return 1;
    ctx->async_checks = on != 0;
    return 0;
}

int kexp_synchronize(kexp_ctx* ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
    KEXP_CUDA(cudaGetLastError());
    return poll_deferred(ctx, true);
}

int kexp_timer_start(kexp_ctx* ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!ctx->ev_t0) { KEXP_CUDA(cudaEventCreate(&ctx->ev_t0)); KEXP_CUDA(cudaEventCreate(&ctx->ev_t1)); }
    KEXP_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return 0;
}

int kexp_timer_stop(kexp_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->ev_t0) return 1;
    cudaSetDevice(ctx->device);
    KEXP_CUDA(cudaEventRecord(ctx->ev_t1, ctx->stream));
    KEXP_CUDA(cudaEventSynchronize(ctx->ev_t1));
    float f = 0.f;
    KEXP_CUDA(cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
    *ms = f;
    return 0;
}

int kexp_last_device_ms(kexp_ctx* ctx, double* ms) {
    if (!ctx || !ms) return 1;
    cudaSetDevice(ctx->device);
    KEXP_CUDA(cudaEventSynchronize(ctx->ev_b));
    float f = 0.f;
    KEXP_CUDA(cudaEventElapsedTime(&f, ctx->ev_a, ctx->ev_b));
    *ms = f;
    return 0;
}

int kexp_batch_average_dev(kexp_batch* b, const double* w, double* v_dev) {
    if (!b || !w || !v_dev) return 1;
    kexp_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    if (panels_unset(b)) return kexp::fail(ctx, "a shared-input batch holds one vector; there are no columns to average");
    if (finish_upload(b)) return 1;
    std::vector<double> ws;
    if (!b->order.empty()) {
        ws.resize((size_t)b->B);
        for (int j = 0; j < b->B; ++j) ws[j] = w[b->order[j]];
        w = ws.data();
    }
    for (size_t p = 0; p < b->panel.size(); ++p)
        if (colsum_panel(ctx, b->panel[p], w, (int)p * b->stride, b->B, b->panelW[p], v_dev, p > 0)) return 1;
    return 0;
}

// scratch batches of the host-pointer entry points, kept between calls (cudaMalloc/cudaFree
// of ~100 MB panels would otherwise dominate small applies)
static int scratch_batches(kexp_ctx* ctx, int B, kexp_batch** bi, kexp_batch** bo) {
    if (ctx->tmp_in && ctx->tmp_in->B != B) {
        kexp_batch_destroy(ctx->tmp_in); kexp_batch_destroy(ctx->tmp_out);
        ctx->tmp_in = ctx->tmp_out = nullptr;
    }
    if (!ctx->tmp_in) {
        if (kexp_batch_create(ctx, B, &ctx->tmp_in)) return 1;
        if (kexp_batch_create(ctx, B, &ctx->tmp_out)) { kexp_batch_destroy(ctx->tmp_in); ctx->tmp_in = nullptr; return 1; }
    }
    *bi = ctx->tmp_in; *bo = ctx->tmp_out;
    return 0;
}

// Host-array entry points own their batches, so they are free to store the columns sorted by |t_k| when the
// economised series are in use (more than one panel: otherwise the order makes no difference).
static int set_scratch_order(kexp_ctx* ctx, int id, kexp_batch* bi, kexp_batch* bo, const double* t, int B) {
    if (finish_upload(bi) || finish_upload(bo)) return 1;
    bi->order.clear(); bo->order.clear();
    if (bo->panel.size() > 1 && economise(ctx, ctx->cheb[id], t, B)) {
        order_by_abs_time(t, B, bi->order);
        bo->order = bi->order;
    }
    return 0;
}

int kexp_cheby_apply_batch(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                           double* Y, int64_t ldy, const double* t, int B) {
    if (check_ready(ctx, id)) return 1;
    if (!X || !Y || !t || B < 1) return kexp::fail(ctx, "bad arguments");
    if (X == Y) return kexp::fail(ctx, "x and y must not alias (cheby_exp.py:119,127)");
    kexp_batch *bi = nullptr, *bo = nullptr;
    int rc = scratch_batches(ctx, B, &bi, &bo);
    if (!rc) rc = set_scratch_order(ctx, id, bi, bo, t, B);
    if (!rc) rc = batch_upload_impl(bi, X, ldx, x_is_shared, false);
    if (!rc) rc = kexp_cheby_apply_dev(ctx, id, bi, bo, t);
    if (!rc) rc = kexp_batch_download(bo, Y, ldy);
    return rc;
}

// kexp_cheby_apply_batch whose device-to-host copy of Y is left in flight: returns once everything is enqueued and
// the compute has been checked; Y is complete after kexp_wait_transfers.
int kexp_cheby_apply_batch_async(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                                 double* Y, int64_t ldy, const double* t, int B) {
    if (check_ready(ctx, id)) return 1;
    if (!X || !Y || !t || B < 1) return kexp::fail(ctx, "bad arguments");
    if (X == Y) return kexp::fail(ctx, "x and y must not alias (cheby_exp.py:119,127)");
    kexp_batch *bi = nullptr, *bo = nullptr;
    int rc = scratch_batches(ctx, B, &bi, &bo);
    if (!rc) rc = set_scratch_order(ctx, id, bi, bo, t, B);
    if (!rc) rc = batch_upload_impl(bi, X, ldx, x_is_shared, false);
    if (!rc) rc = kexp_cheby_apply_dev(ctx, id, bi, bo, t);
    if (!rc) rc = kexp_batch_download_async(bo, Y, ldy);
    return rc;
}

int kexp_cheby_apply(kexp_ctx* ctx, int id, const double* x, double* y, double t) {
    if (!ctx) return 1;
    return kexp_cheby_apply_batch(ctx, id, x, ctx->n, 0, y, ctx->n, &t, 1);
}

int kexp_allreduce_sum_dev(kexp_ctx* ctx, double* v_dev, int64_t count);

int kexp_cheby_apply_average(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                             const double* t, const double* w, int B, double* v_out, int do_allreduce) {
    if (check_ready(ctx, id)) return 1;
    if (!X || !t || !w || !v_out || B < 1) return kexp::fail(ctx, "bad arguments");
    kexp_batch *bi = nullptr, *bo = nullptr;
    int rc = scratch_batches(ctx, B, &bi, &bo);
    if (!rc) rc = set_scratch_order(ctx, id, bi, bo, t, B);
    if (!rc) rc = batch_upload_impl(bi, X, ldx, x_is_shared, false);
    if (!rc) rc = kexp_cheby_apply_dev(ctx, id, bi, bo, t);
    if (!rc) rc = kexp_batch_average_dev(bo, w, ctx->d_v);
    if (!rc && do_allreduce) rc = kexp_allreduce_sum_dev(ctx, ctx->d_v, ctx->n);
    if (!rc) {
        ctx->st_d2h_bytes += (int64_t)sizeof(double) * ctx->n;
        cudaError_t e = cudaMemcpyAsync(v_out, ctx->d_v, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = kexp::fail(ctx, cudaGetErrorString(e));
    }
    return rc;
}

// Host-only helper (no device, no context): the re-expansion  sum_n a_n S_n(s z) = sum_m b_m S_m(z)  used by the
// economised series, the shared-input basis and the fused average, and the number of leading coefficients kept
// for a dropped tail <= eps * sum|b_m|.  Exposed so that the coefficient algebra can be checked on a CPU.
int kexp_series_reexpand(const double* a, int ncheb, double s, double eps, double* b_out, int* keep_out) {
    if (!a || !b_out || ncheb < 0 || !(std::fabs(s) <= 1.0)) return 1;
    std::vector<double> av(a, a + ncheb + 1);
    expansion_column(av, s, b_out, 1);
    if (keep_out) *keep_out = truncated_length(b_out, ncheb + 1, 1, eps > 0.0 ? eps : 0.0);
    return 0;
}

int kexp_cheby_average_fused_dev(kexp_ctx* ctx, int id, kexp_batch* in, const double* t, const double* w, double* v_dev) {
    if (!ctx) return 1;
    if (!in || !t || !w || !v_dev) return kexp::fail(ctx, "null argument");
    std::vector<double> ts, ws;
    if (!in->order.empty()) {
        ts.resize((size_t)in->B); ws.resize((size_t)in->B);
        for (int j = 0; j < in->B; ++j) { ts[j] = t[in->order[j]]; ws[j] = w[in->order[j]]; }
        t = ts.data(); w = ws.data();
    }
    return run_checked(ctx, [&]() { return cheby_average_fused_impl(ctx, id, in, t, w, v_dev); });
}

int kexp_cheby_apply_average_fused(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                                   const double* t, const double* w, int B, double* v_out, int do_allreduce) {
    if (check_ready(ctx, id)) return 1;
    if (!X || !t || !w || !v_out || B < 1) return kexp::fail(ctx, "bad arguments");
    kexp_batch *bi = nullptr, *bo = nullptr;
    int rc = scratch_batches(ctx, B, &bi, &bo);
    if (!rc) rc = finish_upload(bi);
    if (!rc) { bi->order.clear(); rc = batch_upload_impl(bi, X, ldx, x_is_shared, false); }
    if (!rc) rc = kexp_cheby_average_fused_dev(ctx, id, bi, t, w, ctx->d_v);
    if (!rc && do_allreduce) rc = kexp_allreduce_sum_dev(ctx, ctx->d_v, ctx->n);
    if (!rc) {
        ctx->st_d2h_bytes += (int64_t)sizeof(double) * ctx->n;
        cudaError_t e = cudaMemcpyAsync(v_out, ctx->d_v, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = kexp::fail(ctx, cudaGetErrorString(e));
    }
    return rc;
}

int kexp_operator_apply_dev(kexp_ctx* ctx, kexp_batch* in, kexp_batch* out) {
    if (!ctx) return 1;
    if (!ctx->have_op) return kexp::fail(ctx, "operator not set");
    if (!in || !out || in->B != out->B) return kexp::fail(ctx, "bad batches");
    if (panels_unset(in)) return kexp::fail(ctx, "shared-input batches are only valid as the input of kexp_cheby_apply_dev / kexp_cheby_average_fused_dev");
    cudaSetDevice(ctx->device);
    if (swap_lane(ctx, 0)) return 1;
    if (finish_upload(in)) return 1;
    if (in != out) out->order = in->order;
    if (in->stride != out->stride || in->panelW != out->panelW) return kexp::fail(ctx, "batches have different panel layouts");
    int maxW = 0;
    for (int W : in->panelW) maxW = std::max(maxW, W);
    if (ensure_panels(ctx, maxW)) return 1;
    for (int W : in->panelW) DISPATCH_W(W, warm_occupancy<kW>(ctx));
    // same policy as the Chebyshev applies: a solve that misses rtol makes the call repeat itself with more iterations and
    // fail at max_it; in asynchronous mode the record is inspected by a later call
    return run_checked(ctx, [&]() -> int {
        KEXP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
        for (size_t p = 0; p < in->panel.size(); ++p) {
            const int W = in->panelW[p];
            DISPATCH_W(W, launch_operator_full<kW>(ctx, in->panel[p], out->panel[p]));
            record_solve_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_sc);
        }
        KEXP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
        KEXP_CUDA(cudaGetLastError());
        return 0;
    });
}

int kexp_spmm_K_dev(kexp_ctx* ctx, kexp_batch* in, kexp_batch* out) {
    if (!ctx) return 1;
    if (!ctx->have_op) return kexp::fail(ctx, "operator not set");
    if (!in || !out || in->B != out->B || in == out) return kexp::fail(ctx, "bad batches");
    if (panels_unset(in)) return kexp::fail(ctx, "shared-input batches are only valid as the input of kexp_cheby_apply_dev / kexp_cheby_average_fused_dev");
    cudaSetDevice(ctx->device);
    if (finish_upload(in)) return 1;
    out->order = in->order;
    KEXP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    for (size_t p = 0; p < in->panel.size(); ++p) {
        const int W = in->panelW[p];
        TimedLaunch tl(ctx, KC_SPMM_K);
        DISPATCH_W(W, launch_spmm_K<kW>(ctx, in->panel[p], out->panel[p]));
    }
    KEXP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    KEXP_CUDA(cudaGetLastError());
    if (ctx->timing) drain_events(ctx);
    return 0;
}

// ------------------------------------------------------------------ stats / timing
int kexp_set_kernel_timing(kexp_ctx* ctx, int on) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->timing = on != 0;
    return 0;
}

int kexp_get_spmm_cg_ms(kexp_ctx* ctx, double* ms) {
    if (!ctx || !ms) return 1;
    *ms = ctx->cls[KC_SPMM_CG].ms;
    return 0;
}

// ms / launches / model bytes of one kernel class since the last reset (class ids: KClass)
int kexp_get_class_stats(kexp_ctx* ctx, int cls, double* ms, int64_t* launches, double* bytes_per_launch_per_col) {
    if (!ctx || cls < 0 || cls >= KC_COUNT) return 1;
    if (ms) *ms = ctx->cls[cls].ms;
    if (launches) *launches = ctx->cls[cls].launches;
    if (bytes_per_launch_per_col) {
        const double ns = (double)ctx->n_s, n = (double)ctx->n;
        double b = 0.0;
        switch (cls) {
            case KC_SPMM_K: b = 16.0 * n; break;
            case KC_CG_INIT: b = 24.0 * ns; break;
            case KC_SPMM_CG: b = (ctx->solver == 1 ? 32.0 : 24.0) * ns; break;
            case KC_CG_XR: b = 56.0 * ns; break;
            case KC_CG_P: b = 24.0 * ns; break;
            case KC_CHEB: b = 40.0 * n; break;
            default: b = 0.0;
        }
        *bytes_per_launch_per_col = b;
    }
    return 0;
}

// matrix bytes read once per launch by a kernel class (independent of the batch width)
int kexp_get_class_matrix_bytes(kexp_ctx* ctx, int cls, double* bytes) {
    if (!ctx || !bytes) return 1;
    double b = 0.0;
    const double pinv = 24.0 * (double)ctx->n_s;
    switch (cls) {
        case KC_SPMM_K: b = ctx->Kt.bytes(); break;
        case KC_CG_INIT: b = pinv; break;
        case KC_SPMM_CG: b = ctx->As.bytes(); break;
        case KC_CG_XR: b = pinv; break;
        case KC_CG_P: b = pinv; break;
        case KC_CHEB: b = ctx->Rt.bytes() + (ctx->d_Dblk ? 24.0 * (double)ctx->n_d : 0.0); break;
        default: b = 0.0;
    }
    *bytes = b;
    return 0;
}

int kexp_stats(kexp_ctx* ctx, kexp_stats_t* out) {
    if (!ctx || !out) return 1;
    cudaSetDevice(ctx->device);
    memset(out, 0, sizeof(*out));
    out->n = ctx->n;
    out->n_solve = ctx->n_s;
    int rc = for_each_lane(ctx, [&]() -> int {
        KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
        CgScalars sc;
        KEXP_CUDA(cudaMemcpy(&sc, ctx->d_sc, sizeof(CgScalars), cudaMemcpyDeviceToHost));
        out->cg_solves += (int64_t)(sc.total_solves - ctx->base_solves);
        out->cg_iterations += (int64_t)(sc.total_iters - ctx->base_iters);
        out->last_max_rel_residual = std::max(out->last_max_rel_residual, sc.max_rel_res);
        out->unconverged_solves += sc.unconverged;
        out->worst_rel_residual = std::max(out->worst_rel_residual, sc.worst_rel_res);
        return 0;
    });
    if (rc) return rc;
    out->mass_iterations_per_solve = ctx->graph_its;
    out->h2d_vector_bytes = ctx->st_h2d_bytes;
    out->d2h_vector_bytes = ctx->st_d2h_bytes;
    out->solve_path = ctx->solve_path;
    if (poll_deferred(ctx, true)) return 1;
    out->op_applications = ctx->st_op_apps;
    out->kernel_launches = ctx->st_launches;
    out->bicg_iterations = ctx->st_bicg;
    out->spmm_cg_launches = out->cg_iterations;
    out->spmm_cg_bytes = 0.0;     // per-column figures: see kexp_get_class_stats
    out->model_bytes = 0.0;
    return 0;
}

int kexp_stats_reset(kexp_ctx* ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    int rc = for_each_lane(ctx, [&]() -> int {
        KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
        CgScalars sc;
        KEXP_CUDA(cudaMemcpy(&sc, ctx->d_sc, sizeof(CgScalars), cudaMemcpyDeviceToHost));
        ctx->base_iters = sc.total_iters; ctx->base_solves = sc.total_solves;
        return 0;
    });
    if (rc) return rc;
    ctx->st_op_apps = 0; ctx->st_launches = 0; ctx->st_bicg = 0;
    ctx->st_h2d_bytes = 0; ctx->st_d2h_bytes = 0;
    for (int i = 0; i < KC_COUNT; ++i) ctx->cls[i] = ClassStat();
    return 0;
}

// ------------------------------------------------------------------ NCCL (dlopen: uses the
// libnccl already loaded in the process, e.g. torch's, else the system one)
typedef struct { char internal[128]; } kexp_nccl_id_t;

static void* nccl_handle(kexp_ctx* ctx) {
    void* h = ctx ? ctx->nccl_lib : nullptr;
    if (h) return h;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (ctx) ctx->nccl_lib = h;
    return h;
}

int kexp_comm_unique_id(void* id128) {
    if (!id128) return 1;
    void* h = nccl_handle(nullptr);
    if (!h) return 2;
    typedef int (*fn_t)(kexp_nccl_id_t*);
    fn_t f = (fn_t)dlsym(h, "ncclGetUniqueId");
    if (!f) return 3;
    return f((kexp_nccl_id_t*)id128);
}

int kexp_comm_init(kexp_ctx* ctx, const void* id128, int rank, int nranks) {
    if (!ctx || !id128) return 1;
    cudaSetDevice(ctx->device);
    void* h = nccl_handle(ctx);
    if (!h) return kexp::fail(ctx, "libnccl.so.2 not found");
    typedef int (*fn_t)(void**, int, kexp_nccl_id_t, int);
    fn_t f = (fn_t)dlsym(h, "ncclCommInitRank");
    if (!f) return kexp::fail(ctx, "ncclCommInitRank not found");
    kexp_nccl_id_t id;
    memcpy(&id, id128, sizeof(id));
    int rc = f(&ctx->nccl_comm, nranks, id, rank);
    if (rc) return kexp::fail(ctx, "ncclCommInitRank failed: " + std::to_string(rc));
    ctx->rank = rank; ctx->nranks = nranks;
    return 0;
}

int kexp_allreduce_sum_dev(kexp_ctx* ctx, double* v_dev, int64_t count) {
    if (!ctx || !v_dev) return 1;
    if (!ctx->nccl_comm) {
        if (ctx->nranks == 1) return 0;
        return kexp::fail(ctx, "kexp_comm_init not called");
    }
    cudaSetDevice(ctx->device);
    typedef int (*fn_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    fn_t f = (fn_t)dlsym(ctx->nccl_lib, "ncclAllReduce");
    if (!f) return kexp::fail(ctx, "ncclAllReduce not found");
    const int ncclFloat64 = 8, ncclSum = 0;
    int rc = f(v_dev, v_dev, (size_t)count, ncclFloat64, ncclSum, ctx->nccl_comm, ctx->stream);
    if (rc) return kexp::fail(ctx, "ncclAllReduce failed: " + std::to_string(rc));
    return 0;
}

}  // extern "C"
