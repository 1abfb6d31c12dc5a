// kexp_internal.h -- host-side state shared by the translation units of libkexp.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/kexp.h"
#include "kexp_kernels.cuh"

namespace kexp {

struct DevCsr {
    int* rowptr = nullptr;
    int* col = nullptr;
    double* val = nullptr;
    int nrows = 0;
    int ncols = 0;
    int64_t nnz = 0;
    // BSR(3x3) copy (80-byte blocks: 9 values + block column), built when the matrix has
    // the 3-dofs-per-entity structure with little fill
    int* browptr = nullptr;
    double* blocks = nullptr;
    int* rowmap = nullptr;         // stored block row -> output block row (locality ordering), or null
    int nbr = 0;
    int max_bpr = 1;               // most blocks in any block row
    bool diag_first = false;       // every block row starts with its diagonal block
    int64_t nblocks = 0;
    std::vector<int> h_browptr;    // host copy of browptr (stored row order): shared-memory budgeting of the persistent kernels
    std::vector<int> h_rowmap;     // host copy of rowmap (empty: identity)
    std::vector<int> h_rowptr;     // host copy of rowptr
    std::vector<double> h_blocks;  // host copy of the blocks (square matrices only: S)
    bool bsr() const { return blocks != nullptr; }
    double bytes() const {
        return bsr() ? 80.0 * (double)nblocks + 4.0 * (double)(nbr + 1) : 12.0 * (double)nnz + 4.0 * (double)(nrows + 1);
    }
};

// kernel classes for statistics / timing (profiles/ and bench.py's roofline use these)
enum KClass { KC_SPMM_K = 0, KC_CG_INIT, KC_SPMM_CG, KC_CG_XR, KC_CG_P, KC_CHEB, KC_MISC, KC_COUNT };

struct ClassStat {
    double ms = 0.0;
    int64_t launches = 0;
    double bytes = 0.0;
};

struct ChebTable {
    int ncheb = -1;
    double L = 0.0;
    std::vector<double> h_coef;    // a_n, n = 0..ncheb (panel plans and re-expansions are built from it on the host)
    std::vector<double> h_relax;   // per-term tolerance relaxation factor of the mass solve (>= 1)
    double t_valid = 0.0;          // largest |t| the series is valid for (L / spectral radius); 0 = unknown
};

// What one panel of columns runs: number of terms, per-column operator scale and series coefficients
// (reference series, or its per-column economised re-expansion), per-term tolerance relaxation.
struct PanelPlan {
    int M = 0;                          // highest term index = number of operator applications
    double tscale[64];                  // per column: B_k = tscale_k * M^-1 K
    std::vector<double> coef;           // (M+1) x kMaxW, row m = coefficient of S_m per column
    std::vector<double> relax;          // M+1
    unsigned long long sig = 0;         // identifies (M, relax) -- what a captured graph depends on
};

struct GraphKey {
    int id, W;
    unsigned long long sig;
    bool operator<(const GraphKey& o) const {
        if (id != o.id) return id < o.id;
        if (W != o.W) return W < o.W;
        return sig < o.sig;
    }
};

struct KrylovState;               // kexp_krylov.cu

// Device arrays of the per-CTA local problems of S (persistent narrow-panel kernel)
struct NarrowPlanDev {
    double* lblocks = nullptr; int* lrowptr = nullptr; int* cta_blk_ptr = nullptr; int* cta_row_ptr = nullptr;
    int* halo_ptr = nullptr; int* halo_cols = nullptr; int* srowptr = nullptr; int* cta_send_ptr = nullptr; int* send_dst = nullptr;
    int* grow = nullptr; int* cta_nb = nullptr;    // local row order of every CTA (rows that reference halo slots first)
    // K~ for the product of the persistent kernel: per CTA the rows sorted by block count, in groups of 32 rows whose blocks
    // are stored transposed ([block slot][16-byte part][row of the group]) so that a warp's loads are contiguous
    double* kt2 = nullptr; int* kt2_cnt = nullptr; int* kt2_orow = nullptr; int* kt2_gbase = nullptr; int* kt2_grp_ptr = nullptr;
    int* kt2_cta_ptr = nullptr;    // rows of K~ per CTA, cut by block count
    int cap_kt_groups = 0, cap_kt_rows = 0;
    void* inbox = nullptr;         // flagged halo inboxes (two buffers)
    int G = 0;                     // grid the plan was built for (< 0: built and found not applicable)
    int cap_blocks = 0, cap_rows = 0, cap_halo = 0, cap_send = 0, inbox_slots = 0;
};

// Everything one in-flight panel needs: its stream, work panels, CG scalars, reduction scratch and
// captured graphs.  The context's own fields are the ACTIVE lane; swap_lane() exchanges them with a
// stored set, so the enqueue code is lane-agnostic (host code is single-threaded per context).
struct LaneState {
    bool allocated = false;
    cudaStream_t stream = nullptr;
    int curW = 0;
    double *S0 = nullptr, *S1 = nullptr, *Y = nullptr, *T = nullptr, *X = nullptr, *P = nullptr, *Q = nullptr;
    CgScalars* d_sc = nullptr;
    double* d_partials = nullptr;
    unsigned* d_counter = nullptr;
    double* d_tscale = nullptr;
    double* h_pin = nullptr;
    std::map<GraphKey, cudaGraphExec_t> graphs;
    unsigned long long base_iters = 0, base_solves = 0;
};

}  // namespace kexp

struct kexp_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;

    // operator plan
    int64_t n = 0, n_s = 0, n_b = 0, n_d = 0;
    kexp::DevCsr Kt, As, Rt;
    double* d_Dblk = nullptr;      // n_d/3 blocks or null
    double* d_Pinv = nullptr;      // n_s/3 blocks
    int* d_perm = nullptr;         // user -> internal or null
    bool have_op = false;
    // raw user matrices kept for the plain plan (jacobi rebuild) -- host copies
    std::vector<int32_t> h_As_rowptr, h_As_col;
    std::vector<double> h_As_val;

    // mass solver
    double rtol = 1e-12;
    int max_it = 60;
    int precond = 3;
    int solver = 0;                // 0: PCG, 1: Chebyshev semi-iteration (kexp_set_mass_solver_chebyshev)
    double cheb_lo = 0.0, cheb_hi = 0.0;
    int graph_its = 60;            // CG iterations captured per solve (trimmed adaptively, see kexp_cheby_apply_dev)
    bool its_tuned = false;
    bool adaptive_its = true;
    int unconverged_seen = 0;
    // deferred convergence check (kexp_set_async): the counters of an apply are copied to pinned memory behind it
    // and looked at by the next call that finds the copy finished, at the latest by kexp_synchronize / download
    // persistent narrow-panel apply kernel (kexp_narrow.cu)
    bool use_narrow = true;        // KEXP_NARROW=0: always the multi-launch path
    int narrow_max_w = 4;          // widest panel that takes it (KEXP_NARROW_MAXW, up to 8).  8-column panels: 5.5 ms against 4.2 ms for the
                                   // multi-launch path at ref 5 (the K~ product of the persistent kernel is the slow part); 4 columns: a tie
    bool narrow_mreg = true;       // panels of up to 4 columns keep their rows of S in registers (KEXP_NARROW_MREG=0: shared memory)
    int narrow_slot = 0;           // constant-memory argument slot of this context
    kexp::NarrowPlanDev narrow_plan;      // per-CTA local problems of S, halos and send lists (kexp_narrow.cu: build_narrow_plan)
    int* d_nktcols = nullptr;      // block columns of K~
    int* d_nk = nullptr;           // per-step iteration counts
    double* d_nrtol2 = nullptr;    // per-step squared tolerances
    double* d_npartials = nullptr; // per-CTA block sums
    int solve_path = -1;           // how the most recently enqueued / captured mass solve runs (kexp_stats_t.solve_path)
    bool async_checks = false;
    kexp::CgScalars* h_sc = nullptr;       // pinned
    cudaEvent_t ev_chk = nullptr;
    bool chk_pending = false;
    int spmm_blocks_per_sm = 0;    // 0: from occupancy
    bool use_relax = true;         // relax the mass-solve tolerance of late Chebyshev terms (KEXP_RELAX=0: uniform rtol)
    bool use_fused_update = false; // KEXP_FUSED_UPDATE=1: recurrence update of the solved rows inside the last mass-solve iteration and of the
                                   // depth rows in the K~ epilogue.  Measured neutral to slightly slower at M1 (profiles/r02_summary.md): off.
    bool depth_rows_fused = false; // set by enqueue_operator: the K~ product updated the rows that need no solve in its epilogue
    bool use_coop = true;          // whole Chebyshev mass solve as one cooperative kernel (KEXP_COOP=0: one kernel per iteration)
    int red_blocks_per_sm = 4;     // grid of the reduction kernels

    kexp::ChebTable cheb[KEXP_MAX_CHEBY];

    // work panels (width curW), internal order
    int curW = 0;
    double *S0 = nullptr, *S1 = nullptr, *Y = nullptr, *T = nullptr, *X = nullptr, *P = nullptr, *Q = nullptr;
    kexp::CgScalars* d_sc = nullptr;
    double* d_partials = nullptr;
    unsigned* d_counter = nullptr;
    double* d_tscale = nullptr;    // kMaxW doubles: t_k / L
    double* d_w = nullptr;         // kMaxW doubles: averaging weights
    double* h_pin = nullptr;       // pinned staging (kMaxW*2 doubles)
    double* d_stage = nullptr;     // device staging for host-layout vectors
    size_t stage_bytes = 0;
    double* h_stage = nullptr;     // pinned host staging
    size_t h_stage_bytes = 0;
    double* d_v = nullptr;         // n doubles (average result)
    double* d_basis = nullptr;     // n x basis_ld: U_m of the shared-input path
    int basis_ld = 0;
    double* d_D = nullptr;         // (ncheb+1) x 64 combination coefficients
    size_t d_D_cap = 0;
    bool use_shared_basis = true;  // KEXP_SHARED_BASIS=0: treat shared-input batches as 64 independent columns
    std::map<unsigned long long, std::vector<kexp::PanelPlan>> plan_cache;   // per (table, layout, sample times)
    bool use_econ = true;          // KEXP_ECONOMISE=0: always run the full reference series (no per-column re-expansion)
    double econ_eps = 1e-13;       // dropped tail of a re-expanded series, relative to sum |b_m|   (KEXP_ECON_EPS)
    double econ_margin = 0.01;     // safety on the spectral interval: s_k = |t_k| (1 + margin) / t_valid
    bool capture_basis = false;    // enqueue_cheby stores every S_m (column 0) into d_basis

    // graphs keyed by (cheby id, W)
    std::map<kexp::GraphKey, cudaGraphExec_t> graphs;
    bool use_graph = true;
    bool use_pdl = true;           // programmatic dependent launch between the kernels of the graph (KEXP_PDL=0 disables): +3 %

    // timing / stats
    bool timing = false;
    kexp::ClassStat cls[kexp::KC_COUNT];
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_pending;   // (class, index of start event)
    size_t ev_used = 0;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    struct kexp_batch* tmp_in = nullptr;
    struct kexp_batch* tmp_out = nullptr;
    double last_ms = 0.0;
    int64_t st_op_apps = 0, st_launches = 0, st_bicg = 0;
    int64_t st_h2d_bytes = 0, st_d2h_bytes = 0;    // state-vector bytes copied between host and device since reset
    unsigned long long base_iters = 0, base_solves = 0;
    std::map<const void*, int> occ_cache;

    // lanes: two half-width panels of one batch can be in flight on two streams, so that the
    // latency-bound CG product of one overlaps the bandwidth-bound vector update of the other
    kexp::LaneState lane_store[2];
    int cur_lane = 0;
    int nlanes = 1;
    int panel_cap = 64;            // widest panel a new batch gets (32 with two lanes)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_prev = nullptr;
    cudaStream_t copy_stream = nullptr;   // host->device uploads overlap the compute of the previous panel
    double* d_stage_up = nullptr;         // landing area of host->device uploads (copy stream only; d_stage belongs to the compute stream)
    size_t stage_up_bytes = 0;
    // asynchronous downloads (kexp_batch_download_async): own staging buffer and stream
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t ev_unpacked = nullptr, ev_d2h_done = nullptr;
    double* d_stage_out = nullptr;
    size_t stage_out_bytes = 0;
    bool d2h_pending = false;

    // krylov
    kexp::KrylovState* kry = nullptr;

    // nccl (dlopen'ed)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
};

struct kexp_batch {
    kexp_ctx* ctx = nullptr;
    int B = 0;
    std::vector<int> panelW;       // width of each panel
    std::vector<double*> panel;    // [n][W] device buffers, internal order
    int stride = 64;               // first column of panel p is p*stride
    double* shared_vec = nullptr;  // [n][2] panel holding the one input of a shared-input batch (column 1 = 0)
    bool is_shared = false;
    std::vector<cudaEvent_t> ready;   // per panel: upload + pack finished (copy stream)
    bool pending = false;          // an asynchronous upload may still be in flight
    cudaEvent_t ev_consumed = nullptr;   // scratch input batch of the host-pointer entry points: the last apply has read the panels
    // stored column j holds the caller's column order[j] (empty = identity).  Sorting the columns by |t_k|
    // groups sample times of similar economised series length into the same panel.
    std::vector<int> order;
    int user_col(int j) const { return order.empty() ? j : order[j]; }
};

namespace kexp {
int fail(kexp_ctx* ctx, const std::string& msg);
int upload_csr(kexp_ctx* ctx, DevCsr& d, int nrows, int ncols, const int32_t* rp, const int32_t* ci, const double* v, bool force_bsr = false);
void free_csr(DevCsr& d);
int pick_W(int B);
void krylov_free(kexp_ctx* ctx);
// kexp_narrow.cu
bool narrow_eligible(kexp_ctx* ctx, int W);
void free_narrow_plan(NarrowPlanDev& P);
int launch_narrow_apply(kexp_ctx* ctx, int W, int M, int mode, const int* kterm, const double* rtol2, bool with_basis, const double** result);
}  // namespace kexp

#define KEXP_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return kexp::fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)
