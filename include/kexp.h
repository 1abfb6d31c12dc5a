/*
 * kexp.h -- C ABI of libkexp.so: B200 (sm_100a) implementation of the exp(sL)v hot path
 * of colinjcotter/krylov_exp.  Plain pointers and sizes only; every function returns 0 on
 * success, non-zero on error (message via kexp_last_error).  One host thread per context.
 * The caller owns all host arrays; the library owns device memory until kexp_destroy.
 *
 * Reference interfaces replaced (file:line in /root/reference):
 *   operator_solver.solve()  == one application of M^-1 K
 *        forms  averaged_sw_explicit.py:142-148, solver :152-164 (Firedrake/PETSc fieldsplit CG)
 *   cheby_exp.apply(x, y, t)                         cheby_exp.py:110-165
 *   krylov_exp.apply(x, y, t, its)                   krylov_exp.py:50-109
 *   DU *= wt; ensemble.allreduce(DU, V)              timestepping_methods.py:10-11
 * Host-side coefficient generation (cheby_exp.py:31-96) stays in the Python mirror
 * (krylov_exp_b200/cheby.py); the table is handed over with kexp_set_cheby.
 *
 * Vector layout at this boundary: a state vector is one contiguous array of
 * n = n_u + n_eta doubles, [BDM2 velocity dofs ; DG1 depth dofs] -- the order of
 * Firedrake's mixed Function (f.subfunctions[0].dat.data, f.subfunctions[1].dat.data).
 * A batch of B vectors is B such arrays, vector k starting at X + k*ldx.
 */
#ifndef KEXP_H
#define KEXP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kexp_ctx kexp_ctx;

typedef struct kexp_stats_t {
    int64_t n;                 /* state size */
    int64_t n_solve;           /* rows of the mass system actually solved by CG */
    int64_t cg_solves;         /* batched mass solves since last reset */
    int64_t cg_iterations;     /* CG iterations executed since last reset (all columns advance together) */
    int64_t op_applications;   /* batched applications of M^-1 K since last reset */
    int64_t kernel_launches;   /* kernels of this library launched since last reset */
    int64_t bicg_iterations;   /* shifted-solve iterations (krylov path) */
    double  last_max_rel_residual; /* max over columns of ||r||/||b|| of the last mass solve */
    double  model_bytes;       /* algorithmic DRAM bytes of the work since last reset (DESIGN.md model) */
    double  spmm_cg_bytes;     /* of which: bytes of the CG-loop SpMM launches */
    int64_t spmm_cg_launches;  /* number of CG-loop SpMM launches since last reset */
    int64_t unconverged_solves;    /* mass solves that missed their tolerance since the context was created (each one made
                                      its apply repeat itself with more iterations, or fail at max_it) */
    double  worst_rel_residual;    /* largest ||r||/||b|| of any such solve */
    int64_t mass_iterations_per_solve; /* current fixed length of a Chebyshev mass solve / cap of a PCG solve */
    int64_t h2d_vector_bytes;      /* state-vector bytes copied host -> device / device -> host by the batch and host-pointer */
    int64_t d2h_vector_bytes;      /* entry points since last reset (a device-resident time step keeps both at 0) */
    int64_t solve_path;            /* kernel path of the most recently enqueued (or graph-captured) mass solve: 0 batched PCG,
                                      1 Chebyshev, one kernel per iteration (bsr3_cheb_iter_kernel), 2 Chebyshev, one cooperative
                                      kernel per solve (cheb_solve_coop_kernel), 3 persistent narrow-panel apply kernel */
} kexp_stats_t;

/* ---- lifetime -------------------------------------------------------------------- */
int  kexp_create(kexp_ctx** out, int device);
int  kexp_destroy(kexp_ctx* ctx);
const char* kexp_last_error(kexp_ctx* ctx);
const char* kexp_version(void);

/* ---- operator: K and M exported once from the host assembler ----------------------
 * replaces the construction at averaged_sw_explicit.py:142-164 (L_eigs.py:44-45 is the
 * reference's own precedent for pulling K and M out as AIJ/CSR matrices).
 * K: n x n CSR; Mu: n_u x n_u SPD CSR; Meta_inv_blocks: n_cells dense 3x3 inverses of the
 * DG1 mass blocks (row-major).  n_u must be a multiple of 3 (BDM2: 3 dofs per entity).
 */
int kexp_set_operator(kexp_ctx* ctx, int64_t n_u, int64_t n_eta,
                      const int32_t* K_rowptr, const int32_t* K_col, const double* K_val,
                      const int32_t* Mu_rowptr, const int32_t* Mu_col, const double* Mu_val,
                      const double* Meta_inv_blocks, int64_t n_cells);

/* Pre-processed ("statically condensed") operator plan, same mathematical operator:
 *   internal order = perm[user index];  rows [0,n_s) solved by CG with As,
 *   rows [n_s, n_s+n_b): W = T + Rt * W_s,  rows [n_s+n_b, n): W = Dblk * T (or T if NULL),
 *   T = Kt * S.   Built by krylov_exp_b200.operator.condense() from K, Mu, Meta.
 * Pinv_blocks: n_s/3 dense 3x3 blocks of the block-Jacobi preconditioner (row-major).
 */
int kexp_set_operator_plan(kexp_ctx* ctx, int64_t n, int64_t n_s, int64_t n_b,
                           const int32_t* Kt_rowptr, const int32_t* Kt_col, const double* Kt_val,
                           const int32_t* As_rowptr, const int32_t* As_col, const double* As_val,
                           const int32_t* Rt_rowptr, const int32_t* Rt_col, const double* Rt_val,
                           const double* Dblk, const double* Pinv_blocks,
                           const int32_t* perm);

/* fieldsplit_0 CG of averaged_sw_explicit.py:155-157: relative residual tolerance,
 * iteration cap (the whole apply is one CUDA graph with max_it early-exit iterations),
 * precond: 1 = point Jacobi, 3 = 3x3 block Jacobi.  A solve that misses rtol within max_it makes the calling apply
 * FAIL ("mass solve did not reach rtol"); rtol below 2e-14 is clamped to 2e-14 (attainable accuracy in fp64). */
int kexp_set_mass_solver(kexp_ctx* ctx, double rtol, int max_it, int precond);
/* Same solve by Chebyshev semi-iteration on the block-Jacobi preconditioned system: spectrum of
 * P^-1 S enclosed by [lmin, lmax] (host estimate); no inner products, one fused kernel per
 * iteration; the residual of every solve is still measured against rtol. */
int kexp_set_mass_solver_chebyshev(kexp_ctx* ctx, double rtol, double lmin, double lmax, int max_it);

/* ---- Chebyshev:  cheby_exp.__init__ hands over a_n = Re(c_n (-i)^n), n = 0..ncheb,
 * and L (cheby_exp.py:51-98).  id selects one of KEXP_MAX_CHEBY tables (cheby, cheby2). */
#define KEXP_MAX_CHEBY 8
int kexp_set_cheby(kexp_ctx* ctx, int id, const double* a, int ncheb, double L);
/* Optional: the largest |t| the table is meant for, t_valid = L / rho(M^-1 K) -- the drivers build L exactly
 * so, L = eigs[ref]*dt*rho (averaged_sw_explicit.py:77-82,168), i.e. t_valid = dt*rho, while the sample times
 * only reach |t_k| <= dt*rho/2 (:91-95).  With it every column evaluates the SAME polynomial
 * sum_n a_n S_n((t_k/L) A) through its exact re-expansion on the part of the Chebyshev interval the spectrum
 * of (t_k/L) A actually fills, which needs fewer terms (dropped tail <= 1e-13 sum|b_m|).  A call with any
 * |t_k| * 1.01 > t_valid, t_valid = 0, or KEXP_ECONOMISE=0 runs the full series. */
int kexp_set_cheby_time_range(kexp_ctx* ctx, int id, double t_valid);

/* cheby_exp.apply(x, y, t)  (cheby_exp.py:110): y = sum_n a_n S_n, x != y, x unchanged. */
int kexp_cheby_apply(kexp_ctx* ctx, int id, const double* x, double* y, double t);

/* B sample times at once (one ensemble member per column, averaged_sw_explicit.py:233-235).
 * x_is_shared != 0: every column starts from the same vector X (timestepping_methods.py:7). */
int kexp_cheby_apply_batch(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                           double* Y, int64_t ldy, const double* t, int B);

/* v = sum_k w_k exp(t_k L) X_k   (DU *= wt; ensemble.allreduce, timestepping_methods.py:9-11).
 * do_allreduce != 0: sum v over the ranks of kexp_comm_init (NCCL) before returning. */
int kexp_cheby_apply_average(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                             const double* t, const double* w, int B, double* v_out, int do_allreduce);

/* The same weighted sum WITHOUT forming the B individual results: every column is a polynomial in the
 * same operator, so  sum_k w_k exp(t_k L) X_k = sum_m S_m(C) z_m  with  z_m = sum_k w_k b_m^(k) X_k,  one
 * dense combination of the inputs followed by ONE Clenshaw recurrence on a single column
 * (timestepping_methods.py:9-11 only ever uses the sum).  Same result to round-off. */
int kexp_cheby_apply_average_fused(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                                   const double* t, const double* w, int B, double* v_out, int do_allreduce);

/* Host-only (no device needed): b = E(s)^T a with  sum_n a_n S_n(s z) = sum_m b_m S_m(z),  |s| <= 1, S_0 = 1, S_1 = z,
 * S_n = 2 z S_{n-1} + S_{n-2} (the real form of cheby_exp.py:110-165), and the number of leading b_m to keep for a
 * dropped tail <= eps * sum|b_m|.  This is the coefficient algebra behind kexp_set_cheby_time_range, the shared-input
 * batches and kexp_cheby_apply_average_fused.  b_out: ncheb+1 doubles. */
int kexp_series_reexpand(const double* a, int ncheb, double s, double eps, double* b_out, int* keep_out);

/* ---- device-resident batches (the hot loop without host<->device copies) ---------- */
typedef struct kexp_batch kexp_batch;
int kexp_batch_create(kexp_ctx* ctx, int B, kexp_batch** out);
int kexp_batch_destroy(kexp_batch* b);
int kexp_batch_num_panels(kexp_batch* b);     /* a batch is stored as panels of <= 64 (or 32) sample columns */
/* Column order inside the batch: stored column j holds the caller's column order[j] (a permutation of
 * 0..B-1; NULL = identity).  Upload, download, the sample times t and the weights w of the calls below
 * always use the CALLER's order.  Sorting by |t_k| puts sample times of similar series length into one
 * panel; the host-array entry points above do so themselves.  Set the order BEFORE uploading. */
int kexp_batch_set_order(kexp_batch* b, const int32_t* order);
int kexp_batch_upload(kexp_batch* b, const double* X, int64_t ldx, int x_is_shared);   /* host -> HBM */
int kexp_batch_download(kexp_batch* b, double* Y, int64_t ldy);                        /* HBM -> host */
/* The same with DEVICE arrays in the caller's layout (B vectors of n doubles, user dof order, vector k at d + k*ld; e.g. a
 * torch tensor): enqueued on the library's stream, no host copy, no synchronisation.  Together with kexp_get_stream this
 * keeps a whole averaged time step (timestepping_methods.py:34-77) in HBM: the nonlinear slow right-hand side
 * (SlowSolver.solve(), averaged_sw_explicit.py:177-213) runs as the caller's own device code between
 * kexp_batch_export_dev and kexp_batch_upload_dev. */
int kexp_batch_upload_dev(kexp_batch* b, const double* d_X, int64_t ldx, int x_is_shared);
int kexp_batch_export_dev(kexp_batch* b, double* d_Y, int64_t ldy);
int kexp_get_stream(kexp_ctx* ctx, void** cuda_stream_out);
/* Asynchronous forms: the device-to-host copy is left in flight on its own stream and overlaps whatever is enqueued
 * next (the backward half of an averaged evaluation, its uploads); Y is complete after kexp_wait_transfers and must
 * stay valid until then (pinned memory for a truly asynchronous copy). */
int kexp_batch_download_async(kexp_batch* b, double* Y, int64_t ldy);
int kexp_cheby_apply_batch_async(kexp_ctx* ctx, int id, const double* X, int64_t ldx, int x_is_shared,
                                 double* Y, int64_t ldy, const double* t, int B);
int kexp_wait_transfers(kexp_ctx* ctx);
int kexp_cheby_apply_dev(kexp_ctx* ctx, int id, kexp_batch* in, kexp_batch* out, const double* t);
/* device-resident form of kexp_cheby_apply_average_fused; v_dev as in kexp_batch_average_dev */
int kexp_cheby_average_fused_dev(kexp_ctx* ctx, int id, kexp_batch* in, const double* t, const double* w, double* v_dev);
/* v_dev: device pointer to n doubles (e.g. a torch tensor's data_ptr), user dof order */
int kexp_batch_average_dev(kexp_batch* b, const double* w, double* v_dev);
int kexp_synchronize(kexp_ctx* ctx);
/* on != 0: kexp_cheby_apply_dev / kexp_cheby_average_fused_dev return as soon as their work is enqueued (no host
 * synchronisation once the iteration count is tuned), so applies queue back to back and overlap with uploads.  The
 * convergence record of the mass solves is then inspected by a later call; a solve that missed rtol is reported as an
 * error by that call, at the latest by kexp_synchronize / kexp_batch_download / kexp_stats.  Default off: every apply
 * checks (and if needed repeats) itself before returning. */
int kexp_set_async(kexp_ctx* ctx, int on);
/* ms of device time between the start and end of the last *_dev / apply call (CUDA events) */
int kexp_last_device_ms(kexp_ctx* ctx, double* ms);
/* CUDA events on the library's stream around an arbitrary sequence of calls */
int kexp_timer_start(kexp_ctx* ctx);
int kexp_timer_stop(kexp_ctx* ctx, double* ms);
/* CUDA-event time accumulated over the CG-loop SpMM launches since reset; only collected
 * when timing mode is on (it disables graph replay and adds two event records per launch) */
int kexp_set_kernel_timing(kexp_ctx* ctx, int on);
int kexp_get_spmm_cg_ms(kexp_ctx* ctx, double* ms);

/* per kernel class (0 spmm_K, 1 cg_init, 2 spmm_cg, 3 cg_update_xr, 4 cg_update_p, 5 cheb_update,
 * 6 misc): event time and launches since reset (timing mode), algorithmic vector bytes per
 * launch and sample column, and matrix bytes per launch (DESIGN.md byte model) */
int kexp_get_class_stats(kexp_ctx* ctx, int cls, double* ms, int64_t* launches, double* bytes_per_launch_per_col);
int kexp_get_class_matrix_bytes(kexp_ctx* ctx, int cls, double* bytes);

/* one batched application of M^-1 K (operator_solver.solve(), cheby_exp.py:128) */
int kexp_operator_apply_dev(kexp_ctx* ctx, kexp_batch* in, kexp_batch* out);
/* Y = K X only (the SpMM building block, for the SpMM-only sweep of SURVEY 8d) */
int kexp_spmm_K_dev(kexp_ctx* ctx, kexp_batch* in, kexp_batch* out);

/* ---- shift-and-invert Arnoldi:  krylov_exp.apply  (krylov_exp.py:50-109) ----------
 * Shifted system (M + gamma K) w = M v  (test_krylov.py:34-62) with eta eliminated:
 *   Ash u = Brhs v,  eta = v_eta + Eback u.   Mfull (n x n CSR) gives the M inner product
 *   assemble(inner(a,b)*dx) of krylov_exp.py:55-56,74.  Built by the Python mirror. */
int kexp_set_krylov_operator(kexp_ctx* ctx, double gamma,
                             const int32_t* Ash_rowptr, const int32_t* Ash_col, const double* Ash_val,
                             const int32_t* Brhs_rowptr, const int32_t* Brhs_col, const double* Brhs_val,
                             const int32_t* Eback_rowptr, const int32_t* Eback_col, const double* Eback_val,
                             const int32_t* M_rowptr, const int32_t* M_col, const double* M_val,
                             double solve_rtol, int solve_max_it);
/* H_out: (kdim+1) x kdim row-major; u_out: its doubles; residuals_out: its doubles
 * (the "Residual" prints of krylov_exp.py:95-101) or NULL to skip that diagnostic. */
int kexp_krylov_apply(kexp_ctx* ctx, const double* x, double* y, double t, int kdim, int its,
                      double* H_out, double* u_out, double* residuals_out);
/* B sample times at once.  x_is_shared != 0: the shift-and-invert Krylov space does not depend on t, so ONE Arnoldi
 * process serves every t_k (only expm(-t_k H1) e1 and the final combination differ); distinct inputs run their Arnoldi
 * processes back to back on the device queue.  The Arnoldi loop itself never waits for the host: ||w||_M, H and the
 * normalisation scalars stay on the device, breakdown / non-convergence are recorded there and reported at the end. */
int kexp_krylov_apply_batch(kexp_ctx* ctx, const double* X, int64_t ldx, int x_is_shared, double* Y, int64_t ldy,
                            const double* t, int B, int kdim, int its);
/* device small dense expm (Pade-13 scaling and squaring), exposed for tests:
 * E = expm(A), A m x m row-major, m <= 64 */
int kexp_dense_expm(kexp_ctx* ctx, const double* A, int m, double* E);

/* ---- nonlinear slow right-hand side (the step between the two exponentials of an averaged evaluation) -------------
 * L(U) of /root/reference/averaged_sw_explicit.py:177-213 (vector-invariant, upwinded) for B ensemble members at once:
 * U, L are DEVICE arrays of B vectors of n = n_u + n_eta doubles in the user dof order (vector k at U + k*ldu), the
 * tables are the device copies of krylov_exp_b200.slow_rhs.SlowTables (per-cell geometry factors at the quadrature
 * points, basis tabulations, facet neighbours; host-side input generation like the assembled K and M).  Enqueued on
 * `cuda_stream`; SlowSolver.solve() = M^-1 L is this followed by kexp_operator_apply_dev of a context whose K is the
 * identity.  nq <= 36 cell points, nf <= 6 facet points. */
int kexp_slow_form_dev(void* cuda_stream, int n_cells, int n_u, int n_eta, int nq, int nf,
                       const double* cell, const double* fac, const double* Phi, const double* dPhi, const double* gPhi,
                       const double* Psi, const double* gPsi, const double* PhiF, const double* PsiF, const double* qw,
                       const double* sign, const int32_t* dofs, const int32_t* nbr_cell, const int32_t* nbr_edge, const double* b,
                       const double* U, int64_t ldu, double* L, int64_t ldl, int B, double dT);

/* ---- multi-GPU: one context per GPU/process; unique id from ncclGetUniqueId -------- */
int kexp_comm_unique_id(void* id128);                       /* 128-byte buffer */
int kexp_comm_init(kexp_ctx* ctx, const void* id128, int rank, int nranks);
int kexp_allreduce_sum_dev(kexp_ctx* ctx, double* v_dev, int64_t count);

/* ---- statistics --------------------------------------------------------------------- */
int kexp_stats(kexp_ctx* ctx, kexp_stats_t* out);
int kexp_stats_reset(kexp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KEXP_H */
