// kexp_krylov.cu -- shift-and-invert Arnoldi (krylov_exp.py:50-109) on the device.
//
//   beta = ||x||_M, V_0 = x/beta                                     krylov_exp.py:58-64
//   for j < its:  w = (M + gamma K)^-1 M V_j                         :70-71  (BiCGStab on the eta-eliminated system)
//                 MGS in the M inner product, H(:,j)                  :73-80  (one cooperative kernel)
//                 Hinv, H1 = (Hinv - I)/gamma, expm(-H1 s t) e1       :87-99  (one CTA)
//   y = beta sum_j u_j V_j                                            :102-109
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <type_traits>
#include <vector>
#include "kexp_internal.h"
#include "kexp_krylov_kernels.cuh"

namespace kexp {

struct KrylovState {
    double gamma = 0.0;
    double rtol = 1e-13;
    int max_it = 2000;
    int64_t n = 0, n_u = 0;
    DevCsr Ash, Brhs, Eback, M;
    double* d_Pinv = nullptr;          // 3x3 block-Jacobi of Ash
    // Arnoldi basis
    int kdim = 0;
    double *V = nullptr, *Z = nullptr; // (kdim+1) x n
    double *w = nullptr, *mw = nullptr, *tmp = nullptr;
    // BiCGStab work (n_u)
    double *b = nullptr, *x = nullptr, *r = nullptr, *r0 = nullptr, *p = nullptr, *v = nullptr, *s = nullptr, *t = nullptr, *y = nullptr, *z = nullptr;
    // scalars
    double* d_dots = nullptr;          // 4
    double* h_dots = nullptr;          // pinned 4
    double* d_partials = nullptr;
    unsigned* d_counter = nullptr;
    double* d_H = nullptr;             // (kdim+1) x kdim row-major
    double* d_small = nullptr;         // dense scratch
    double* d_u = nullptr;
    double* d_fac = nullptr;
    double* d_svals = nullptr;
    double* d_scal = nullptr;          // [0] 1/||x||_M, [1] ||x||_M, [2] 1/H[j+1][j], [3] H[j+1][j]   (device scalars of the loop)
    double* d_res = nullptr;           // residual estimates (64)
    double* d_out = nullptr;           // batch results [B][n]
    double* d_ub = nullptr;            // batch coefficient vectors [B][64]
    size_t out_cap = 0;

    int grid = 148;
    int coop_grid = 148;
    // one-kernel BiCGStab (default); KEXP_BICG_HOST=1 keeps the host-driven loop for comparison
    bool bicg_coop = true;
    int bicg_grid = 148;
    double* d_bicg_partials = nullptr;
    int* d_info = nullptr;             // [0] iterations (accumulated), [1] sticky status, [2] last solve
    int* h_info = nullptr;             // pinned copy
    bool info_pending = false;
    // matrix-in-shared-memory variant (bicgstab_smem_kernel): sliced-ELL chunks, one per SM
    bool bicg_smem = false;
    double* sm_vals = nullptr;
    int *sm_cols = nullptr, *sm_chunk_slice = nullptr, *sm_slice_slot = nullptr, *sm_chunk_row = nullptr, *sm_rowmap = nullptr;
    size_t sm_bytes = 0;               // dynamic shared memory per CTA
    int sm_grid = 0;
};

static void free_state(KrylovState* k) {
    if (!k) return;
    free_csr(k->Ash); free_csr(k->Brhs); free_csr(k->Eback); free_csr(k->M);
    double* bufs[] = {k->d_Pinv, k->V, k->Z, k->w, k->mw, k->tmp, k->b, k->x, k->r, k->r0, k->p, k->v, k->s, k->t, k->y, k->z,
                      k->d_dots, k->d_partials, k->d_H, k->d_small, k->d_u, k->d_fac, k->d_svals, k->d_scal, k->d_res, k->d_out, k->d_ub};
    for (double* b : bufs) cudaFree(b);
    cudaFree(k->d_counter);
    cudaFree(k->d_bicg_partials);
    cudaFree(k->d_info);
    cudaFree(k->sm_vals); cudaFree(k->sm_cols); cudaFree(k->sm_chunk_slice); cudaFree(k->sm_slice_slot);
    cudaFree(k->sm_chunk_row); cudaFree(k->sm_rowmap);
    if (k->h_dots) cudaFreeHost(k->h_dots);
    if (k->h_info) cudaFreeHost(k->h_info);
    delete k;
}

void krylov_free(kexp_ctx* ctx) {
    free_state(ctx->kry);
    ctx->kry = nullptr;
}

}  // namespace kexp

using namespace kexp;

namespace {

inline int grid_for(int n, int cap) { return std::max(1, std::min(cap, (n + kKryThreads - 1) / kKryThreads)); }

void spmv(kexp_ctx* ctx, const DevCsr& A, const double* x, double* y) {
    KrylovState* k = ctx->kry;
    const int grid = std::max(1, std::min(8 * k->grid, (A.nrows + 7) / 8));
    spmv_warp_kernel<<<grid, kKryThreads, 0, ctx->stream>>>(A.rowptr, A.col, A.val, x, y, A.nrows);
    ctx->st_launches++;
}

void lincomb(kexp_ctx* ctx, double* out, double a, const double* x, double b, const double* y, double c, const double* z, int n) {
    lincomb3_kernel<<<grid_for(n, 4 * ctx->kry->grid), kKryThreads, 0, ctx->stream>>>(out, a, x, b, y, c, z, n);
    ctx->st_launches++;
}

// up to 4 dots, result on the host (one stream sync)
int dots(kexp_ctx* ctx, int count, const double* const* a, const double* const* b, const int* n, double* out) {
    KrylovState* k = ctx->kry;
    DotArgs args;
    memset(&args, 0, sizeof(args));
    args.count = count;
    for (int i = 0; i < count; ++i) { args.a[i] = a[i]; args.b[i] = b[i]; args.n[i] = n[i]; }
    multidot_kernel<<<k->grid, kKryThreads, 0, ctx->stream>>>(args, k->d_partials, k->d_counter, k->d_dots);
    ctx->st_launches++;
    KEXP_CUDA(cudaMemcpyAsync(k->h_dots, k->d_dots, sizeof(double) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; ++i) out[i] = k->h_dots[i];
    return 0;
}

int dot1(kexp_ctx* ctx, const double* a, const double* b, int n, double* out) {
    const double* aa[1] = {a}; const double* bb[1] = {b}; int nn[1] = {n};
    return dots(ctx, 1, aa, bb, nn, out);
}

// w <- (M + gamma K)^-1 M v : right-preconditioned BiCGStab on the eliminated system
int shifted_solve(kexp_ctx* ctx, const double* vin, double* wout) {
    KrylovState* k = ctx->kry;
    const int nu = (int)k->n_u, n = (int)k->n;
    const int nblk = nu / 3;
    auto precond = [&](const double* in, double* out) {
        blk3_apply_kernel<<<grid_for(nblk, 4 * k->grid), kKryThreads, 0, ctx->stream>>>(k->d_Pinv, in, out, nblk);
        ctx->st_launches++;
    };
    spmv(ctx, k->Brhs, vin, k->b);                                   // b = M_u v_u - gamma K_ue v_eta
    if (k->bicg_smem) {
        BicgSmemArgs a;
        a.vals = k->sm_vals; a.cols = k->sm_cols; a.chunk_slice = k->sm_chunk_slice; a.slice_slot = k->sm_slice_slot;
        a.chunk_row = k->sm_chunk_row; a.rowmap = k->sm_rowmap; a.Pinv = k->d_Pinv; a.b = k->b;
        a.x = k->x; a.y = k->y; a.z = k->z; a.max_it = k->max_it; a.rtol = k->rtol;
        a.partials = k->d_bicg_partials; a.info = k->d_info;
        void* args[] = {&a};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)bicgstab_smem_kernel, dim3(k->sm_grid), dim3(kBicgSmemThreads), args, k->sm_bytes, ctx->stream);
        if (e != cudaSuccess) return kexp::fail(ctx, std::string("cooperative BiCGStab (shared-memory matrix) launch: ") + cudaGetErrorString(e));
        ctx->st_launches++;
        KEXP_CUDA(cudaMemcpyAsync(k->h_info, k->d_info, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        k->info_pending = true;
        KEXP_CUDA(cudaMemcpyAsync(wout, k->x, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
        spmv(ctx, k->Eback, k->x, k->tmp);
        lincomb(ctx, wout + nu, 1.0, vin + nu, 1.0, k->tmp, 0.0, nullptr, n - nu);
        return 0;
    }
    if (k->bicg_coop) {
        BicgArgs a;
        a.rowptr = k->Ash.rowptr; a.col = k->Ash.col; a.val = k->Ash.val; a.Pinv = k->d_Pinv; a.b = k->b;
        a.x = k->x; a.r = k->r; a.r0 = k->r0; a.p = k->p; a.v = k->v; a.s = k->s; a.t = k->t; a.y = k->y; a.z = k->z;
        a.nu = nu; a.max_it = k->max_it; a.rtol = k->rtol; a.partials = k->d_bicg_partials; a.info = k->d_info;
        void* args[] = {&a};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)bicgstab_coop_kernel, dim3(k->bicg_grid), dim3(kKryThreads), args, 0, ctx->stream);
        if (e != cudaSuccess) return kexp::fail(ctx, std::string("cooperative BiCGStab launch: ") + cudaGetErrorString(e));
        ctx->st_launches++;
        KEXP_CUDA(cudaMemcpyAsync(k->h_info, k->d_info, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        k->info_pending = true;
        KEXP_CUDA(cudaMemcpyAsync(wout, k->x, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
        spmv(ctx, k->Eback, k->x, k->tmp);
        lincomb(ctx, wout + nu, 1.0, vin + nu, 1.0, k->tmp, 0.0, nullptr, n - nu);
        return 0;
    }
    KEXP_CUDA(cudaMemsetAsync(k->x, 0, sizeof(double) * nu, ctx->stream));
    KEXP_CUDA(cudaMemcpyAsync(k->r, k->b, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
    KEXP_CUDA(cudaMemcpyAsync(k->r0, k->b, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
    KEXP_CUDA(cudaMemsetAsync(k->p, 0, sizeof(double) * nu, ctx->stream));
    KEXP_CUDA(cudaMemsetAsync(k->v, 0, sizeof(double) * nu, ctx->stream));
    double bb = 0.0;
    if (dot1(ctx, k->b, k->b, nu, &bb)) return 1;
    double rho = 1.0, alpha = 1.0, omega = 1.0, rho_new = bb;         // r0.r = b.b at the start
    const double tol2 = k->rtol * k->rtol * bb;
    if (bb > 0.0) {
        int it = 0;
        for (; it < k->max_it; ++it) {
            if (rho_new == 0.0 || omega == 0.0) return kexp::fail(ctx, "BiCGStab breakdown in the shifted solve");
            const double beta = (rho_new / rho) * (alpha / omega);
            lincomb(ctx, k->p, 1.0, k->r, beta, k->p, -beta * omega, k->v, nu);      // p = r + beta (p - omega v)
            precond(k->p, k->y);
            spmv(ctx, k->Ash, k->y, k->v);
            double r0v = 0.0;
            if (dot1(ctx, k->r0, k->v, nu, &r0v)) return 1;
            alpha = rho_new / r0v;
            lincomb(ctx, k->s, 1.0, k->r, -alpha, k->v, 0.0, nullptr, nu);           // s = r - alpha v
            precond(k->s, k->z);
            spmv(ctx, k->Ash, k->z, k->t);
            double d3[3];
            const double* a3[3] = {k->t, k->t, k->s};
            const double* b3[3] = {k->s, k->t, k->s};
            int n3[3] = {nu, nu, nu};
            if (dots(ctx, 3, a3, b3, n3, d3)) return 1;
            ctx->st_bicg++;
            if (d3[2] <= tol2) {                                                       // s already small enough
                lincomb(ctx, k->x, 1.0, k->x, alpha, k->y, 0.0, nullptr, nu);
                break;
            }
            omega = (d3[1] > 0.0) ? d3[0] / d3[1] : 0.0;
            lincomb(ctx, k->x, 1.0, k->x, alpha, k->y, omega, k->z, nu);             // x += alpha y + omega z
            lincomb(ctx, k->r, 1.0, k->s, -omega, k->t, 0.0, nullptr, nu);           // r = s - omega t
            rho = rho_new;
            double d2[2];
            const double* a2[2] = {k->r0, k->r};
            const double* b2[2] = {k->r, k->r};
            int n2[2] = {nu, nu};
            if (dots(ctx, 2, a2, b2, n2, d2)) return 1;
            rho_new = d2[0];
            if (d2[1] <= tol2) break;
        }
        if (it >= k->max_it) return kexp::fail(ctx, "shifted solve did not converge (raise solve_max_it)");
    }
    // w_u = x ; w_eta = v_eta + Eback w_u
    KEXP_CUDA(cudaMemcpyAsync(wout, k->x, sizeof(double) * nu, cudaMemcpyDeviceToDevice, ctx->stream));
    spmv(ctx, k->Eback, k->x, k->tmp);
    lincomb(ctx, wout + nu, 1.0, vin + nu, 1.0, k->tmp, 0.0, nullptr, n - nu);
    return 0;
}

// Sliced-ELL chunks of the shifted operator for bicgstab_smem_kernel: block rows split into `G`
// contiguous ranges of about equal block count; inside a range rows are sorted by block count and
// cut into slices of 32.  Returns false (and leaves the state untouched) when a chunk does not fit.
bool build_smem_chunks(kexp_ctx* ctx, KrylovState* k, const int32_t* rowptr, const int32_t* col, const double* val,
                       int64_t nu, int G, size_t smem_limit) {
    const int nbr = (int)(nu / 3);
    // 3x3 block structure
    std::vector<int> bptr(nbr + 1, 0);
    std::vector<int> bcol;
    std::vector<double> bval;
    {
        std::vector<int> seen(nbr, -1);
        for (int br = 0; br < nbr; ++br) {
            const size_t first = bcol.size();
            for (int i = 0; i < 3; ++i)
                for (int q = rowptr[3 * br + i]; q < rowptr[3 * br + i + 1]; ++q) {
                    const int bc = col[q] / 3;
                    if (seen[bc] < (int)first) {                  // not yet in this block row
                        seen[bc] = (int)bcol.size();
                        bcol.push_back(bc);
                        bval.insert(bval.end(), 9, 0.0);
                    }
                    bval[(size_t)seen[bc] * 9 + 3 * i + (col[q] % 3)] += val[q];
                }
            bptr[br + 1] = (int)bcol.size();
        }
    }
    const int64_t nblocks = bptr[nbr];
    std::vector<int> chunk_row(G + 1, 0), chunk_slice(G + 1, 0), slice_slot(1, 0), rowmap(nbr);
    {
        int br = 0;
        for (int c = 0; c < G; ++c) {
            const int64_t target = nblocks * (c + 1) / G;
            chunk_row[c] = br;
            while (br < nbr && bptr[br + 1] <= target) ++br;
            if (c == G - 1) br = nbr;
        }
        chunk_row[G] = nbr;
    }
    std::vector<double> vals;
    std::vector<int> cols;
    size_t worst = 0;
    for (int c = 0; c < G; ++c) {
        const int r0 = chunk_row[c], r1 = chunk_row[c + 1];
        if (r1 - r0 > kBicgSmemThreads) return false;
        std::vector<int> order(r1 - r0);
        for (int i = 0; i < r1 - r0; ++i) order[i] = r0 + i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bptr[a + 1] - bptr[a] > bptr[b + 1] - bptr[b]; });
        for (int i = 0; i < r1 - r0; ++i) rowmap[r0 + i] = order[i];
        chunk_slice[c] = (int)slice_slot.size() - 1;
        const int slot_first = slice_slot.back();
        for (int s0 = 0; s0 < r1 - r0; s0 += 32) {
            const int lanes = std::min(32, r1 - r0 - s0);
            int width = 0;
            for (int l = 0; l < lanes; ++l) width = std::max(width, bptr[order[s0 + l] + 1] - bptr[order[s0 + l]]);
            const size_t slot0 = (size_t)slice_slot.back();
            vals.resize((slot0 + width) * 9 * 32, 0.0);
            cols.resize((slot0 + width) * 32, 0);
            for (int l = 0; l < 32; ++l) {
                const int br = (l < lanes) ? order[s0 + l] : order[s0];
                const int cnt = (l < lanes) ? bptr[br + 1] - bptr[br] : 0;
                for (int q = 0; q < width; ++q) {
                    cols[(slot0 + q) * 32 + l] = (q < cnt) ? bcol[bptr[br] + q] : br;      // padding: own column, zero values
                    if (q < cnt)
                        for (int e = 0; e < 9; ++e) vals[((slot0 + q) * 9 + e) * 32 + l] = bval[(size_t)(bptr[br] + q) * 9 + e];
                }
            }
            slice_slot.push_back((int)(slot0 + width));
        }
        const size_t nslots = (size_t)(slice_slot.back() - slot_first);
        worst = std::max(worst, nslots * (9 * 32 * sizeof(double) + 32 * sizeof(int)));
    }
    chunk_slice[G] = (int)slice_slot.size() - 1;
    const size_t bytes = worst + (size_t)9 * kBicgSmemThreads * sizeof(double);
    if (bytes > smem_limit) return false;
    auto up = [&](auto** dst, const auto& v) -> bool {
        using T = std::remove_reference_t<decltype(**dst)>;
        if (cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1)) != cudaSuccess) return true;
        return cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) != cudaSuccess;
    };
    if (up(&k->sm_vals, vals) || up(&k->sm_cols, cols) || up(&k->sm_chunk_slice, chunk_slice) || up(&k->sm_slice_slot, slice_slot) ||
        up(&k->sm_chunk_row, chunk_row) || up(&k->sm_rowmap, rowmap)) return false;
    if (cudaFuncSetAttribute(bicgstab_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bicgstab_smem_kernel, kBicgSmemThreads, bytes);
    if (occ < 1) return false;
    k->sm_bytes = bytes;
    k->sm_grid = G;
    (void)ctx;
    return true;
}

int alloc_d(kexp_ctx* ctx, double** p, size_t count) {
    KEXP_CUDA(cudaMalloc(p, sizeof(double) * std::max<size_t>(count, 1)));
    return 0;
}

}  // namespace

extern "C" {

int kexp_set_krylov_operator(kexp_ctx* ctx, double gamma,
                             const int32_t* Ash_rowptr, const int32_t* Ash_col, const double* Ash_val,
                             const int32_t* Brhs_rowptr, const int32_t* Brhs_col, const double* Brhs_val,
                             const int32_t* Eback_rowptr, const int32_t* Eback_col, const double* Eback_val,
                             const int32_t* M_rowptr, const int32_t* M_col, const double* M_val,
                             double solve_rtol, int solve_max_it) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!ctx->have_op) return kexp::fail(ctx, "set the operator first (kexp_set_operator)");
    if (!Ash_rowptr || !Brhs_rowptr || !Eback_rowptr || !M_rowptr) return kexp::fail(ctx, "null matrix");
    if (!(gamma != 0.0) || !(solve_rtol > 0.0) || solve_max_it < 1) return kexp::fail(ctx, "bad shifted-solve parameters");
    krylov_free(ctx);
    KrylovState* k = new KrylovState();
    ctx->kry = k;
    k->gamma = gamma; k->rtol = solve_rtol; k->max_it = solve_max_it;
    k->n = ctx->n;
    // n_u from the shapes: Ash is n_u x n_u; the caller passes consistent CSR arrays
    int64_t nu = 0;
    {   // Brhs has n_u rows and n columns: find n_u as the largest r with a valid rowptr is not possible
        // without the size, so Eback (n_eta x n_u) and M (n x n) fix it: n_u = n - n_eta, n_eta = 3*cells
        // is not known here either -> n_u is derived from the plan: rows solved + rows back-substituted.
        nu = ctx->n_s + ctx->n_b;
    }
    k->n_u = nu;
    const int64_t ne = k->n - nu;
    if (nu % 3) { krylov_free(ctx); return kexp::fail(ctx, "n_u must be a multiple of 3"); }
    if (upload_csr(ctx, k->Ash, (int)nu, (int)nu, Ash_rowptr, Ash_col, Ash_val) ||
        upload_csr(ctx, k->Brhs, (int)nu, (int)k->n, Brhs_rowptr, Brhs_col, Brhs_val) ||
        upload_csr(ctx, k->Eback, (int)ne, (int)nu, Eback_rowptr, Eback_col, Eback_val) ||
        upload_csr(ctx, k->M, (int)k->n, (int)k->n, M_rowptr, M_col, M_val)) { krylov_free(ctx); return 1; }
    // 3x3 block-Jacobi of Ash
    std::vector<double> Pinv((size_t)nu * 3, 0.0);
    for (int64_t b = 0; b < nu / 3; ++b) {
        double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < 3; ++i)
            for (int q = Ash_rowptr[3 * b + i]; q < Ash_rowptr[3 * b + i + 1]; ++q) {
                const int64_t c = Ash_col[q];
                if (c >= 3 * b && c < 3 * b + 3) A[i][c - 3 * b] += Ash_val[q];
            }
        const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        if (!(fabs(det) > 0.0)) { krylov_free(ctx); return kexp::fail(ctx, "singular diagonal block in the shifted operator"); }
        double* P = &Pinv[(size_t)b * 9];
        P[0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det; P[1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det; P[2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
        P[3] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det; P[4] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det; P[5] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
        P[6] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det; P[7] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det; P[8] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
    }
    bool bad = alloc_d(ctx, &k->d_Pinv, Pinv.size());
    if (!bad) bad = cudaMemcpy(k->d_Pinv, Pinv.data(), sizeof(double) * Pinv.size(), cudaMemcpyHostToDevice) != cudaSuccess;
    double** un[] = {&k->b, &k->x, &k->r, &k->r0, &k->p, &k->v, &k->s, &k->t, &k->y, &k->z};
    for (auto pp : un) bad = bad || alloc_d(ctx, pp, (size_t)nu);
    bad = bad || alloc_d(ctx, &k->w, (size_t)k->n) || alloc_d(ctx, &k->mw, (size_t)k->n) || alloc_d(ctx, &k->tmp, (size_t)k->n);
    k->grid = ctx->num_sms;
    bad = bad || alloc_d(ctx, &k->d_dots, 4) || alloc_d(ctx, &k->d_partials, (size_t)8 * ctx->num_sms * 4 + 64);
    bad = bad || cudaMalloc(&k->d_counter, sizeof(unsigned)) != cudaSuccess || cudaMemset(k->d_counter, 0, sizeof(unsigned)) != cudaSuccess;
    bad = bad || cudaMallocHost(&k->h_dots, sizeof(double) * 4) != cudaSuccess;
    bad = bad || alloc_d(ctx, &k->d_small, 12 * 64 * 64) || alloc_d(ctx, &k->d_u, 64) || alloc_d(ctx, &k->d_fac, 8) || alloc_d(ctx, &k->d_svals, 8);
    if (bad) { krylov_free(ctx); return kexp::fail(ctx, "out of memory in kexp_set_krylov_operator"); }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mgs_sweep_kernel, kKryThreads, 0);
    k->coop_grid = std::max(1, occ) * ctx->num_sms;
    {
        const char* e = getenv("KEXP_BICG_HOST");
        k->bicg_coop = !(e && atoi(e) != 0);
        int per_sm = 4;
        if ((e = getenv("KEXP_BICG_BLOCKS")) && atoi(e) > 0) per_sm = atoi(e);
        occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bicgstab_coop_kernel, kKryThreads, 0);
        k->bicg_grid = std::max(1, std::min(per_sm, occ)) * ctx->num_sms;
        if (alloc_d(ctx, &k->d_bicg_partials, (size_t)k->bicg_grid * 16) ||
            cudaMalloc(&k->d_info, sizeof(int) * 4) != cudaSuccess || cudaMemset(k->d_info, 0, sizeof(int) * 4) != cudaSuccess ||
            cudaMallocHost(&k->h_info, sizeof(int) * 4) != cudaSuccess) { krylov_free(ctx); return kexp::fail(ctx, "out of memory in kexp_set_krylov_operator"); }
        k->h_info[0] = k->h_info[1] = k->h_info[2] = k->h_info[3] = 0;
        // matrix resident in shared memory when it fits (ref <= 5); KEXP_BICG_SMEM=0 keeps it in L2
        e = getenv("KEXP_BICG_SMEM");
        if (k->bicg_coop && !(e && atoi(e) == 0)) {
            int optin = 0;
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
            const size_t limit = (size_t)std::max(0, optin - 1024);             // static shared memory of the kernel
            k->bicg_smem = build_smem_chunks(ctx, k, Ash_rowptr, Ash_col, Ash_val, nu, ctx->num_sms, limit);
            if (k->bicg_smem && k->sm_grid > k->bicg_grid) {                     // partials sized for bicg_grid CTAs
                cudaFree(k->d_bicg_partials); k->d_bicg_partials = nullptr;
                if (alloc_d(ctx, &k->d_bicg_partials, (size_t)k->sm_grid * 16)) { krylov_free(ctx); return 1; }
            }
        }
    }
    return 0;
}

int kexp_dense_expm(kexp_ctx* ctx, const double* A, int m, double* E) {
    if (!ctx || !A || !E) return 1;
    cudaSetDevice(ctx->device);
    if (m < 1 || m > 64) return kexp::fail(ctx, "dense expm supports 1 <= m <= 64");
    double* d = nullptr;
    KEXP_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)m * m * 10));
    KEXP_CUDA(cudaMemcpyAsync(d, A, sizeof(double) * m * m, cudaMemcpyHostToDevice, ctx->stream));
    dense_expm_kernel<<<1, kDenseThreads, 0, ctx->stream>>>(d, d + m * m, d + 2 * m * m, m);
    ctx->st_launches++;
    cudaError_t e = cudaMemcpyAsync(E, d + m * m, sizeof(double) * m * m, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return kexp::fail(ctx, cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Arnoldi process of krylov_exp.py:58-80 for ONE start vector that is already on the device (k->w), entirely on
// the library's stream: H (device), V, Z = M V and ||x||_M (k->d_scal[1]) are left in the state; nothing here
// waits for the device except the optional residual diagnostic (its operator application checks its mass solve).
// ---------------------------------------------------------------------------------------------
static int arnoldi_dev(kexp_ctx* ctx, int kdim, int its, double t, double* d_res /* device, its doubles, or null */) {
    KrylovState* k = ctx->kry;
    const int n = (int)k->n;
    const int ldh = kdim;
    KEXP_CUDA(cudaMemsetAsync(k->d_H, 0, sizeof(double) * (size_t)(kdim + 1) * kdim, ctx->stream));
    // beta = ||x||_M, V0 = x / beta, Z0 = M V0
    spmv(ctx, k->M, k->w, k->mw);
    {
        DotArgs args;
        memset(&args, 0, sizeof(args));
        args.count = 1; args.a[0] = k->w; args.b[0] = k->mw; args.n[0] = n;
        multidot_kernel<<<k->grid, kKryThreads, 0, ctx->stream>>>(args, k->d_partials, k->d_counter, k->d_dots);
        arnoldi_norm_kernel<<<1, 1, 0, ctx->stream>>>(k->d_dots, nullptr, k->d_scal, nullptr, nullptr);
        scale2_kernel<<<grid_for(n, 4 * k->grid), kKryThreads, 0, ctx->stream>>>(k->V, k->Z, k->w, k->mw, k->d_scal, n);
        ctx->st_launches += 3;
    }
    const double svals[3] = {1.0 / 3.0, 2.0 / 3.0, 1.0};            // krylov_exp.py:90
    KEXP_CUDA(cudaMemcpyAsync(k->d_svals, svals, sizeof(svals), cudaMemcpyHostToDevice, ctx->stream));
    kexp_batch *bi = nullptr, *bo = nullptr;
    if (d_res) {
        if (kexp_batch_create(ctx, 1, &bi) || kexp_batch_create(ctx, 1, &bo)) return 1;
    }
    int rc = 0;
    for (int j = 0; j < its && !rc; ++j) {
        const double* Vj = k->V + (size_t)j * n;
        if (shifted_solve(ctx, Vj, k->w)) { rc = 1; break; }
        {   // modified Gram-Schmidt sweep (cooperative launch: grid-wide barriers inside)
            const int slice = kKryThreads * kMgsPerThread;
            int grid = (n + slice - 1) / slice;
            if (grid > k->coop_grid) { rc = kexp::fail(ctx, "vector too long for the cooperative MGS sweep"); break; }
            double* wv = k->w; const double* Vb = k->V; const double* Zb = k->Z; int nn = n; int64_t ldv = n; int jj = j;
            double* parts = k->d_partials; double* Hcol = k->d_small + 11 * 64 * 64;     // its+1 doubles
            void* args[] = {&wv, &Vb, &Zb, &nn, &ldv, &jj, &parts, &Hcol};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)mgs_sweep_kernel, dim3(grid), dim3(kKryThreads), args, 0, ctx->stream);
            if (e != cudaSuccess) { rc = kexp::fail(ctx, std::string("cooperative MGS launch: ") + cudaGetErrorString(e)); break; }
            ctx->st_launches++;
            // H(0..j, j)
            e = cudaMemcpy2DAsync(k->d_H + j, sizeof(double) * ldh, Hcol, sizeof(double), sizeof(double), j + 1, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) { rc = kexp::fail(ctx, cudaGetErrorString(e)); break; }
        }
        // H[j+1,j] = ||w||_M  (krylov_exp.py:78), V_{j+1} = w / H[j+1,j], Z_{j+1} = M V_{j+1}: scalars stay on the device
        spmv(ctx, k->M, k->w, k->mw);
        {
            DotArgs args;
            memset(&args, 0, sizeof(args));
            args.count = 1; args.a[0] = k->w; args.b[0] = k->mw; args.n[0] = n;
            multidot_kernel<<<k->grid, kKryThreads, 0, ctx->stream>>>(args, k->d_partials, k->d_counter, k->d_dots);
            arnoldi_norm_kernel<<<1, 1, 0, ctx->stream>>>(k->d_dots, k->d_H + (size_t)(j + 1) * ldh + j, k->d_scal + 2, k->d_info, k->d_scal + 1);
            scale2_kernel<<<grid_for(n, 4 * k->grid), kKryThreads, 0, ctx->stream>>>(k->V + (size_t)(j + 1) * n, k->Z + (size_t)(j + 1) * n,
                                                                                      k->w, k->mw, k->d_scal + 2, n);
            ctx->st_launches += 3;
        }
        if (d_res) {
            // small dense part on one CTA: inv(H), expm(-H1 s t) e1 for s = 1/3, 2/3, 1   (krylov_exp.py:87-99)
            hessenberg_exp_kernel<<<1, kDenseThreads, 0, ctx->stream>>>(k->d_H, ldh, j + 1, k->gamma, t, k->d_svals, 3, k->d_u, k->d_fac, k->d_small);
            ctx->st_launches++;
            // C = || M^-1 K w ||_M with w the orthogonalised, unnormalised vector (krylov_exp.py:92-94); w stays on the
            // device.  NOTE: the operator is this context's M^-1 K (kexp_set_operator); the reference applies whatever
            // operator_solver the caller passed -- see DESIGN.md for this documented deviation.
            if (kexp_batch_upload_dev(bi, k->w, n, 0) || kexp_operator_apply_dev(ctx, bi, bo) || kexp_batch_export_dev(bo, k->tmp, n)) { rc = 1; break; }
            spmv(ctx, k->M, k->tmp, k->mw);
            DotArgs args;
            memset(&args, 0, sizeof(args));
            args.count = 1; args.a[0] = k->tmp; args.b[0] = k->mw; args.n[0] = n;
            multidot_kernel<<<k->grid, kKryThreads, 0, ctx->stream>>>(args, k->d_partials, k->d_counter, k->d_dots);
            residual_kernel<<<1, 1, 0, ctx->stream>>>(k->d_dots, k->gamma, k->d_fac, 3, d_res + j);
            ctx->st_launches += 2;
        }
    }
    if (bi) kexp_batch_destroy(bi);
    if (bo) kexp_batch_destroy(bo);
    if (!rc) KEXP_CUDA(cudaGetLastError());
    return rc;
}

static int krylov_prepare(kexp_ctx* ctx, int kdim, int* its) {
    KrylovState* k = ctx->kry;
    if (!k) return kexp::fail(ctx, "krylov operator not set (kexp_set_krylov_operator)");
    if (kdim < 1 || kdim > 63) return kexp::fail(ctx, "kdim must be in 1..63");
    if (*its < 0) *its = kdim;                                       // krylov_exp.py:66-67
    if (*its > kdim) return kexp::fail(ctx, "its <= kdim violated (krylov_exp.py:68)");
    if (*its < 1) return kexp::fail(ctx, "its must be >= 1");
    const int n = (int)k->n;
    if (k->kdim < kdim) {
        cudaFree(k->V); cudaFree(k->Z); cudaFree(k->d_H);
        k->V = k->Z = k->d_H = nullptr;
        if (alloc_d(ctx, &k->V, (size_t)(kdim + 1) * n) || alloc_d(ctx, &k->Z, (size_t)(kdim + 1) * n) ||
            alloc_d(ctx, &k->d_H, (size_t)(kdim + 1) * kdim)) return 1;
        k->kdim = kdim;
    }
    if (!k->d_scal) {
        if (alloc_d(ctx, &k->d_scal, 8) || alloc_d(ctx, &k->d_res, 64)) return 1;
    }
    return 0;
}

// after the closing synchronisation: breakdown / non-convergence recorded on the device during the loop
static int krylov_outcome(kexp_ctx* ctx, bool* zero_input) {
    KrylovState* k = ctx->kry;
    double scal[4];
    int info[4];
    KEXP_CUDA(cudaMemcpy(scal, k->d_scal, sizeof(scal), cudaMemcpyDeviceToHost));
    KEXP_CUDA(cudaMemcpy(info, k->d_info, sizeof(info), cudaMemcpyDeviceToHost));
    KEXP_CUDA(cudaMemsetAsync(k->d_info, 0, sizeof(int) * 4, ctx->stream));
    k->info_pending = false;
    ctx->st_bicg += info[0];
    *zero_input = !(scal[1] > 0.0);
    if (*zero_input) return 0;                                       // exp(.) 0 = 0, no error
    if (info[1] == 2) return kexp::fail(ctx, "BiCGStab breakdown in the shifted solve");
    if (info[1] == 1) return kexp::fail(ctx, "shifted solve did not converge (raise solve_max_it)");
    if (info[3]) return kexp::fail(ctx, "Arnoldi breakdown (invariant subspace reached)");
    return 0;
}

int kexp_krylov_apply(kexp_ctx* ctx, const double* x, double* y, double t, int kdim, int its,
                      double* H_out, double* u_out, double* residuals_out) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!x || !y) return kexp::fail(ctx, "null vector");
    if (x == y) return kexp::fail(ctx, "x and y must not alias");
    if (krylov_prepare(ctx, kdim, &its)) return 1;
    KrylovState* k = ctx->kry;
    const int n = (int)k->n;
    KEXP_CUDA(cudaMemcpyAsync(k->w, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (arnoldi_dev(ctx, kdim, its, t, residuals_out ? k->d_res : nullptr)) return 1;
    // inv(H), expm(-H1 t) e1 (krylov_exp.py:87-99 at s = 1), y = beta sum_j u_j V_j (:102-109)
    hessenberg_exp_kernel<<<1, kDenseThreads, 0, ctx->stream>>>(k->d_H, kdim, its, k->gamma, t, k->d_svals, 3, k->d_u, k->d_fac, k->d_small);
    combine_batch_kernel<<<grid_for(n, 4 * k->grid), kKryThreads, sizeof(double) * its, ctx->stream>>>(k->V, n, k->d_u, its, 1, k->d_scal, k->tmp, n, n);
    ctx->st_launches += 2;
    KEXP_CUDA(cudaMemcpyAsync(y, k->tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (H_out) KEXP_CUDA(cudaMemcpyAsync(H_out, k->d_H, sizeof(double) * (size_t)(kdim + 1) * kdim, cudaMemcpyDeviceToHost, ctx->stream));
    if (u_out) KEXP_CUDA(cudaMemcpyAsync(u_out, k->d_u, sizeof(double) * its, cudaMemcpyDeviceToHost, ctx->stream));
    if (residuals_out) KEXP_CUDA(cudaMemcpyAsync(residuals_out, k->d_res, sizeof(double) * its, cudaMemcpyDeviceToHost, ctx->stream));
    KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
    KEXP_CUDA(cudaGetLastError());
    bool zero = false;
    if (krylov_outcome(ctx, &zero)) return 1;
    if (zero) {                                                      // zero input: exp(.) 0 = 0
        memset(y, 0, sizeof(double) * n);
        if (H_out) memset(H_out, 0, sizeof(double) * (size_t)(kdim + 1) * kdim);
        if (u_out) memset(u_out, 0, sizeof(double) * its);
        if (residuals_out) memset(residuals_out, 0, sizeof(double) * its);
    }
    return 0;
}

// B sample times at once (krylov_exp.apply for every ensemble member, averaged_sw_explicit.py:233-235 with the
// Krylov propagator).  x_is_shared: the shift-and-invert Krylov space K_m((I + gamma A)^-1, x) does not depend on t, so
// ONE Arnoldi process serves every t_k and only the small exponentials expm(-t_k H1) e1 and the final combination
// differ.  Distinct inputs run their Arnoldi processes back to back on the device queue (no host wait in between).
int kexp_krylov_apply_batch(kexp_ctx* ctx, const double* X, int64_t ldx, int x_is_shared, double* Y, int64_t ldy,
                            const double* t, int B, int kdim, int its) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (!X || !Y || !t || B < 1) return kexp::fail(ctx, "bad arguments");
    if (krylov_prepare(ctx, kdim, &its)) return 1;
    KrylovState* k = ctx->kry;
    const int n = (int)k->n;
    if (ldy < n || (!x_is_shared && ldx < n)) return kexp::fail(ctx, "leading dimension < n");
    if (k->out_cap < (size_t)B * n) {
        cudaFree(k->d_out); cudaFree(k->d_ub);
        k->d_out = k->d_ub = nullptr; k->out_cap = 0;
        if (alloc_d(ctx, &k->d_out, (size_t)B * n) || alloc_d(ctx, &k->d_ub, (size_t)B * 64)) return 1;
        k->out_cap = (size_t)B * n;
    }
    const int cgrid = grid_for(n, 4 * k->grid);
    const int nproc = x_is_shared ? 1 : B;
    for (int p = 0; p < nproc; ++p) {
        KEXP_CUDA(cudaMemcpyAsync(k->w, X + (size_t)p * (x_is_shared ? 0 : ldx), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        if (arnoldi_dev(ctx, kdim, its, 0.0, nullptr)) return 1;
        const int c0 = x_is_shared ? 0 : p, c1 = x_is_shared ? B : p + 1;
        for (int c = c0; c < c1; ++c) {
            hessenberg_exp_kernel<<<1, kDenseThreads, 0, ctx->stream>>>(k->d_H, kdim, its, k->gamma, t[c], k->d_svals + 2, 1,
                                                                      k->d_ub + (size_t)c * 64, k->d_fac, k->d_small);
            ctx->st_launches++;
        }
        combine_batch_kernel<<<cgrid, kKryThreads, sizeof(double) * its * (c1 - c0), ctx->stream>>>(
            k->V, n, k->d_ub + (size_t)c0 * 64, its, c1 - c0, k->d_scal, k->d_out + (size_t)c0 * n, n, n);
        ctx->st_launches++;
    }
    KEXP_CUDA(cudaMemcpy2DAsync(Y, sizeof(double) * ldy, k->d_out, sizeof(double) * n, sizeof(double) * n, B, cudaMemcpyDeviceToHost, ctx->stream));
    KEXP_CUDA(cudaStreamSynchronize(ctx->stream));
    KEXP_CUDA(cudaGetLastError());
    bool zero = false;
    if (krylov_outcome(ctx, &zero)) return 1;      // sticky status over all columns; `zero` refers to the last start vector
    return 0;
}

}  // extern "C"
