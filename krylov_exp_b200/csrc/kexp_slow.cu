// kexp_slow.cu -- the nonlinear slow 1-form L(U) of the averaged model for a batch of ensemble members on the device
// (SURVEY section 8, row N3).
//
// Reference: /root/reference/averaged_sw_explicit.py:177-213, vector-invariant branch (:192-201):
//     L = dT [ <perp(grad(v.perp(u))), u> dx - <both(perp(n) (v.perp(u))), both(Upwind u)> dS + div(v) K dx
//            + <grad(phi), u> (eta - b) dx  - jump(phi) (uup+ (eta+ - b+) - uup- (eta- - b-)) dS ]
// evaluated in reference coordinates from the per-cell geometry tables of krylov_exp_b200/slow_rhs.py (the formulas are
// derived in that module's docstring; oracle/slow_rhs_ref.py is the NumPy restatement this kernel is checked against).
//
// One thread = one (cell, ensemble member): the 12 signed BDM2 and 3 DG1 coefficients of the cell and, per edge, of the
// neighbour across it live in registers; basis tabulations are staged in shared memory; the geometry tables are read
// once per cell and broadcast to the members that share a warp.  Edge dofs receive exactly two contributions (one per
// adjacent cell) through atomicAdd on a zeroed vector -- two-term sums are order independent, so the result is
// deterministic; interior and depth rows are plain stores.  State layout: B vectors of n doubles in the user dof order
// [u ; eta], vector k at U + k*ldu (the layout of kexp_batch_export_dev / the torch tensors of DeviceAveragedStepper).
#include <cuda_runtime.h>
#include <stdint.h>

namespace kexp {

constexpr int kSlowThreads = 128;
constexpr int kMaxNq = 36;      // cell quadrature points (6 x 6 collapsed Gauss-Jacobi)
constexpr int kMaxNf = 6;       // Gauss points per facet
constexpr int kNCell = 11;      // doubles per (cell, point): Ginv00 Ginv01 Ginv11 G00 G01 G11 w*m mu dmu0 dmu1 w/(2 detJ^2)
constexpr int kNFac = 14;       // doubles per (cell, edge, point): Jd0[3] Jd1[3] n[3] perp(n)[3] ds mu

struct SlowArgs {
    int n_cells, n_u, n_eta, nq, nf;
    const double *cell, *fac, *Phi, *dPhi, *gPhi, *Psi, *gPsi, *PhiF, *PsiF, *qw, *sign, *b;
    const int *dofs, *nbr_cell, *nbr_edge;
    const double* U; int64_t ldu;
    double* L; int64_t ldl;
    int B;
    double dT;
};

__global__ void __launch_bounds__(kSlowThreads)
slow_form_kernel(const SlowArgs a) {
    __shared__ double sPhi[kMaxNq * 24], sgPhi[kMaxNq * 48], sdPhi[kMaxNq * 12], sPsi[kMaxNq * 3], sgPsi[6], sqw[kMaxNq];
    __shared__ double sPhiF[3 * kMaxNf * 24], sPsiF[3 * kMaxNf * 3];
    for (int i = threadIdx.x; i < a.nq * 24; i += kSlowThreads) sPhi[i] = a.Phi[i];
    for (int i = threadIdx.x; i < a.nq * 48; i += kSlowThreads) sgPhi[i] = a.gPhi[i];
    for (int i = threadIdx.x; i < a.nq * 12; i += kSlowThreads) sdPhi[i] = a.dPhi[i];
    for (int i = threadIdx.x; i < a.nq * 3; i += kSlowThreads) sPsi[i] = a.Psi[i];
    for (int i = threadIdx.x; i < a.nq; i += kSlowThreads) sqw[i] = a.qw[i];
    for (int i = threadIdx.x; i < 3 * a.nf * 24; i += kSlowThreads) sPhiF[i] = a.PhiF[i];
    for (int i = threadIdx.x; i < 3 * a.nf * 3; i += kSlowThreads) sPsiF[i] = a.PsiF[i];
    if (threadIdx.x < 6) sgPsi[threadIdx.x] = a.gPsi[threadIdx.x];
    __syncthreads();

    const int64_t total = (int64_t)a.n_cells * a.B;
    for (int64_t idx = (int64_t)blockIdx.x * kSlowThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kSlowThreads) {
        const int c = (int)(idx / a.B), col = (int)(idx % a.B);
        const double* Uc = a.U + (size_t)col * a.ldu;
        double* Lc = a.L + (size_t)col * a.ldl;
        double ul[12], dl[3], eu[12], ep[3];
        int dof[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            dof[j] = __ldg(a.dofs + 12 * c + j);
            ul[j] = Uc[dof[j]] * __ldg(a.sign + 12 * c + j);
            eu[j] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dl[i] = Uc[a.n_u + 3 * c + i] - __ldg(a.b + 3 * c + i);          // eta0 - b
            ep[i] = 0.0;
        }
        // ---------------- cell integrals
        for (int q = 0; q < a.nq; ++q) {
            const double* t = a.cell + ((size_t)c * a.nq + q) * kNCell;
            const double gi00 = __ldg(t), gi01 = __ldg(t + 1), gi11 = __ldg(t + 2), g00 = __ldg(t + 3), g01 = __ldg(t + 4), g11 = __ldg(t + 5);
            const double wm = __ldg(t + 6), mu = __ldg(t + 7), dmu0 = __ldg(t + 8), dmu1 = __ldg(t + 9), ke = __ldg(t + 10);
            const double* P = sPhi + q * 24;            // [j][a]
            const double* gP = sgPhi + q * 48;          // [j][a][b]
            double uh0 = 0.0, uh1 = 0.0, d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;     // d_ab = d uh_a / d xi_b
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                uh0 = fma(ul[j], P[2 * j], uh0); uh1 = fma(ul[j], P[2 * j + 1], uh1);
                d00 = fma(ul[j], gP[4 * j], d00); d01 = fma(ul[j], gP[4 * j + 1], d01);
                d10 = fma(ul[j], gP[4 * j + 2], d10); d11 = fma(ul[j], gP[4 * j + 3], d11);
            }
            const double up0 = -uh1, up1 = uh0;                                // uh_perp
            const double h0 = wm * (gi00 * up0 + gi01 * up1), h1 = wm * (gi01 * up0 + gi11 * up1);
            const double kin = ke * (g00 * uh0 * uh0 + 2.0 * g01 * uh0 * uh1 + g11 * uh1 * uh1);
            const double Dq = dl[0] * sPsi[3 * q] + dl[1] * sPsi[3 * q + 1] + dl[2] * sPsi[3 * q + 2];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const double p0 = P[2 * i], p1 = P[2 * i + 1];
                const double cx = p0 * uh1 - p1 * uh0;                          // Phi_i x uh
                const double dc0 = (gP[4 * i] * uh1 - gP[4 * i + 2] * uh0) + (p0 * d10 - p1 * d00);
                const double dc1 = (gP[4 * i + 1] * uh1 - gP[4 * i + 3] * uh0) + (p0 * d11 - p1 * d01);
                const double t1 = h0 * (dmu0 * cx + mu * dc0) + h1 * (dmu1 * cx + mu * dc1);
                eu[i] += t1 + sdPhi[12 * q + i] * kin;                          // T1 + T3
            }
            const double wq = sqw[q] * Dq;
#pragma unroll
            for (int i = 0; i < 3; ++i) ep[i] += wq * (sgPsi[2 * i] * uh0 + sgPsi[2 * i + 1] * uh1);   // T4
        }
        // ---------------- interior facets
        for (int e = 0; e < 3; ++e) {
            const int nb = __ldg(a.nbr_cell + 3 * c + e), ne = __ldg(a.nbr_edge + 3 * c + e);
            double un_[12], dn_[3];
#pragma unroll
            for (int j = 0; j < 12; ++j) un_[j] = Uc[__ldg(a.dofs + 12 * nb + j)] * __ldg(a.sign + 12 * nb + j);
#pragma unroll
            for (int i = 0; i < 3; ++i) dn_[i] = Uc[a.n_u + 3 * nb + i] - __ldg(a.b + 3 * nb + i);
            for (int f = 0; f < a.nf; ++f) {
                const int fn_ = a.nf - 1 - f;                                   // the neighbour runs along the edge the other way
                const double* fo = a.fac + (((size_t)c * 3 + e) * a.nf + f) * kNFac;
                const double* fg = a.fac + (((size_t)nb * 3 + ne) * a.nf + fn_) * kNFac;
                const double* PO = sPhiF + (e * a.nf + f) * 24;
                const double* PN = sPhiF + (ne * a.nf + fn_) * 24;
                const double* SO = sPsiF + (e * a.nf + f) * 3;
                const double* SN = sPsiF + (ne * a.nf + fn_) * 3;
                double uo0 = 0.0, uo1 = 0.0, un0 = 0.0, un1 = 0.0;
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    uo0 = fma(ul[j], PO[2 * j], uo0); uo1 = fma(ul[j], PO[2 * j + 1], uo1);
                    un0 = fma(un_[j], PN[2 * j], un0); un1 = fma(un_[j], PN[2 * j + 1], un1);
                }
                double u3o[3], u3n[3], dotO = 0.0, dotN = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    u3o[d] = __ldg(fo + d) * uo0 + __ldg(fo + 3 + d) * uo1;     // u0 as a 3-vector on either side
                    u3n[d] = __ldg(fg + d) * un0 + __ldg(fg + 3 + d) * un1;
                    dotO = fma(u3o[d], __ldg(fo + 6 + d), dotO);                // dot(u0, n), each side with its own n
                    dotN = fma(u3n[d], __ldg(fg + 6 + d), dotN);
                }
                const double sgO = (dotO > 0.0) ? 1.0 : ((dotO < 0.0) ? -1.0 : 0.0), sgN = (dotN > 0.0) ? 1.0 : ((dotN < 0.0) ? -1.0 : 0.0);
                const double upO = 0.5 * (sgO + 1.0), upN = 0.5 * (sgN + 1.0);  // Upwind
                double PY = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) PY = fma(__ldg(fo + 9 + d), upO * u3o[d] + upN * u3n[d], PY);
                const double ds = __ldg(fo + 12), muf = __ldg(fo + 13);
                const double DO = dl[0] * SO[0] + dl[1] * SO[1] + dl[2] * SO[2];
                const double DN = dn_[0] * SN[0] + dn_[1] * SN[1] + dn_[2] * SN[2];
                const double flux = 0.5 * (dotO + fabs(dotO)) * DO - 0.5 * (dotN + fabs(dotN)) * DN;
                const double cT2 = ds * muf * PY;
#pragma unroll
                for (int i = 0; i < 12; ++i) eu[i] += cT2 * (PO[2 * i] * uo1 - PO[2 * i + 1] * uo0);       // -T2
#pragma unroll
                for (int i = 0; i < 3; ++i) ep[i] -= ds * SO[i] * flux;                                      // -T5
            }
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const double v = a.dT * eu[j] * __ldg(a.sign + 12 * c + j);
            if (j < 9) atomicAdd(Lc + dof[j], v);                               // edge dofs: shared with the neighbour cell
            else Lc[dof[j]] = v;                                                // interior dofs: this cell only
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) Lc[a.n_u + 3 * c + i] = a.dT * ep[i];
    }
}

}  // namespace kexp

extern "C" int kexp_slow_form_dev(void* stream, int n_cells, int n_u, int n_eta, int nq, int nf,
                                  const double* cell, const double* fac, const double* Phi, const double* dPhi, const double* gPhi,
                                  const double* Psi, const double* gPsi, const double* PhiF, const double* PsiF, const double* qw,
                                  const double* sign, const int* dofs, const int* nbr_cell, const int* nbr_edge, const double* b,
                                  const double* U, int64_t ldu, double* L, int64_t ldl, int B, double dT) {
    using namespace kexp;
    if (n_cells < 1 || B < 1 || nq < 1 || nq > kMaxNq || nf < 1 || nf > kMaxNf || !U || !L) return 1;
    if (ldu < (int64_t)n_u + n_eta || ldl < (int64_t)n_u + n_eta) return 2;
    cudaStream_t s = (cudaStream_t)stream;
    // the edge rows are accumulated with atomicAdd: start from zero (the interior / depth rows are overwritten)
    for (int k = 0; k < B; ++k)
        if (cudaMemsetAsync(L + (size_t)k * ldl, 0, sizeof(double) * (size_t)n_u, s) != cudaSuccess) return 3;
    SlowArgs a;
    a.n_cells = n_cells; a.n_u = n_u; a.n_eta = n_eta; a.nq = nq; a.nf = nf;
    a.cell = cell; a.fac = fac; a.Phi = Phi; a.dPhi = dPhi; a.gPhi = gPhi; a.Psi = Psi; a.gPsi = gPsi; a.PhiF = PhiF; a.PsiF = PsiF;
    a.qw = qw; a.sign = sign; a.b = b; a.dofs = dofs; a.nbr_cell = nbr_cell; a.nbr_edge = nbr_edge;
    a.U = U; a.ldu = ldu; a.L = L; a.ldl = ldl; a.B = B; a.dT = dT;
    const int64_t total = (int64_t)n_cells * B;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)((total + kSlowThreads - 1) / kSlowThreads < (int64_t)sms * 16 ? (total + kSlowThreads - 1) / kSlowThreads : (int64_t)sms * 16);
    slow_form_kernel<<<grid, kSlowThreads, 0, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
