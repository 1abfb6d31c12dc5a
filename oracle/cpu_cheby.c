/* TEST INFRASTRUCTURE (oracle): plain C / OpenMP restatement of the hot path, used as
 *   (1) an independent CPU checker at sizes where the SciPy oracle is slow, and
 *   (2) the CPU baseline timed by bench.py (cpu_baseline, --impl reference).
 * Never linked into or called from the product (krylov_exp_b200/).
 *
 * Follows the reference's algorithm:
 *   operator_solver.solve()  = z = K v ; CG on M_u ; exact 3x3 blocks for M_eta
 *       /root/reference/averaged_sw_explicit.py:142-164 (fieldsplit: cg on field 0,
 *       preonly on the block-diagonal field 1)
 *   cheby_exp.apply           /root/reference/cheby_exp.py:110-165
 *       mode 0: the collapsed real recurrence S_n = 2(t/L) A S_{n-1} + S_{n-2}
 *       mode 1: the literal two-vector recurrence with 2*ncheb-1 operator solves
 *   parallelism               one sample time per core, as the reference's one ensemble
 *       member per MPI rank (/root/reference/averaged_sw_explicit.py:37,233-235)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t n_u, n_eta;
    const int32_t *Krp, *Kci; const double *Kv;
    const int32_t *Mrp, *Mci; const double *Mv;
    const double *Meta_inv;       /* n_eta/3 blocks, row-major 3x3 */
    const double *Pinv;           /* n_u/3 blocks: block-Jacobi preconditioner */
    double rtol; int max_it;
} op_t;

static void spmv(int64_t n, const int32_t *rp, const int32_t *ci, const double *v, const double *x, double *y) {
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k) s += v[k] * x[ci[k]];
        y[i] = s;
    }
}

static void blk3(const double *B, int64_t nb, const double *r, double *z) {
    for (int64_t b = 0; b < nb; ++b) {
        const double *P = B + 9 * b; const double *rr = r + 3 * b; double *zz = z + 3 * b;
        for (int i = 0; i < 3; ++i) zz[i] = P[3 * i] * rr[0] + P[3 * i + 1] * rr[1] + P[3 * i + 2] * rr[2];
    }
}

static double dot(int64_t n, const double *a, const double *b) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* w = M^-1 K v ; work: 4*n_u + n doubles; returns CG iterations */
static int apply_op(const op_t *o, const double *v, double *w, double *work) {
    const int64_t nu = o->n_u, ne = o->n_eta, n = nu + ne;
    double *z = work, *r = z + n, *p = r + nu, *q = p + nu, *pz = q + nu;
    spmv(n, o->Krp, o->Kci, o->Kv, v, z);
    blk3(o->Meta_inv, ne / 3, z + nu, w + nu);
    memset(w, 0, sizeof(double) * nu);
    memcpy(r, z, sizeof(double) * nu);
    blk3(o->Pinv, nu / 3, r, p);
    double rz = dot(nu, r, p), bb = dot(nu, r, r);
    int it = 0;
    if (bb == 0.0) return 0;
    for (; it < o->max_it;) {
        spmv(nu, o->Mrp, o->Mci, o->Mv, p, q);
        double alpha = rz / dot(nu, p, q);
        for (int64_t i = 0; i < nu; ++i) { w[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
        ++it;
        if (dot(nu, r, r) <= o->rtol * o->rtol * bb) break;
        blk3(o->Pinv, nu / 3, r, pz);
        double rz2 = dot(nu, r, pz), beta = rz2 / rz;
        rz = rz2;
        for (int64_t i = 0; i < nu; ++i) p[i] = pz[i] + beta * p[i];
    }
    return it;
}

/* Y[k] = sum_n a_n S_n(X[k], t[k]);  X,Y: B vectors of n doubles (ld = n).
 * mode 0: real recurrence (ncheb solves); mode 1: literal complex recurrence (2 ncheb - 1
 * solves, c = complex coefficients interleaved re,im).  Returns total CG iterations. */
int64_t cpu_cheby_apply_batch(int64_t n_u, int64_t n_eta,
                              const int32_t *Krp, const int32_t *Kci, const double *Kv,
                              const int32_t *Mrp, const int32_t *Mci, const double *Mv,
                              const double *Meta_inv, const double *Pinv,
                              const double *coef, int ncheb, double L, int mode,
                              const double *X, int x_shared, const double *t, int B, double *Y,
                              double rtol, int max_it, int nthreads) {
    op_t o = {n_u, n_eta, Krp, Kci, Kv, Mrp, Mci, Mv, Meta_inv, Pinv, rtol, max_it};
    const int64_t n = n_u + n_eta;
    int64_t total_it = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_it)
    for (int k = 0; k < B; ++k) {
        const double *x = X + (x_shared ? 0 : (int64_t)k * n);
        double *y = Y + (int64_t)k * n;
        double *buf = (double *)malloc(sizeof(double) * (size_t)(8 * n + 4 * n_u));
        double *work = buf, *Sa = buf + (n + 4 * n_u), *Sb = Sa + n, *W = Sb + n;
        double *Ta = W + n, *Tb = Ta + n, *Tc = Tb + n, *Td = Tc + n;   /* literal mode extras */
        const double s = t[k] / L;
        if (mode == 0) {
            memcpy(Sa, x, sizeof(double) * n);
            for (int64_t i = 0; i < n; ++i) y[i] = coef[0] * x[i];
            double *Sm1 = Sa, *Sm2 = Sb;
            for (int m = 1; m <= ncheb; ++m) {
                total_it += apply_op(&o, Sm1, W, work);
                const double cs = (m == 1) ? s : 2.0 * s;
                if (m == 1) for (int64_t i = 0; i < n; ++i) Sm2[i] = cs * W[i];
                else for (int64_t i = 0; i < n; ++i) Sm2[i] = cs * W[i] + Sm2[i];
                for (int64_t i = 0; i < n; ++i) y[i] += coef[m] * Sm2[i];
                double *tmp = Sm1; Sm1 = Sm2; Sm2 = tmp;
            }
        } else {
            /* cheby_exp.py:114-165, c[2m], c[2m+1] = Re, Im of ChebCoeffs[m] */
            double *Tm1r = Sa, *Tm1i = Sb, *Tm2r = Ta, *Tm2i = Tb, *Tr = Tc, *Ti = Td;
            memcpy(Tm1r, x, sizeof(double) * n); memset(Tm1i, 0, sizeof(double) * n);
            for (int64_t i = 0; i < n; ++i) y[i] = coef[0] * Tm1r[i];
            total_it += apply_op(&o, x, W, work);
            memset(Tr, 0, sizeof(double) * n);
            for (int64_t i = 0; i < n; ++i) { Ti[i] = W[i] * (-t[k] / L); y[i] += -coef[3] * Ti[i]; }
            for (int m = 2; m <= ncheb; ++m) {
                double *a = Tm2r, *b = Tm2i;
                Tm2r = Tm1r; Tm2i = Tm1i; Tm1r = Tr; Tm1i = Ti; Tr = a; Ti = b;
                total_it += apply_op(&o, Tm1r, W, work);
                for (int64_t i = 0; i < n; ++i) Ti[i] = W[i] * (-2.0 * t[k] / L) - Tm2i[i];
                total_it += apply_op(&o, Tm1i, W, work);
                for (int64_t i = 0; i < n; ++i) Tr[i] = W[i] * (2.0 * t[k] / L) - Tm2r[i];
                for (int64_t i = 0; i < n; ++i) { y[i] += coef[2 * m] * Tr[i]; y[i] += -coef[2 * m + 1] * Ti[i]; }
            }
        }
        free(buf);
    }
    return total_it;
}

int cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
