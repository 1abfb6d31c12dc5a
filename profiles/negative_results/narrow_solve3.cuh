// Negative result (round 2): W = 2 mass solve with THREE threads per block row and the matrix rows in registers
// (384 threads, two passes).  Bit-identical iterates, no spills, but SLOWER than one thread per block row:
// 2.42 k cycles against 1.96 k cycles per Chebyshev iteration at ref 5 (profiles/r02_summary.md).  The iteration is bound by
// shared-memory WAVEFRONTS of the x gathers: a warp that covers 10 block rows instead of 32 issues three times as many
// load instructions for the same rows, and that costs more than the matrix reads it saves.
// The same solve for W = 2 with THREE threads per block row (thread (row, i) owns scalar row i of the 3x3 blocks): a
// single-column recurrence is pure latency -- one block row per thread means a chain of 90 dependent-ish products behind
// 60 shared-memory loads, and every iteration re-reads the CTA's 75 KB of matrix from shared memory, which costs more
// shared-memory cycles than everything else of the iteration.  Here a thread holds its 15 matrix entries per row (row i of
// the 5 blocks), row i of P^-1 and b_i in REGISTERS for the whole solve and only reads x from shared memory (the three
// threads of a row read the same addresses: one broadcast).  The products of a row run in the same order as in narrow_solve
// and in the wide-panel kernels (block by block, columns r = 0, 1, 2), so the iterates are bit-identical.  z = P^-1 r needs
// the residuals of the two sibling threads: warp shuffles.  Warp layout: lane = 3 * (row in warp) + i, 10 rows per warp,
// NP3 passes of 10 * NT / 32 rows.  Requires n_own <= NP3 * 10 * NT / 32 and at most 5 blocks per row (host check).
template <int NT, int NP3>
KEXP_PHASE_FN unsigned narrow_solve3(int slot, int step, unsigned xid
#if KEXP_NARROW_TIMING
                                     , long long* s_tim
#endif
) {
    constexpr int W = 2, SS = NarrowCfg<W>::SS, kNWarps = NT / 32, RPW = 10;
    const NarrowArgs& a = c_nargs[slot];
    const NarrowSmem sm = narrow_smem_carve<W>(a);
    const StepCtx sx = make_step(a, step);
    const int K = a.kterm[step];
    int s0, s1;
    cta_range(a.s_nbr, s0, s1);
    const int n_own = s1 - s0;
    const int h0 = __ldg(a.halo_ptr + blockIdx.x), n_h1 = __ldg(a.halo_ptr + blockIdx.x + 1) - h0;
    const int nsend = __ldg(a.cta_send_ptr + blockIdx.x + 1) - __ldg(a.cta_send_ptr + blockIdx.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = lane % 3;
    const int sib = lane - i;                            // lane of the row's thread 0
    const int ell_ld = sm.ell_ld;
    const size_t inbox_half = (size_t)a.inbox_slots * (3 * W);
    const uint4* my_inbox = a.inbox + (size_t)h0 * (3 * W);
    const int ne = n_h1 * 3 * W;
    double* cur = sm.xo0;
    double* prev = sm.xo1;
    // the thread's share of the matrix, of P^-1 and of b
    double m[NP3][5][3], P[NP3][3];
    double2 bv[NP3];
    int sl[NP3][5], lrow[NP3];
    bool valid[NP3];
#pragma unroll
    for (int p = 0; p < NP3; ++p) {
        const int lr = (p * kNWarps + warp) * RPW + lane / 3;
        valid[p] = lane < 3 * RPW && lr < n_own;
        lrow[p] = valid[p] ? lr : 0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const bool have = valid[p] && q < sm.scnt[lr];          // shorter rows: zero blocks on the row's own slot
            sl[p][q] = have ? sm.ss[q * ell_ld + lr] : lrow[p];
#pragma unroll
            for (int r = 0; r < 3; ++r) m[p][q][r] = have ? sm.sv[(size_t)(9 * q + 3 * i + r) * ell_ld + lr] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) P[p][r] = valid[p] ? sm.sP[(3 * i + r) * ell_ld + lr] : 0.0;
    }
    NTIM_T0
    // ---- start: b <- T (global), x_1 = alpha P^-1 b, x_0 = 0
    double2 accb = make_double2(0.0, 0.0);
#pragma unroll
    for (int p = 0; p < NP3; ++p) {
        const int ls = lrow[p] * SS + i * W;
        bv[p] = valid[p] ? ldcg2(a.T + (size_t)(3 * (s0 + lrow[p]) + i) * W) : make_double2(0.0, 0.0);
        accb.x = fma(bv[p].x, bv[p].x, accb.x); accb.y = fma(bv[p].y, bv[p].y, accb.y);
        const double2 b0 = make_double2(__shfl_sync(0xffffffffu, bv[p].x, sib), __shfl_sync(0xffffffffu, bv[p].y, sib));
        const double2 b1 = make_double2(__shfl_sync(0xffffffffu, bv[p].x, sib + 1), __shfl_sync(0xffffffffu, bv[p].y, sib + 1));
        const double2 b2 = make_double2(__shfl_sync(0xffffffffu, bv[p].x, sib + 2), __shfl_sync(0xffffffffu, bv[p].y, sib + 2));
        double2 z;
        z.x = a.alpha * fma(P[p][0], b0.x, fma(P[p][1], b1.x, P[p][2] * b2.x));
        z.y = a.alpha * fma(P[p][0], b0.y, fma(P[p][1], b1.y, P[p][2] * b2.y));
        if (valid[p]) {
            *reinterpret_cast<double2*>(cur + ls) = z;
            *reinterpret_cast<double2*>(prev + ls) = make_double2(0.0, 0.0);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        accb.x += __shfl_xor_sync(0xffffffffu, accb.x, off);
        accb.y += __shfl_xor_sync(0xffffffffu, accb.y, off);
    }
    if (lane == 0) { sm.sred[(1 * kNWarps + warp) * W] = accb.x; sm.sred[(1 * kNWarps + warp) * W + 1] = accb.y; }
    __syncthreads();                                    // x_1 on the own rows complete
    narrow_send<NT, 3 * W, SS>(a.inbox + (size_t)(xid & 1u) * inbox_half, sm.sptr, sm.sdst, nsend, cur, xid);
    NTIM_ADD(6)
    for (int it = 1; it < K; ++it) {
        ll_recv<NT, 3 * W, SS>(my_inbox + (size_t)(xid & 1u) * inbox_half, sm.xh, ne, xid);
        __syncthreads();                                // own rows and halo of x_it complete
        NTIM_ADD(7)
        ++xid;
        const double omega = a.omega[it];
        const bool last = (it == K - 1);
        double2 accr = make_double2(0.0, 0.0);
#pragma unroll
        for (int p = 0; p < NP3; ++p) {
            const int ls = lrow[p] * SS + i * W;
            double2 acc = make_double2(0.0, 0.0);
            // x rows of two blocks in flight; the compiler barriers keep ptxas from hoisting the loads of all ten blocks of
            // the two passes (120 registers) and spilling the matrix instead
            double2 xq[2][3];
            auto load_x = [&](int q, double2 (&x)[3]) {
                const int s = sl[p][q];
                const double* xr = (s >= 0 ? cur + s * SS : sm.xh + (-s - 1) * SS);
                x[0] = *reinterpret_cast<const double2*>(xr); x[1] = *reinterpret_cast<const double2*>(xr + W);
                x[2] = *reinterpret_cast<const double2*>(xr + 2 * W);
            };
            load_x(0, xq[0]);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                if (q + 1 < 5) load_x(q + 1, xq[(q + 1) & 1]);
                asm volatile("" ::: "memory");
                const double2 (&x)[3] = xq[q & 1];
                acc.x = fma(m[p][q][0], x[0].x, acc.x); acc.y = fma(m[p][q][0], x[0].y, acc.y);
                acc.x = fma(m[p][q][1], x[1].x, acc.x); acc.y = fma(m[p][q][1], x[1].y, acc.y);
                acc.x = fma(m[p][q][2], x[2].x, acc.x); acc.y = fma(m[p][q][2], x[2].y, acc.y);
            }
            const double2 rv = make_double2(bv[p].x - acc.x, bv[p].y - acc.y);
            const double2 r0 = make_double2(__shfl_sync(0xffffffffu, rv.x, sib), __shfl_sync(0xffffffffu, rv.y, sib));
            const double2 r1 = make_double2(__shfl_sync(0xffffffffu, rv.x, sib + 1), __shfl_sync(0xffffffffu, rv.y, sib + 1));
            const double2 r2 = make_double2(__shfl_sync(0xffffffffu, rv.x, sib + 2), __shfl_sync(0xffffffffu, rv.y, sib + 2));
            const double2 xo = *reinterpret_cast<const double2*>(cur + ls);
            const double2 xmv = *reinterpret_cast<const double2*>(prev + ls);
            double2 z, xn;
            z.x = fma(P[p][0], r0.x, fma(P[p][1], r1.x, P[p][2] * r2.x));
            z.y = fma(P[p][0], r0.y, fma(P[p][1], r1.y, P[p][2] * r2.y));
            xn.x = fma(omega, fma(a.alpha, z.x, xo.x) - xmv.x, xmv.x);
            xn.y = fma(omega, fma(a.alpha, z.y, xo.y) - xmv.y, xmv.y);
            if (!last) {
                if (valid[p]) *reinterpret_cast<double2*>(prev + ls) = xn;
            } else if (valid[p]) {
                accr.x = fma(rv.x, rv.x, accr.x); accr.y = fma(rv.y, rv.y, accr.y);
                const size_t o = (size_t)(3 * (s0 + lrow[p]) + i) * W;
                const UpdCtx u = make_upd<W>(a, sx, 0);
                *reinterpret_cast<double2*>(a.X + o) = xn;
                row_update(u, o, (size_t)(3 * (s0 + lrow[p]) + i), xn);     // edge rows: W = x_K
            }
        }
        NTIM_ADD(1)
        if (!last) {
            __syncthreads();                            // x_{it+1} on the own rows complete (and nobody reads the halo of x_it any more)
            narrow_send<NT, 3 * W, SS>(a.inbox + (size_t)(xid & 1u) * inbox_half, sm.sptr, sm.sdst, nsend, prev, xid);
        } else {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                accr.x += __shfl_xor_sync(0xffffffffu, accr.x, off);
                accr.y += __shfl_xor_sync(0xffffffffu, accr.y, off);
            }
            if (lane == 0) { sm.sred[(0 * kNWarps + warp) * W] = accr.x; sm.sred[(0 * kNWarps + warp) * W + 1] = accr.y; }
        }
        NTIM_ADD(3)
        double* tsw = cur; cur = prev; prev = tsw;
    }
    __syncthreads();
    if (threadIdx.x < 2 * W) {                           // block sums of ||r||^2 and ||b||^2 per column (fixed order)
        const int q = threadIdx.x / W, c = threadIdx.x % W;
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kNWarps; ++w) t += sm.sred[(q * kNWarps + w) * W + c];
        a.partials[((size_t)blockIdx.x * 2 + q) * W + c] = t;
    }
    NTIM_ADD(10)
    return xid;
}

