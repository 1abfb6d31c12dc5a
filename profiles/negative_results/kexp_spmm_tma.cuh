// kexp_spmm_tma.cuh -- EXPERIMENTAL (off by default, KEXP_STAGED_SPMM=1): the CG product q = S p
// for 64-column panels with the gathered rows staged in shared memory by asynchronous copies and
// an mbarrier producer/consumer ring (4 producer warps + 16 consumer warps, 3 stages).
//
// Idea: cut the block rows into chunks of 16 mesh entities; the host precomputes the list of
// DISTINCT block columns of each chunk (<= 43: neighbouring entities share most of their stencil)
// and rewrites every block-column index as a slot in that list, so each gathered row crosses
// L2 -> SM once per chunk and the FMA loop only reads shared memory.
//
// Measured on B200 (ref 5, 64 columns; bsr3_spmm_kernel = 48 us):
//   cp.async.bulk (TMA) per 1.5 KB tile, 1 producer warp, __syncthreads ring   62 us
//   same, dedicated producer warp + full/empty mbarriers                         54 us
//   16-byte LDGSTS copies, 1 producer warp                                       76 us
//   16-byte LDGSTS copies, 4 producer warps                                      57 us
// i.e. ~40 small copies per chunk are request-rate bound in the copy engines before they are
// bandwidth bound, and one CTA per SM leaves nothing to overlap the fill with.  Kept for the
// record and for wider tiles (larger entity blocks would change the balance); not used.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kexp_kernels.cuh"

namespace kexp {

constexpr int kTmaW = 64;                 // panel width this kernel is built for
constexpr int kTmaRows = 16;              // block rows per chunk == warps per CTA
constexpr int kTmaThreads = 32 * kTmaRows;
constexpr int kTmaStages = 3;
constexpr int kTmaProducers = 4;          // producer warps (a warp sustains only ~50 outstanding copy instructions)
constexpr int kTmaTile = 3 * kTmaW * 8;   // bytes of one gathered column block (3 rows)

struct TmaStageLayout {
    int x_bytes;        // max distinct column blocks per chunk * kTmaTile
    int blk_bytes;      // max matrix blocks per chunk * 80
    int p_bytes;        // kTmaRows * 72, rounded to 16
    __host__ __device__ int stage_bytes() const { return x_bytes + blk_bytes + p_bytes; }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// chunk metadata (device arrays, built on the host):
//   ccolptr[nchunks+1], ccols[]    distinct global block columns of each chunk
//   browptr[nbr+1], lblocks[]      BSR blocks with the 10th double holding the SLOT in ccols of the chunk;
//                                  the diagonal block is the first block of every row
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// kTmaRows consumer warps (one block row of the current chunk each) + one producer warp that
// runs ahead through the stage ring: full[s] counts the bytes of a stage, empty[s] the
// consumer warps that are done with it.  Consumer warps never synchronise with each other.
template <int STAGES>
__global__ void __launch_bounds__(kTmaThreads + 32 * kTmaProducers, 1)
bsr3_spmm_tma_kernel(const int* __restrict__ browptr, const double* __restrict__ lblocks,
                     const int* __restrict__ ccolptr, const int* __restrict__ ccols,
                     const double* __restrict__ X, double* __restrict__ Y, int nbr, int nchunks,
                     const double* __restrict__ R, const double* __restrict__ Pinv, TmaStageLayout lay,
                     CgScalars* __restrict__ sc, double* __restrict__ partials, unsigned* __restrict__ counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    if (sc->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = nchunks / gridDim.x, rem = nchunks % gridDim.x;
    const int c_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int c_count = per + ((int)blockIdx.x < rem ? 1 : 0);
    const int stage_bytes = lay.stage_bytes();

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 32 * kTmaProducers); mbar_init(&empty[s], kTmaRows); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double2 dot[3][1] = {{make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0)}};

    if (warp >= kTmaRows) {
        const int pw = warp - kTmaRows;
        // ---------------- producer warp: 16-byte async copies (LDGSTS), 512 B per warp instruction.
        // (Bulk-TMA copies were tried first: at 1.5 KB per gathered tile the per-request cost of the
        // TMA unit, ~40 requests per chunk, made the stage fill slower than the FMA loop.)
        for (int it = 0; it < c_count; ++it) {
            const int s = it % STAGES;
            const int c = c_begin + it;
            const int r0 = c * kTmaRows, r1 = min(r0 + kTmaRows, nbr);
            const int cc0 = __ldg(ccolptr + c), cc1 = __ldg(ccolptr + c + 1);
            const int kb0 = __ldg(browptr + r0), kb1 = __ldg(browptr + r1);
            const int ncol = cc1 - cc0;
            const int col_a = (lane < ncol) ? __ldg(ccols + cc0 + lane) : 0;
            const int col_b = (32 + lane < ncol) ? __ldg(ccols + cc0 + 32 + lane) : 0;
            if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
            unsigned char* st = smem + (size_t)s * stage_bytes;
            for (int j = pw; j < ncol; j += kTmaProducers) {
                const int cj = (j < 32) ? __shfl_sync(0xffffffffu, col_a, j) : ((j < 64) ? __shfl_sync(0xffffffffu, col_b, j - 32) : __ldg(ccols + cc0 + j));
                const unsigned char* src = reinterpret_cast<const unsigned char*>(X + (size_t)(3 * cj) * kTmaW) + 16 * lane;
                unsigned char* dst = st + (size_t)j * kTmaTile + 16 * lane;
#pragma unroll
                for (int r = 0; r < 3; ++r) cp_async16(dst + 512 * r, src + 512 * r);
            }
            {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(lblocks + (size_t)10 * kb0);
                unsigned char* dst = st + lay.x_bytes;
                for (int e = 16 * lane + 512 * pw; e < (kb1 - kb0) * 80; e += 512 * kTmaProducers) cp_async16(dst + e, src + e);
                const unsigned char* psrc = reinterpret_cast<const unsigned char*>(Pinv + (size_t)9 * r0);
                unsigned char* pdst = st + lay.x_bytes + lay.blk_bytes;
                for (int e = 16 * lane + 512 * pw; e < (r1 - r0) * 72; e += 512 * kTmaProducers) cp_async16(pdst + e, psrc + e);
            }
            cp_async_arrive(&full[s]);      // this lane's copies of the stage: arrives when they have landed
        }
    } else {
        // ---------------- consumer warps
        double2 rv_next[3];
        int k0_next = 0, k1_next = 0, kb0_next = 0;
        auto prefetch = [&](int it) {
            const int br = (c_begin + it) * kTmaRows + warp;
            if (it < c_count && br < nbr) {
                k0_next = __ldg(browptr + br); k1_next = __ldg(browptr + br + 1);
                kb0_next = __ldg(browptr + (c_begin + it) * kTmaRows);
                const size_t o = (size_t)(3 * br) * kTmaW + 2 * lane;
#pragma unroll
                for (int i = 0; i < 3; ++i) rv_next[i] = ld2_stream(R + o + (size_t)i * kTmaW);
            }
        };
        prefetch(0);
        for (int it = 0; it < c_count; ++it) {
            const int s = it % STAGES;
            const int br = (c_begin + it) * kTmaRows + warp;
            const int k0 = k0_next, k1 = k1_next, kb0 = kb0_next;
            double2 rv[3] = {rv_next[0], rv_next[1], rv_next[2]};
            prefetch(it + 1);
            const unsigned char* st = smem + (size_t)s * stage_bytes;
            mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
            if (br < nbr) {
                const double2* sb = reinterpret_cast<const double2*>(st + lay.x_bytes) + (size_t)5 * (k0 - kb0);
                const double* sP = reinterpret_cast<const double*>(st + lay.x_bytes + lay.blk_bytes) + 9 * warp;
                double2 acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
                double2 pv[3];
                for (int k = 0; k < k1 - k0; ++k) {
                    const double2 v01 = sb[5 * k], v23 = sb[5 * k + 1], v45 = sb[5 * k + 2], v67 = sb[5 * k + 3], v8c = sb[5 * k + 4];
                    const int slot = __double2loint(v8c.y);
                    const double2* xt = reinterpret_cast<const double2*>(st + (size_t)slot * kTmaTile) + lane;
                    const double2 x0 = xt[0], x1 = xt[kTmaW / 2], x2 = xt[kTmaW];
                    if (k == 0) { pv[0] = x0; pv[1] = x1; pv[2] = x2; }      // the diagonal block comes first
                    acc[0].x = fma(v01.x, x0.x, fma(v01.y, x1.x, fma(v23.x, x2.x, acc[0].x)));
                    acc[0].y = fma(v01.x, x0.y, fma(v01.y, x1.y, fma(v23.x, x2.y, acc[0].y)));
                    acc[1].x = fma(v23.y, x0.x, fma(v45.x, x1.x, fma(v45.y, x2.x, acc[1].x)));
                    acc[1].y = fma(v23.y, x0.y, fma(v45.x, x1.y, fma(v45.y, x2.y, acc[1].y)));
                    acc[2].x = fma(v67.x, x0.x, fma(v67.y, x1.x, fma(v8c.x, x2.x, acc[2].x)));
                    acc[2].y = fma(v67.x, x0.y, fma(v67.y, x1.y, fma(v8c.x, x2.y, acc[2].y)));
                }
                double Pb[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) Pb[e] = sP[e];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);          // this warp no longer reads the stage
                const size_t o = (size_t)(3 * br) * kTmaW + 2 * lane;
#pragma unroll
                for (int i = 0; i < 3; ++i) st2(Y + o + (size_t)i * kTmaW, acc[i]);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double2 Pq;
                    Pq.x = fma(Pb[3 * i], acc[0].x, fma(Pb[3 * i + 1], acc[1].x, Pb[3 * i + 2] * acc[2].x));
                    Pq.y = fma(Pb[3 * i], acc[0].y, fma(Pb[3 * i + 1], acc[1].y, Pb[3 * i + 2] * acc[2].y));
                    dot[0][0].x = fma(pv[i].x, acc[i].x, dot[0][0].x); dot[0][0].y = fma(pv[i].y, acc[i].y, dot[0][0].y);
                    dot[1][0].x = fma(Pq.x, rv[i].x, dot[1][0].x);     dot[1][0].y = fma(Pq.y, rv[i].y, dot[1][0].y);
                    dot[2][0].x = fma(Pq.x, acc[i].x, dot[2][0].x);    dot[2][0].y = fma(Pq.y, acc[i].y, dot[2][0].y);
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        }
    }
    __syncthreads();           // all stages drained: shared memory is reused by the reduction

    // grid-wide reduction of the three dot products (two ticketed levels, fixed order); the
    // 16-warp CTA reduces in two halves of 8 warps through the same helper layout
    {
        double2 (*sh)[kTmaRows][kTmaW / 2] = reinterpret_cast<double2 (*)[kTmaRows][kTmaW / 2]>(smem);   // stages are drained
        __shared__ bool s_flag;
        const int G = gridDim.x, NG = (G + kGroup - 1) / kGroup;
        const int grp = blockIdx.x / kGroup;
        const int gsize = min(kGroup, G - grp * kGroup);
        double* gpart = partials + (size_t)3 * G * kTmaW;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (warp < kTmaRows) sh[a][warp][lane] = dot[a][0];
        __syncthreads();
        for (int e = threadIdx.x; e < 3 * (kTmaW / 2); e += kTmaThreads + 32 * kTmaProducers) {
            const int a = e / (kTmaW / 2), cp = e % (kTmaW / 2);
            double2 s = sh[a][0][cp];
#pragma unroll
            for (int w = 1; w < kTmaRows; ++w) { s.x += sh[a][w][cp].x; s.y += sh[a][w][cp].y; }
            st2(partials + ((size_t)a * G + blockIdx.x) * kTmaW + 2 * cp, s);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(counter + 1 + grp, 1u);
            s_flag = (ticket == (unsigned)gsize - 1u);
            if (s_flag) counter[1 + grp] = 0;
        }
        __syncthreads();
        if (!s_flag) return;
        __threadfence();
        for (int e = threadIdx.x; e < 3 * kTmaW; e += kTmaThreads + 32 * kTmaProducers) {
            const int a = e / kTmaW, c = e % kTmaW;
            const double* base = partials + ((size_t)a * G + (size_t)grp * kGroup) * kTmaW + c;
            double v[kGroup];
#pragma unroll
            for (int u = 0; u < kGroup; ++u) v[u] = (u < gsize) ? __ldcg(base + (size_t)u * kTmaW) : 0.0;
            double s = v[0];
#pragma unroll
            for (int u = 1; u < kGroup; ++u) s += v[u];
            gpart[((size_t)a * NG + grp) * kTmaW + c] = s;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(counter, 1u);
            s_flag = (ticket == (unsigned)NG - 1u);
            if (s_flag) counter[0] = 0;
        }
        __syncthreads();
        if (!s_flag) return;
        __threadfence();
        // final sums: thread c < 64 adds the NG group partials of column c in order, 3 accumulators
        if (threadIdx.x < kTmaW) {
            const int c = threadIdx.x;
            double s[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double t = 0.0;
                const double* base = gpart + (size_t)a * NG * kTmaW + c;
                for (int b = 0; b < NG; ++b) t += __ldcg(base + (size_t)b * kTmaW);
                s[a] = t;
            }
            const double rz = sc->rz[c];
            double alpha = 0.0, beta = 0.0;
            if (sc->active[c] && s[0] > 0.0 && rz > 0.0) {
                alpha = rz / s[0];
                const double rz_new = fma(alpha * alpha, s[2], fma(-2.0 * alpha, s[1], rz));
                beta = fmax(rz_new, 0.0) / rz;
            }
            sc->pq[c] = s[0];
            sc->alpha[c] = alpha;
            sc->beta[c] = beta;
        }
    }
}

}  // namespace kexp
