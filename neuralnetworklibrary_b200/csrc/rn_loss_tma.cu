// rn_loss_tma.cu -- the flat [B,A,C] loss kernel of rn_loss.cu with `clas` staged through shared memory by bulk
// asynchronous copies (cp.async.bulk + mbarrier, the TMA engine in its 1-D form) instead of through registers.
//
// Why: in rn_loss_kernel the bytes in flight live in registers (8 x 128-bit loads per thread = 32 of its 80 registers),
// which caps the kernel at three CTAs per SM and serialises "wait for loads" and "math" inside every warp.  Here one
// elected thread keeps RN_TMA_STAGES sub-tiles of 32 KB in flight per CTA, independent of what the warps are doing, and
// the warps read their vectors from shared memory one at a time.  A sub-tile of the flat layout is one contiguous 32 KB
// run, i.e. exactly one bulk copy.  Same element math (rn_loss_vector4), same partial sums, same final reduction.
// Experimental: selected with the environment variable RN_LOSS_TMA (profiles/r01_summary.md has the A/B numbers).
#include <stdlib.h>

#include "rn_loss_math.cuh"

#ifndef RN_TMA_U
#define RN_TMA_U 8
#endif
#define RN_TMA_TILE (RN_THREADS * RN_TMA_U)  // vectors per sub-tile (32 KB)
#ifndef RN_TMA_STAGES
#define RN_TMA_STAGES 2
#endif
#ifndef RN_TMA_CTAS
#define RN_TMA_CTAS 3
#endif

__device__ __forceinline__ uint32_t rn_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rn_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rn_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void rn_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rn_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rn_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rn_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(rn_smem_addr(bar))
                 : "memory");
}
// Bounded wait: a mistake in the pipeline must not hang the GPU (results would be wrong and the tests would say so).
__device__ __forceinline__ void rn_mbar_wait(uint64_t *bar, unsigned parity) {
    const uint32_t a = rn_smem_addr(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (done) return;
    }
}

template <int CVT, bool G2, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(RN_THREADS, RN_TMA_CTAS)
rn_loss_tma_kernel(const __grid_constant__ RnLossParams P, const __grid_constant__ RnGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: stages x 32 KB of class vectors | gt boxes float4[M] | gt cats int[M]
    float4 *s_x = reinterpret_cast<float4 *>(smem);
    float4 *s_box = s_x + RN_TMA_STAGES * RN_TMA_TILE;
    int *s_cat = reinterpret_cast<int *>(s_box + P.M);
    __shared__ __align__(8) uint64_t s_bar[RN_TMA_STAGES];
    __shared__ float s_red[2][RN_THREADS / 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int A = P.A;
    const int CV = CVT ? CVT : P.CV;
    const int nvec = A * CV;
    const int span = RN_THREADS * 8 * P.iters;  // the host plans CTAs in units of 2048 vectors (rn_loss.cu)
    const int cta0 = blockIdx.x * span;
    const int cta1 = min(nvec, cta0 + span);
    const int ntiles = (cta1 - cta0 + RN_TMA_TILE - 1) / RN_TMA_TILE;
    const float4 *x_img = reinterpret_cast<const float4 *>(P.clas) + (size_t)b * nvec;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RN_TMA_STAGES; ++s) rn_mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int i) {  // sub-tile i of this CTA -> stage i % STAGES (one elected thread)
        const int t0 = cta0 + i * RN_TMA_TILE;
        const unsigned bytes = (unsigned)(min(RN_TMA_TILE, cta1 - t0)) * 16u;
        uint64_t *bar = &s_bar[i % RN_TMA_STAGES];
        rn_mbar_expect_tx(bar, bytes);
        rn_bulk_g2s(s_x + (i % RN_TMA_STAGES) * RN_TMA_TILE, x_img + t0, bytes, bar);
    };
    if (tid == 0) {  // `clas` does not depend on the assignment kernel this launch overlaps with (PDL)
        for (int i = 0; i < RN_TMA_STAGES && i < ntiles; ++i) issue(i);
    }

    if (tid < 32) rn_compact_gt(P.gt_boxes + (size_t)b * P.M, P.gt_cats + (size_t)b * P.M, P.M, s_box, nullptr, s_cat);
    rn_pdl_wait();
    const int n_pos = P.npos[b];
    const float n_norm = fmaxf((float)n_pos, 1.0f);    // clamp(min=1), Vision.py:1530
    const float gl = __fdiv_rn(P.wc_over_bs, n_norm);  // upstream of every focal term
    float4 *dx_img = GRAD ? reinterpret_cast<float4 *>(P.dclas) + (size_t)b * nvec : nullptr;
    float4 *pr_img = (LOGITS && P.probs) ? reinterpret_cast<float4 *>(P.probs) + (size_t)b * nvec : nullptr;
    const int32_t *m_img = P.matches + (size_t)b * A;
    __syncthreads();  // s_cat / s_box visible

    float acc_neg = 0.0f, acc_pos = 0.0f;
#pragma unroll 1
    for (int i = 0; i < ntiles; ++i) {
        const int stage = i % RN_TMA_STAGES;
        const int t0 = cta0 + i * RN_TMA_TILE;
        int mrow[RN_TMA_U];
#pragma unroll
        for (int u = 0; u < RN_TMA_U; ++u) {  // the rows' assignments (L2 hits), requested before the wait
            const int v = t0 + tid + u * RN_THREADS;
            mrow[u] = (v < cta1) ? __ldg(m_img + v / CV) : RN_MATCH_IGNORE;
        }
        rn_mbar_wait(&s_bar[stage], (unsigned)((i / RN_TMA_STAGES) & 1));
        const float4 *sx = s_x + stage * RN_TMA_TILE;
#pragma unroll
        for (int u = 0; u < RN_TMA_U; ++u) {
            const int v = t0 + tid + u * RN_THREADS;
            if (v < cta1) {
                const float4 gv = rn_loss_vector4<CVT, G2, GRAD, LOGITS>(P, sx[tid + u * RN_THREADS], mrow[u], v, CV, s_cat, gl,
                                                                         pr_img ? reinterpret_cast<float *>(pr_img + v) : nullptr,
                                                                         acc_neg, acc_pos);
                if (GRAD) rn_stg_stream(dx_img + v, gv);
            }
        }
        __syncthreads();  // every thread is done reading this stage
        if (tid == 0 && i + RN_TMA_STAGES < ntiles) issue(i + RN_TMA_STAGES);
    }

    // ---- regression rows whose first vector lies in this CTA's range: smooth L1 (Vision.py:1532-1566) ----
    float acc_reg = 0.0f;
    {
        const int r0 = (cta0 + CV - 1) / CV;
        const int r1 = min(A, (cta1 + CV - 1) / CV);
        const float numel = (float)(4 * n_pos);
        const float ge = n_pos > 0 ? __fdiv_rn(P.wr_over_bs, numel) : 0.0f;  // mean() backward
        const float4 *reg4 = reinterpret_cast<const float4 *>(P.reg) + (size_t)b * A;
        float4 *dreg4 = GRAD ? reinterpret_cast<float4 *>(P.dreg) + (size_t)b * A : nullptr;
        const float knee = (float)(1.0 / 9.0), off = (float)(0.5 / 9.0);  // Vision.py:1565
        for (int row = r0 + tid; row < r1; row += RN_THREADS) {
            const int m = __ldg(m_img + row);
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) {
                const float4 an = rn_anchor_from_param(g, P.table, row);
                const float4 tg = s_box[m];
                const float4 pr = __ldg(reg4 + row);
                const float aw = __fsub_rn(an.z, an.x), ah = __fsub_rn(an.w, an.y);
                const float acx = __fadd_rn(an.x, __fmul_rn(0.5f, aw)), acy = __fadd_rn(an.y, __fmul_rn(0.5f, ah));
                float tw = __fsub_rn(tg.z, tg.x), th = __fsub_rn(tg.w, tg.y);
                const float tcx = __fadd_rn(tg.x, __fmul_rn(0.5f, tw)), tcy = __fadd_rn(tg.y, __fmul_rn(0.5f, th));
                tw = fmaxf(tw, 1.0f);  // Vision.py:1553-1554
                th = fmaxf(th, 1.0f);
                float ts[4], pv[4] = {pr.x, pr.y, pr.z, pr.w}, gg[4];
                ts[0] = __fdiv_rn(__fdiv_rn(__fsub_rn(tcx, acx), aw), 0.1f);  // Vision.py:1556, :1562
                ts[1] = __fdiv_rn(__fdiv_rn(__fsub_rn(tcy, acy), ah), 0.1f);
                ts[2] = __fdiv_rn(logf(__fdiv_rn(tw, aw)), 0.2f);             // Vision.py:1558
                ts[3] = __fdiv_rn(logf(__fdiv_rn(th, ah)), 0.2f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = __fsub_rn(ts[k], pv[k]);
                    const float diff = fabsf(d);
                    float l, gd;
                    if (diff < knee) {
                        l = __fmul_rn(4.5f, __fmul_rn(diff, diff));
                        gd = __fmul_rn(__fmul_rn(ge, 4.5f), __fmul_rn(2.0f, diff));
                    } else {
                        l = __fsub_rn(diff, off);
                        gd = ge;
                    }
                    acc_reg += l;
                    gg[k] = d > 0.0f ? -gd : (d < 0.0f ? gd : 0.0f);  // -sign(t - p) * gd
                }
                g4 = make_float4(gg[0], gg[1], gg[2], gg[3]);
            }
            if (GRAD) dreg4[row] = g4;
        }
    }

    rn_pdl_trigger();
    float c = rn_warp_sum(acc_neg + acc_pos);
    float r = rn_warp_sum(acc_reg);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = c;
        s_red[1][tid >> 5] = r;
    }
    __syncthreads();
    if (tid == 0) {
        float cs = 0.f, rs = 0.f;
#pragma unroll
        for (int w = 0; w < RN_THREADS / 32; ++w) {
            cs += s_red[0][w];
            rs += s_red[1][w];
        }
        reinterpret_cast<float2 *>(P.partials)[(size_t)b * gridDim.x + blockIdx.x] = make_float2(cs, rs);
    }
}

template <int CVT, bool LOGITS>
static cudaError_t rn_launch_tma_t(bool g2, bool grad, dim3 grid, size_t smem, cudaStream_t s, const RnLossParams &P, const RnGeom &g) {
#define RN_TMA_GO(G2, GRAD)                                                                                                       \
    do {                                                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(rn_loss_tma_kernel<CVT, G2, GRAD, LOGITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem);                                                                          \
        if (e != cudaSuccess) return e;                                                                                           \
        return rn_launch_pdl(rn_loss_tma_kernel<CVT, G2, GRAD, LOGITS>, grid, dim3(RN_THREADS), smem, s, P, g);                    \
    } while (0)
    if (g2 && grad) RN_TMA_GO(true, true);
    else if (g2) RN_TMA_GO(true, false);
    else if (grad) RN_TMA_GO(false, true);
    else RN_TMA_GO(false, false);
#undef RN_TMA_GO
}

// Called by rn_loss_impl (rn_loss.cu) instead of its own launch when RN_LOSS_TMA is set and C % 4 == 0.  Returns false if
// this variant does not apply (the caller then launches the register-staged kernel).
bool rn_launch_loss_tma(bool logits, bool g2, bool grad, dim3 grid, cudaStream_t s, const RnLossParams &P, const RnGeom &g) {
    if (P.C % 4 != 0) return false;
    const size_t smem = (size_t)RN_TMA_STAGES * RN_TMA_TILE * sizeof(float4) + (size_t)P.M * (sizeof(float4) + sizeof(int));
    if (smem > 200 * 1024) return false;
    cudaError_t e;
    if (logits) {
        if (P.C == 80) e = rn_launch_tma_t<20, true>(g2, grad, grid, smem, s, P, g);
        else if (P.C == 20) e = rn_launch_tma_t<5, true>(g2, grad, grid, smem, s, P, g);
        else e = rn_launch_tma_t<0, true>(g2, grad, grid, smem, s, P, g);
    } else {
        if (P.C == 80) e = rn_launch_tma_t<20, false>(g2, grad, grid, smem, s, P, g);
        else if (P.C == 20) e = rn_launch_tma_t<5, false>(g2, grad, grid, smem, s, P, g);
        else e = rn_launch_tma_t<0, false>(g2, grad, grid, smem, s, P, g);
    }
    return e == cudaSuccess;
}
