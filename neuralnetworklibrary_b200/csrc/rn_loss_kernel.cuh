// rn_loss_kernel.cuh -- the flat [B,A,C] loss kernel template and its launch dispatch (see rn_loss.cu for the description of
// the path).  A header because the 40 instantiations are compiled in four translation units.
#pragma once
#include "rn_loss_math.cuh"

#ifndef RN_LOSS_CTAS
#define RN_LOSS_CTAS 3
#endif
#ifndef RN_LOSS_CTAS_LOGIT
#define RN_LOSS_CTAS_LOGIT 3
#endif

// V: floats per vector (4 when C % 4 == 0, else 1).  CVT: compile-time vectors per row (0 = runtime).
// Each CTA handles P.iters consecutive sub-tiles of one image (the prologue -- ground-truth compaction,
// per-image scalars -- and the block reduction are paid once per CTA).
// __launch_bounds__(256, 3): three resident CTAs per SM need <= 85 registers per thread.  Without the
// bound ptxas drifted from 80 to 88 registers after an unrelated parameter-struct change, dropping
// occupancy to two CTAs per SM and the kernel from 370 us to 419 us (profiles/r01_summary.md).
// MT: where the assignment comes from (RnMatchI32: rn_assign's int32 matches; RnMatchU8: the byte map of rn_loss_step).
template <int V, int CVT, bool G2, bool GRAD, bool LOGITS, typename MT>
__global__ void __launch_bounds__(RN_THREADS, LOGITS ? RN_LOSS_CTAS_LOGIT : RN_LOSS_CTAS)
rn_loss_kernel(const __grid_constant__ RnLossParams P, const __grid_constant__ RnGeom g) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: gt boxes float4[M] | gt cats int[M]
    float4 *s_box = reinterpret_cast<float4 *>(smem);
    int *s_cat = reinterpret_cast<int *>(s_box + P.M);
    __shared__ float s_red[2][RN_THREADS / 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int A = P.A;
    const int CV = CVT ? CVT : P.CV;
    const int nvec = A * CV;  // vectors in one image
    const int cta0 = blockIdx.x * (RN_LOSS_TILE * P.iters);
    const int cta1 = min(nvec, cta0 + RN_LOSS_TILE * P.iters);

    if (tid < 32) rn_compact_gt(P.gt_boxes + (size_t)b * P.M, P.gt_cats + (size_t)b * P.M, P.M, s_box, nullptr, s_cat);
    {   // Launched with PDL right behind rn_assign: while the assignment kernel (8 us of dependent latencies during which
        // HBM idles) drains, pull this CTA's sub-tiles of `clas` (independent of the assignment) towards L2 -- one 128-byte
        // line per thread and sub-tile -- then wait for its matches / npos.  P.prefetch: 1 = first sub-tile only,
        // 2 = all of them (default), 3 = also those of the CTA that will take this CTA's place (first wave only).
        const int linear = blockIdx.y * gridDim.x + blockIdx.x;
        const int rounds = (P.prefetch >= 3 && linear < P.resident) ? 2 : 1;
        for (int r = 0; r < rounds; ++r) {
            const int lin = linear + r * P.resident;
            const int pb = lin / (int)gridDim.x, px = lin - pb * (int)gridDim.x;
            if (pb >= P.B) break;
            const int p0 = px * (RN_LOSS_TILE * P.iters);
            const int nsub = P.prefetch >= 2 ? P.iters : 1;
            for (int it = 0; it < nsub; ++it) {
                const int v = p0 + it * RN_LOSS_TILE + tid * RN_LOSS_U;
                if (v < nvec) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.clas + ((size_t)pb * nvec + v) * V));
            }
        }
        rn_pdl_wait();
    }

    const int n_pos = P.npos[b];
    const float n_norm = fmaxf((float)n_pos, 1.0f);    // clamp(min=1), Vision.py:1530
    const float gl = __fdiv_rn(P.wc_over_bs, n_norm);  // upstream of every focal term
    const float *x_img = P.clas + (size_t)b * A * P.C;
    float *dx_img = GRAD ? P.dclas + (size_t)b * A * P.C : nullptr;
    const typename MT::T *m_img = MT::base(P) + (size_t)b * A;
    __syncthreads();  // s_cat / s_box visible

    float acc_neg = 0.0f, acc_pos = 0.0f;
#pragma unroll 1
    for (int tile0 = cta0; tile0 < cta1; tile0 += RN_LOSS_TILE) {
        if (tile0 + RN_LOSS_TILE <= nvec)
            rn_loss_subtile<V, CVT, G2, GRAD, true, LOGITS, MT>(P, b, x_img, dx_img, m_img, s_cat, CV, nvec, nvec, tile0, gl, acc_neg, acc_pos);
        else
            rn_loss_subtile<V, CVT, G2, GRAD, false, LOGITS, MT>(P, b, x_img, dx_img, m_img, s_cat, CV, nvec, nvec, tile0, gl, acc_neg, acc_pos);
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
        for (int row = r0 + tid; row < r1; row += RN_THREADS) {
            const int m = MT::load(m_img, row);
            if (sizeof(typename MT::T) == 1 && P.matches_out) P.matches_out[(size_t)b * A + row] = m;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) {
                g4 = rn_smooth_l1_row(rn_anchor_from_param(g, P.table, row), s_box[m], __ldg(reg4 + row), ge, acc_reg);
            }
            if (GRAD) dreg4[row] = g4;
        }
    }

    rn_pdl_trigger();  // the final-reduction kernel may be scheduled as the last CTAs retire
    // ---- block reduction (fixed order) -> one partial pair per CTA ----
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

template <int V, int CVT, bool LOGITS, typename MT>
static void rn_launch_loss(bool g2, bool grad, dim3 grid, size_t smem, cudaStream_t s, const RnLossParams &P,
                           const RnGeom &g) {
    // the training configuration (gamma == 2 with gradients) has row-width specialisations; the rarer variants (forward only,
    // general gamma) share the generic row width -- fewer instantiations, a smaller library
    if (g2 && grad) {
        rn_launch_pdl(rn_loss_kernel<V, CVT, true, true, LOGITS, MT>, grid, dim3(RN_THREADS), smem, s, P, g);
    } else if constexpr (CVT != 0) {
        rn_launch_loss<V, 0, LOGITS, MT>(g2, grad, grid, smem, s, P, g);
    } else {
        if (g2) rn_launch_pdl(rn_loss_kernel<V, 0, true, false, LOGITS, MT>, grid, dim3(RN_THREADS), smem, s, P, g);
        else if (grad) rn_launch_pdl(rn_loss_kernel<V, 0, false, true, LOGITS, MT>, grid, dim3(RN_THREADS), smem, s, P, g);
        else rn_launch_pdl(rn_loss_kernel<V, 0, false, false, LOGITS, MT>, grid, dim3(RN_THREADS), smem, s, P, g);
    }
}

// One quarter of the kernel family (10 instantiations): probabilities or logits x the source of the assignment.  Each
// quarter is instantiated in its own translation unit (rn_loss.cu, rn_loss_inst_*.cu) so that the four compile in parallel --
// ptxas needs ~30 s per quarter; rn_loss.cu declares the other three `extern`.
template <bool LOGITS, typename MT>
void rn_dispatch_loss_part(int V, int C, bool g2, bool grad, dim3 grid, size_t smem, cudaStream_t s, const RnLossParams &P,
                           const RnGeom &g) {
    if (V == 4 && C == 80) rn_launch_loss<4, 20, LOGITS, MT>(g2, grad, grid, smem, s, P, g);
    else if (V == 4 && C == 20) rn_launch_loss<4, 5, LOGITS, MT>(g2, grad, grid, smem, s, P, g);
    else if (V == 4) rn_launch_loss<4, 0, LOGITS, MT>(g2, grad, grid, smem, s, P, g);
    else rn_launch_loss<1, 0, LOGITS, MT>(g2, grad, grid, smem, s, P, g);
}
#define RN_LOSS_PART_ARGS int, int, bool, bool, dim3, size_t, cudaStream_t, const RnLossParams &, const RnGeom &

