// rn_loss.cu -- sigmoid-focal + smooth-L1 loss, forward AND backward in one streaming pass.
//
// Replaces (reference file:line): ssd1 Vision.py:1568-1605 (one-hot build, gathers), focal_loss_retina
// Vision.py:1513-1530, smoothL1_loss_retina Vision.py:1532-1566, SSD_loss.__call__ Vision.py:1620-1644
// and the autograd replay of all of it (General/Learner.py:514).
//
// Data movement: clas [B,A,C] is read once and dclas written once with 128-bit accesses in a flat,
// perfectly coalesced mapping (a row of C floats is C/4 vectors; the row index is recovered with a
// compile-time-constant division); reg is only read for positive anchors; dreg is written once.
// Algorithmic bytes per image: 8*A*(C+4) with gradients, 4*A*(C+4) forward only.  The kernel is
// HBM-bound by design but sits close to the fp32 issue limit (~28 instructions per class element), so
// the logarithm is a branch-free polynomial (no special cases are reachable: its argument is clamped
// to [1e-4, 1-1e-4]) and the one division per element is MUFU.RCP + FMUL.  No tensor cores: there is no
// contraction anywhere on this path.
//
// Reductions: per-thread fp32 partial sums over <= 32 elements, fixed-order warp/block trees, one
// partial per CTA, then a one-CTA final kernel (one warp per image) that sums each image's partials in a
// fixed order in float64 and combines the images in image order.  No floating-point atomics anywhere
// => run-to-run bit-identical results.
#include <string.h>

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

// In-place scale by a device scalar; exits at once when the scalar is exactly 1.
__global__ void __launch_bounds__(RN_THREADS)
rn_scale_kernel(float *__restrict__ a, size_t na, float *__restrict__ bptr, size_t nb, const float *__restrict__ s) {
    const float k = *s;
    if (k == 1.0f) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 *a4 = reinterpret_cast<float4 *>(a);
    for (size_t i = i0; i < na / 4; i += stride) {
        float4 v = a4[i];
        v.x *= k; v.y *= k; v.z *= k; v.w *= k;
        a4[i] = v;
    }
    for (size_t i = (na / 4) * 4 + i0; i < na; i += stride) a[i] *= k;
    float4 *b4 = reinterpret_cast<float4 *>(bptr);
    for (size_t i = i0; i < nb / 4; i += stride) {
        float4 v = b4[i];
        v.x *= k; v.y *= k; v.z *= k; v.w *= k;
        b4[i] = v;
    }
    for (size_t i = (nb / 4) * 4 + i0; i < nb; i += stride) bptr[i] *= k;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// Sub-tiles per CTA: as many as possible (amortises the CTA prologue / reduction) while the grid still
// has >= ~4 waves of 148 SMs x 3 resident CTAs (measured: Pascal B=32 is best with 2, COCO B=16 with 4).
static int rn_loss_iters(int B, int A, int C) {
    const int V = (C % 4 == 0) ? 4 : 1;
    const long long sub = ((long long)A * (C / V) + RN_LOSS_TILE - 1) / RN_LOSS_TILE;
    int it = 4;
    while (it > 1 && (long long)B * ((sub + it - 1) / it) < 4LL * 148 * 3) it >>= 1;
    if (rn_opt(RN_OPT_LOSS_ITERS) > 0) it = rn_opt(RN_OPT_LOSS_ITERS);  // tuning override (rn_set_option)
    return it;
}
static int rn_loss_tiles(int B, int A, int C) {  // CTAs (= partials) per image
    const int V = (C % 4 == 0) ? 4 : 1;
    const long long sub = ((long long)A * (C / V) + RN_LOSS_TILE - 1) / RN_LOSS_TILE;
    const int it = rn_loss_iters(B, A, C);
    return (int)((sub + it - 1) / it);
}

extern "C" size_t rn_loss_workspace_bytes(int B, int A, int C) {
    if (B <= 0 || A <= 0 || C <= 0) return 256;
    // worst case over the tiling choice (iters = 1), so the size does not depend on the heuristic
    const int V = (C % 4 == 0) ? 4 : 1;
    const size_t sub = (size_t)(((long long)A * (C / V) + RN_LOSS_TILE - 1) / RN_LOSS_TILE);
    size_t partials = sizeof(float2) * (size_t)B * sub;
    size_t per_image = sizeof(float) * 2 * (size_t)B;
    return ((partials + 255) / 256) * 256 + ((per_image + 255) / 256) * 256;
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

template <typename MT>
static void rn_dispatch_loss(bool logits, int V, int C, bool g2, bool grad, dim3 grid, size_t smem, cudaStream_t s,
                             const RnLossParams &P, const RnGeom &g) {
    if (logits) {
        if (V == 4 && C == 80) rn_launch_loss<4, 20, true, MT>(g2, grad, grid, smem, s, P, g);
        else if (V == 4 && C == 20) rn_launch_loss<4, 5, true, MT>(g2, grad, grid, smem, s, P, g);
        else if (V == 4) rn_launch_loss<4, 0, true, MT>(g2, grad, grid, smem, s, P, g);
        else rn_launch_loss<1, 0, true, MT>(g2, grad, grid, smem, s, P, g);
    } else {
        if (V == 4 && C == 80) rn_launch_loss<4, 20, false, MT>(g2, grad, grid, smem, s, P, g);
        else if (V == 4 && C == 20) rn_launch_loss<4, 5, false, MT>(g2, grad, grid, smem, s, P, g);
        else if (V == 4) rn_launch_loss<4, 0, false, MT>(g2, grad, grid, smem, s, P, g);
        else rn_launch_loss<1, 0, false, MT>(g2, grad, grid, smem, s, P, g);
    }
}

extern "C" int rn_loss(const float *clas, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
                       const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
                       const double *base, int K, const float *anchors, double alpha, double gamma, double beta,
                       int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                       size_t workspace_bytes, void *stream) {
    return rn_loss_impl(false, nullptr, clas, reg, gt_boxes, gt_cats, matches, npos, nullptr, B, A, C, M, H, W, base, K, anchors,
                        alpha, gamma, beta, B_global, dclas, dreg, out3, workspace, workspace_bytes, stream);
}

extern "C" int rn_loss_logits(const float *logits, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
                              const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
                              const double *base, int K, const float *anchors, double alpha, double gamma,
                              double beta, int B_global, float *dlogits, float *dreg, float *probs_out, float *out3,
                              void *workspace, size_t workspace_bytes, void *stream) {
    return rn_loss_impl(true, probs_out, logits, reg, gt_boxes, gt_cats, matches, npos, nullptr, B, A, C, M, H, W, base, K, anchors,
                        alpha, gamma, beta, B_global, dlogits, dreg, out3, workspace, workspace_bytes, stream);
}

// bytes != NULL: the assignment is the byte map of rn_loss_step (matches is then the optional dense OUTPUT, npos the
// persistent counters); the final kernel also restores the map and the counters (rn_step.cu).
int rn_loss_impl(bool logits, float *probs, const float *clas, const float *reg, const float *gt_boxes,
                 const int64_t *gt_cats, const int32_t *matches, const int32_t *npos, const RnLossBytes *bytes, int B, int A,
                 int C, int M, int H, int W, const double *base, int K, const float *anchors, double alpha, double gamma,
                 double beta, int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                 size_t workspace_bytes, void *stream) {
    if (B <= 0 || A <= 0 || C <= 0 || M < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: B=%d A=%d C=%d M=%d", B, A, C, M);
    if (!clas || !reg || (!matches && !bytes) || !npos || !out3 || (M > 0 && (!gt_boxes || !gt_cats)))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: null pointer");
    if ((dclas == nullptr) != (dreg == nullptr))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: dclas and dreg must both be given or both be NULL");
    if (B_global < B) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: B_global=%d < B=%d", B_global, B);
    const int V = (C % 4 == 0) ? 4 : 1;
    if ((long long)A * (C / V) > 0x7fffffffLL - RN_LOSS_TILE) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: A*C too large");
    if (V == 4 && ((((uintptr_t)clas) | ((uintptr_t)dclas)) & 15))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: clas/dclas must be 16-byte aligned");
    if ((((uintptr_t)reg) | ((uintptr_t)dreg) | ((uintptr_t)gt_boxes) | ((uintptr_t)anchors)) & 15)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: reg/dreg/gt_boxes/anchors must be 16-byte aligned");
    if (workspace_bytes < rn_loss_workspace_bytes(B, A, C) || !workspace || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_loss: workspace needs %zu bytes, 256-byte aligned", rn_loss_workspace_bytes(B, A, C));
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, anchors, A);
    if (rc) return rc;

    const int tiles = rn_loss_tiles(B, A, C);
    RnLossParams P;
    P.clas = clas; P.reg = reg;
    P.gt_boxes = reinterpret_cast<const float4 *>(gt_boxes); P.gt_cats = gt_cats;
    P.matches = bytes ? nullptr : matches; P.m8 = bytes ? bytes->m8 : nullptr;
    P.matches_out = bytes ? const_cast<int32_t *>(matches) : nullptr;
    P.npos = npos; P.table = reinterpret_cast<const float4 *>(anchors);
    P.dclas = dclas; P.dreg = dreg; P.probs = probs;
    P.B = B; P.A = A; P.C = C; P.CV = C / V; P.M = M; P.tiles = tiles; P.iters = rn_loss_iters(B, A, C);
    P.prefetch = rn_opt(RN_OPT_LOSS_PREFETCH) > 0 ? rn_opt(RN_OPT_LOSS_PREFETCH) : 2;
    P.resident = 148 * RN_LOSS_CTAS;
    P.a_pos = (float)alpha; P.a_neg = (float)(1.0 - alpha);  // Vision.py:1526
    P.gamma = (float)gamma;
    P.lo = (float)1e-4; P.hi = (float)(1.0 - 1e-4);          // Vision.py:1524
    const float bs = (float)B_global;
    const float w_reg = (float)(1.0 - beta), w_clas = (float)beta;  // Vision.py:1644
    P.wc_over_bs = w_clas / bs;
    P.wr_over_bs = w_reg / bs;
    unsigned char *wsb = reinterpret_cast<unsigned char *>(workspace);
    const size_t ws_total = rn_loss_workspace_bytes(B, A, C);
    P.partials = reinterpret_cast<float *>(wsb);
    float *per_image = reinterpret_cast<float *>(wsb + ws_total - ((sizeof(float) * 2 * (size_t)B + 255) / 256) * 256);

    const size_t smem = (size_t)M * (sizeof(float4) + sizeof(int));
    if (smem > 48 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: M=%d too large", M);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(tiles, B);
    const bool g2 = (gamma == 2.0), grad = dclas != nullptr;
    if (bytes) rn_dispatch_loss<RnMatchU8NC>(logits, V, C, g2, grad, grid, smem, s, P, g);
    else rn_dispatch_loss<RnMatchI32>(logits, V, C, g2, grad, grid, smem, s, P, g);
    rc = rn_check_launch("rn_loss");
    if (rc) return rc;
    RnFinalClean clean;
    memset(&clean, 0, sizeof(clean));
    if (bytes) {
        clean.m8 = bytes->m8; clean.clean_list = bytes->clean_list; clean.clean_cnt = bytes->clean_cnt;
        clean.npos_acc = bytes->npos_acc; clean.npos_out = bytes->npos_out; clean.A = A;
    }
    rn_launch_pdl(rn_loss_final_kernel, dim3(1), dim3(1024), 0, s, reinterpret_cast<const float2 *>(P.partials), npos, B,
                  tiles, w_reg, w_clas, bs, per_image, out3, clean);
    return rn_check_launch("rn_loss_final");
}

// ------------------------------------------------------------------------------------------------
// The one exchange of the multi-GPU path: every rank's three loss scalars summed over the image
// shards -- 12 bytes per rank, pure latency.  Instead of a NCCL all-gather on a side stream (an event pair and a collective
// launch per step) one tiny kernel on the SAME stream as the loss, capturable in the step's CUDA graph: the rank stores its
// scalars into its slot of every peer's buffer through peer-mapped (symmetric) memory over NVLink, publishes them with a
// release store of the step's sequence number, waits for the peers' numbers with acquire loads, and sums the slots in rank
// order -- the same fixed order on every rank, so the result is bit-identical everywhere.
// Buffer of rank r (peer-mapped into every process):  float slots[2][world][4] | uint32 flags[2][world].  Two slot sets,
// indexed by the parity of the sequence number: a rank can be at most one step ahead of the slowest one (its step k+1
// cannot finish before every peer has published step k+1, i.e. has finished reading step k), so a set is never overwritten
// while a peer still reads it.
// ------------------------------------------------------------------------------------------------
#define RN_PEER_MAX 16
struct RnPeers {
    float *buf[RN_PEER_MAX];
};

__global__ void __launch_bounds__(32)
rn_peer_exchange_kernel(const float *in3, float *out3, const RnPeers peers, int rank, int world, uint32_t *__restrict__ seq) {
    const int t = threadIdx.x;
    const uint32_t e = *seq + 1u;
    const int set = (int)(e & 1u);
    const float v0 = in3[0], v1 = in3[1], v2 = in3[2];  // in3 may equal out3 (rn_peer_exchange)
    bool timed_out = false;
    if (t < world) {
        float *slot = peers.buf[t] + ((size_t)set * world + rank) * 4;
        slot[0] = v0;
        slot[1] = v1;
        slot[2] = v2;
        __threadfence_system();
        uint32_t *flag = reinterpret_cast<uint32_t *>(peers.buf[t] + (size_t)2 * world * 4) + (size_t)set * world + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(e) : "memory");
        // wait until rank t has published this step into MY buffer -- but never forever: a peer that died (or never launched
        // its step) must not hang the GPU; after ~5 s the exchange gives up and poisons the result
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(peers.buf[rank] + (size_t)2 * world * 4) + (size_t)set * world + t;
        uint32_t got;
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
            if ((int)(got - e) >= 0) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 5000000000ull) {
                timed_out = true;
                break;
            }
            __nanosleep(100);
        } while (true);
    }
    timed_out = __any_sync(RN_FULL_MASK, timed_out);
    if (t == 0 && timed_out) {
        out3[0] = out3[1] = out3[2] = __int_as_float(0x7fc00000);  // NaN: loud, and the next step is not blocked by this one
        *seq = e;
        return;
    }
    if (timed_out) return;
    if (t == 0) {
        const float *slots = peers.buf[rank] + (size_t)set * world * 4;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int r = 0; r < world; ++r) {  // rank order: the same on every rank
            s0 = __fadd_rn(s0, slots[4 * r + 0]);
            s1 = __fadd_rn(s1, slots[4 * r + 1]);
            s2 = __fadd_rn(s2, slots[4 * r + 2]);
        }
        out3[0] = s0;
        out3[1] = s1;
        out3[2] = s2;
        *seq = e;
    }
}

extern "C" size_t rn_peer_exchange_bytes(int world) {
    if (world < 1) return 0;
    return (size_t)2 * world * 4 * sizeof(float) + (size_t)2 * world * sizeof(uint32_t);
}

extern "C" int rn_peer_exchange_to(const float *in3, float *out3, void *const *peer_bufs, int rank, int world, uint32_t *seq,
                                   void *stream) {
    if (world < 1 || world > RN_PEER_MAX || rank < 0 || rank >= world)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_peer_exchange: rank=%d world=%d (at most %d ranks)", rank, world, RN_PEER_MAX);
    if (!in3 || !out3 || !peer_bufs || !seq) return rn_set_error(RN_ERR_INVALID_ARG, "rn_peer_exchange: null pointer");
    RnPeers P;
    memset(&P, 0, sizeof(P));
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || (((uintptr_t)peer_bufs[r]) & 15)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_peer_exchange: bad buffer of rank %d", r);
        P.buf[r] = reinterpret_cast<float *>(peer_bufs[r]);
    }
    rn_peer_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(in3, out3, P, rank, world, seq);
    return rn_check_launch("rn_peer_exchange");
}

extern "C" int rn_peer_exchange(float *out3, void *const *peer_bufs, int rank, int world, uint32_t *seq, void *stream) {
    return rn_peer_exchange_to(out3, out3, peer_bufs, rank, world, seq, stream);
}

extern "C" int rn_scale_grads(float *dclas, size_t n_clas, float *dreg, size_t n_reg, const float *grad_out,
                              void *stream) {
    if (!grad_out || (n_clas && !dclas) || (n_reg && !dreg)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_scale_grads: null pointer");
    if ((((uintptr_t)dclas) | ((uintptr_t)dreg)) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_scale_grads: 16-byte alignment required");
    rn_scale_kernel<<<148 * 8, RN_THREADS, 0, (cudaStream_t)stream>>>(dclas, n_clas, dreg, n_reg, grad_out);
    return rn_check_launch("rn_scale_grads");
}
