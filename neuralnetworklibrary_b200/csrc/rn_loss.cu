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

#include "rn_loss_kernel.cuh"

// this translation unit instantiates <probabilities, int32 matches>; the other three quarters live in rn_loss_inst_*.cu
extern template void rn_dispatch_loss_part<true, RnMatchI32>(RN_LOSS_PART_ARGS);
#ifdef RN_EXPERIMENTAL
extern template void rn_dispatch_loss_part<false, RnMatchU8NC>(RN_LOSS_PART_ARGS);
extern template void rn_dispatch_loss_part<true, RnMatchU8NC>(RN_LOSS_PART_ARGS);
#endif


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
    if (bytes) {
#ifdef RN_EXPERIMENTAL
        if (logits) rn_dispatch_loss_part<true, RnMatchU8NC>(V, C, g2, grad, grid, smem, s, P, g);
        else rn_dispatch_loss_part<false, RnMatchU8NC>(V, C, g2, grad, grid, smem, s, P, g);
#else
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: the byte-map variant needs a library built with -DRN_EXPERIMENTAL");
#endif
    } else {
        if (logits) rn_dispatch_loss_part<true, RnMatchI32>(V, C, g2, grad, grid, smem, s, P, g);
        else rn_dispatch_loss_part<false, RnMatchI32>(V, C, g2, grad, grid, smem, s, P, g);
    }
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
