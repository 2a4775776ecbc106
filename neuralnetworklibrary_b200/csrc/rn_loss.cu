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
#include "rn_common.cuh"

#define RN_LOSS_U 8  // vectors per thread
#define RN_LOSS_TILE (RN_THREADS * RN_LOSS_U)

struct RnLossParams {
    const float *clas;
    const float *reg;
    const float4 *gt_boxes;
    const int64_t *gt_cats;
    const int32_t *matches;
    const int32_t *npos;
    const float4 *table;
    float *dclas;
    float *dreg;
    float *probs;     // LOGITS only, may be NULL: sigmoid(logits) as used by the kernel (for checking / reuse)
    float *partials;  // [B][tiles][2] : {sum of focal terms, sum of smooth-L1 terms}
    int B, A, C, CV, M, tiles, iters;
    float a_pos, a_neg, gamma, lo, hi;
    float wc_over_bs, wr_over_bs;  // beta / B_global, (1-beta) / B_global   (Vision.py:1644)
};

// -2*log(v) for v in [2^-20, 1]; max relative error 1.5e-7 (degree-6 minimax on [sqrt(.5), sqrt(2)),
// fitted for the relative error of log itself).  Branch free, no special cases.
__device__ __forceinline__ float rn_neg2log(float v) {
    const int i = __float_as_int(v);
    const int t = (i - 0x3f3504f3) & 0xff800000;  // exponent (as a float-field multiple of 2^23)
    const float f = __int_as_float(i - t) - 1.0f;  // mantissa in [sqrt(.5), sqrt(2)) minus 1
    const float e23 = (float)t;
    float p = -2.0f * 8.700362962e-02f;
    p = fmaf(p, f, -2.0f * -1.426749380e-01f);
    p = fmaf(p, f, -2.0f * 1.491478973e-01f);
    p = fmaf(p, f, -2.0f * -1.657758280e-01f);
    p = fmaf(p, f, -2.0f * 1.996306205e-01f);
    p = fmaf(p, f, -2.0f * -2.500133718e-01f);
    p = fmaf(p, f, -2.0f * 3.333391077e-01f);
    const float z = f * f;
    const float w0 = fmaf(-2.0f, f, z);  // -2*(f - f^2/2)
    const float r = fmaf(z * f, p, w0);
    return fmaf(e23, -2.0f * 0.69314718056f / 8388608.0f, r);
}

__device__ __forceinline__ float rn_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// LOGITS variant: the class activations are logits and the head's
// nn.Sigmoid (reference retinanet.py:258,286) is fused here, y = 1 / (1 + exp(-z)), and the gradient is
// chained through sigmoid's backward, grad * (1 - y) * y.  This scalar form (accurate expf + IEEE divide,
// the operations of torch's CUDA sigmoid kernel) serves the C % 4 != 0 path; the vector path uses
// rn_sigmoid_pair below.
__device__ __forceinline__ float rn_sigmoid(float z) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-z))); }

// One class element.  POS selects the target (t = 1 for the matched class of a positive anchor).
// Returns the gradient w.r.t. the probability (already scaled by `ga` = alpha-weight * upstream) and
// adds the focal term divided by (alpha-weight/2) to `acc`.
//   t = 0:  l = -(1-a) r^g log(q),  dl/dp = (1-a) ( r^g / q - g r^(g-1) log q ),  r = 1-(1-p), q = 1-p
//   t = 1:  l = -a q^g log(p),      dl/dp = -a ( q^g / p - g q^(g-1) log p )
// (r, not p, on purpose: the reference computes (1-pt) with pt = 1-p in fp32, Vision.py:1525-1527.)
template <bool POS, bool G2, bool GRAD>
__device__ __forceinline__ float rn_focal_elem(float x, float lo, float hi, float gamma, float ga, float &acc) {
    const float p = fminf(fmaxf(x, lo), hi);  // Vision.py:1524
    const float q = 1.0f - p;
    const float u = POS ? q : (1.0f - q);
    const float v = POS ? p : q;
    const float l2 = rn_neg2log(v);  // -2 log v  >= 0
    float pw, pw1;
    if (G2) {
        pw1 = u;
        pw = u * u;  // pow(x, 2.0) == x*x in torch
    } else {
        pw1 = powf(u, gamma - 1.0f);
        pw = pw1 * u;
    }
    acc = fmaf(pw, l2, acc);
    if (!GRAD) return 0.0f;
    // pw / v - gamma * pw1 * log v  =  pw * rcp(v) + (gamma/2) * pw1 * l2
    float g = G2 ? fmaf(pw1, l2, pw * rn_rcp_approx(v)) : fmaf(0.5f * gamma * pw1, l2, pw * rn_rcp_approx(v));
    g *= POS ? -ga : ga;
    return (p == x) ? g : 0.0f;  // clamp backward: pass-through iff lo <= x <= hi (inclusive)
}

// ------------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): one instruction issues two fp32
// operations.  The loss kernel is limited by instruction issue, not by the fp32 pipes, so packing the
// arithmetic of two neighbouring class elements halves the issue slots the polynomial and the
// focal-term algebra need (profiles/r01_summary.md).  Each half is an ordinary IEEE fp32 operation.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long rn_f2;
__device__ __forceinline__ rn_f2 rn_pack(float a, float b) {
    rn_f2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void rn_unpack(rn_f2 v, float &a, float &b) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ rn_f2 rn_fma2(rn_f2 a, rn_f2 b, rn_f2 c) {
    rn_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ rn_f2 rn_mul2(rn_f2 a, rn_f2 b) {
    rn_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ rn_f2 rn_add2(rn_f2 a, rn_f2 b) {
    rn_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ rn_f2 rn_splat(float a) { return rn_pack(a, a); }

// sigmoid of two logits, arithmetic packed two-wide: 2^(z * -log2 e) with the product's rounding error
// carried in a correction term, MUFU.EX2, 1 + e, MUFU.RCP refined by one Newton step.  Within ~5 ulp of the
// correctly rounded value (MUFU.EX2 itself is a 2-ulp approximation); ~5 issue slots per element instead of
// the ~17 of expf + an IEEE divide.
__device__ __forceinline__ void rn_sigmoid_pair(float z0, float z1, float &y0, float &y1) {
    const rn_f2 z = rn_pack(z0, z1);
    const rn_f2 c_hi = rn_splat(-1.4426950216293335f);       // float32(-log2 e)
    const rn_f2 t_hi = rn_mul2(z, c_hi);
    rn_f2 t_lo = rn_fma2(z, c_hi, t_hi ^ 0x8000000080000000ull);          // exact residual of the product
    t_lo = rn_fma2(z, rn_splat(-1.925963033500011e-08f), t_lo);           // + z * (-log2 e - float32(-log2 e))
    float a0, a1;
    rn_unpack(t_hi, a0, a1);
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    const rn_f2 eb = rn_pack(e0, e1);
    const rn_f2 e = rn_fma2(rn_mul2(eb, rn_splat(0.6931471805599453f)), t_lo, eb);  // 2^(t_hi + t_lo)
    const rn_f2 one = rn_splat(1.0f);
    const rn_f2 d = rn_add2(e, one);
    float d0, d1;
    rn_unpack(d, d0, d1);
    const rn_f2 r0 = rn_pack(rn_rcp_approx(d0), rn_rcp_approx(d1));
    const rn_f2 err = rn_fma2(d ^ 0x8000000080000000ull, r0, one);       // 1 - d * r0
    rn_unpack(rn_fma2(r0, err, r0), y0, y1);
}

// Two background (target 0) class elements with gamma == 2: same mathematics as
// rn_focal_elem<false, true, GRAD>, arithmetic packed two-wide.  acc2 accumulates pw * (-2 log q).
template <bool GRAD>
__device__ __forceinline__ void rn_focal_pair_neg(float x0, float x1, float lo, float hi, float ga, rn_f2 &acc2,
                                                  float &g0, float &g1) {
    const float p0 = fminf(fmaxf(x0, lo), hi), p1 = fminf(fmaxf(x1, lo), hi);  // Vision.py:1524
    const rn_f2 one = rn_splat(1.0f), mone = rn_splat(-1.0f);
    const rn_f2 q = rn_fma2(rn_pack(p0, p1), mone, one);  // 1 - p  (exact product, one rounding)
    const rn_f2 u = rn_fma2(q, mone, one);                // 1 - (1 - p), Vision.py:1525-1527
    float q0, q1;
    rn_unpack(q, q0, q1);
    // -2 log(q): range reduction per element (integer), polynomial packed
    const int i0 = __float_as_int(q0), i1 = __float_as_int(q1);
    const int t0 = (i0 - 0x3f3504f3) & 0xff800000, t1 = (i1 - 0x3f3504f3) & 0xff800000;
    const rn_f2 f = rn_add2(rn_pack(__int_as_float(i0 - t0), __int_as_float(i1 - t1)), mone);
    const rn_f2 e23 = rn_pack((float)t0, (float)t1);
    rn_f2 p = rn_splat(-2.0f * 8.700362962e-02f);
    p = rn_fma2(p, f, rn_splat(-2.0f * -1.426749380e-01f));
    p = rn_fma2(p, f, rn_splat(-2.0f * 1.491478973e-01f));
    p = rn_fma2(p, f, rn_splat(-2.0f * -1.657758280e-01f));
    p = rn_fma2(p, f, rn_splat(-2.0f * 1.996306205e-01f));
    p = rn_fma2(p, f, rn_splat(-2.0f * -2.500133718e-01f));
    p = rn_fma2(p, f, rn_splat(-2.0f * 3.333391077e-01f));
    const rn_f2 z = rn_mul2(f, f);
    const rn_f2 w0 = rn_fma2(rn_splat(-2.0f), f, z);
    const rn_f2 r = rn_fma2(rn_mul2(z, f), p, w0);
    const rn_f2 l2 = rn_fma2(e23, rn_splat(-2.0f * 0.69314718056f / 8388608.0f), r);
    const rn_f2 pw = rn_mul2(u, u);  // pow(x, 2.0) == x*x in torch
    acc2 = rn_fma2(pw, l2, acc2);
    if (!GRAD) return;
    // dl/dp / (alpha weight) = pw / q - 2 u log q = pw * rcp(q) + u * l2
    rn_f2 g = rn_fma2(u, l2, rn_mul2(pw, rn_pack(rn_rcp_approx(q0), rn_rcp_approx(q1))));
    g = rn_mul2(g, rn_splat(ga));
    rn_unpack(g, g0, g1);
    g0 = (p0 == x0) ? g0 : 0.0f;  // clamp backward: pass-through iff lo <= x <= hi (inclusive)
    g1 = (p1 == x1) ? g1 : 0.0f;
}

template <int V>
struct RnVec;
template <>
struct RnVec<4> {
    float4 d;
    __device__ __forceinline__ void load(const float *p) { d = rn_ldg_stream(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ void store(float *p) const { rn_stg_stream(reinterpret_cast<float4 *>(p), d); }
    __device__ __forceinline__ float &at(int e) { return e == 0 ? d.x : (e == 1 ? d.y : (e == 2 ? d.z : d.w)); }
};
template <>
struct RnVec<1> {
    float d;
    __device__ __forceinline__ void load(const float *p) { d = __ldg(p); }
    __device__ __forceinline__ void store(float *p) const { *p = d; }
    __device__ __forceinline__ float &at(int) { return d; }
};

// One sub-tile of RN_LOSS_TILE vectors: U independent 128-bit loads per thread are issued first, then
// the element math, then the stores.  FULL = the sub-tile lies completely inside the image, so there is
// no per-vector bounds predicate; addresses are one 64-bit base per thread plus immediates.
template <int V, int CVT, bool G2, bool GRAD, bool FULL, bool LOGITS>
__device__ __forceinline__ void rn_loss_subtile(const RnLossParams &P, const float *__restrict__ x_img,
                                                float *__restrict__ dx_img, const int32_t *__restrict__ m_img,
                                                const int *s_cat, int CV, int nvec, int tile0, float gl,
                                                float &acc_neg, float &acc_pos) {
    const int tid = threadIdx.x;
    const int v0 = tile0 + tid;
    const float *xp = x_img + (size_t)v0 * V;
    RnVec<V> xv[RN_LOSS_U];
    int mrow[RN_LOSS_U];
#pragma unroll
    for (int u = 0; u < RN_LOSS_U; ++u) {
        const int v = v0 + u * RN_THREADS;
        if (FULL || v < nvec) {
            xv[u].load(xp + (size_t)u * RN_THREADS * V);
            mrow[u] = __ldg(m_img + v / CV);
        } else {
            mrow[u] = RN_MATCH_IGNORE;
#pragma unroll
            for (int e = 0; e < V; ++e) xv[u].at(e) = 0.5f;
        }
    }
    float *dp = GRAD ? dx_img + (size_t)v0 * V : nullptr;
#pragma unroll
    for (int u = 0; u < RN_LOSS_U; ++u) {
        const int m = mrow[u];
        const float a_row = (m == RN_MATCH_IGNORE) ? 0.0f : P.a_neg;  // ignored anchors contribute nothing
        const float ga = a_row * gl;
        float part = 0.0f;
        RnVec<V> gv;
        float y[V];  // probabilities: the input itself, or sigmoid(logit)
        if (LOGITS && V == 4) {
            rn_sigmoid_pair(xv[u].at(0), xv[u].at(1), y[0], y[1]);
            rn_sigmoid_pair(xv[u].at(2), xv[u].at(3), y[2 % V], y[3 % V]);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) y[e] = LOGITS ? rn_sigmoid(xv[u].at(e)) : xv[u].at(e);
        }
        bool slow = false;
        int pe = -1;
        if (m >= 0) {  // rare: a positive anchor; is its class inside this vector?  (Vision.py:1588-1593)
            const int v = v0 + u * RN_THREADS;
            pe = s_cat[m] - (v - (v / CV) * CV) * V;
            slow = (unsigned)pe < (unsigned)V;
        }
        if (!slow) {  // common case: every element has target 0
            if (V == 4 && G2) {
                rn_f2 acc2 = 0ull;  // (+0.0f, +0.0f)
                rn_focal_pair_neg<GRAD>(y[0], y[1], P.lo, P.hi, ga, acc2, gv.at(0), gv.at(1));
                rn_focal_pair_neg<GRAD>(y[2], y[3], P.lo, P.hi, ga, acc2, gv.at(2), gv.at(3));
                float s0, s1;
                rn_unpack(acc2, s0, s1);
                part = s0 + s1;
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) gv.at(e) = rn_focal_elem<false, G2, GRAD>(y[e], P.lo, P.hi, P.gamma, ga, part);
            }
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (e == pe) {
                    float pp = 0.0f;
                    gv.at(e) = rn_focal_elem<true, G2, GRAD>(y[e], P.lo, P.hi, P.gamma, P.a_pos * gl, pp);
                    acc_pos = fmaf(0.5f * P.a_pos, pp, acc_pos);
                } else {
                    gv.at(e) = rn_focal_elem<false, G2, GRAD>(y[e], P.lo, P.hi, P.gamma, ga, part);
                }
            }
        }
        if (LOGITS && P.probs != nullptr && (FULL || v0 + u * RN_THREADS < nvec)) {
            float *pp = P.probs + ((size_t)blockIdx.y * nvec + v0 + (size_t)u * RN_THREADS) * V;
#pragma unroll
            for (int e = 0; e < V; ++e) pp[e] = y[e];
        }
        if (LOGITS && GRAD) {  // sigmoid backward: grad * (1 - y) * y
#pragma unroll
            for (int e = 0; e < V; ++e) gv.at(e) = (gv.at(e) * (1.0f - y[e])) * y[e];
        }
        acc_neg = fmaf(0.5f * a_row, part, acc_neg);
        if (GRAD && (FULL || v0 + u * RN_THREADS < nvec)) gv.store(dp + (size_t)u * RN_THREADS * V);
    }
}

// V: floats per vector (4 when C % 4 == 0, else 1).  CVT: compile-time vectors per row (0 = runtime).
// Each CTA handles P.iters consecutive sub-tiles of one image (the prologue -- ground-truth compaction,
// per-image scalars -- and the block reduction are paid once per CTA).
// __launch_bounds__(256, 3): three resident CTAs per SM need <= 85 registers per thread.  Without the
// bound ptxas drifted from 80 to 88 registers after an unrelated parameter-struct change, dropping
// occupancy to two CTAs per SM and the kernel from 370 us to 419 us (profiles/r01_summary.md).
template <int V, int CVT, bool G2, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(RN_THREADS, 3)
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
    {   // Launched with PDL right behind rn_assign: pull this CTA's first sub-tile of `clas` (independent of the
        // assignment) towards L2 while the assignment kernel drains, then wait for its matches / npos.
        const float *first = P.clas + ((size_t)b * nvec + (size_t)cta0 + (size_t)tid * RN_LOSS_U) * V;
        if (cta0 + tid * RN_LOSS_U < nvec) asm volatile("prefetch.global.L2 [%0];" ::"l"(first));
        rn_pdl_wait();
    }

    const int n_pos = P.npos[b];
    const float n_norm = fmaxf((float)n_pos, 1.0f);    // clamp(min=1), Vision.py:1530
    const float gl = __fdiv_rn(P.wc_over_bs, n_norm);  // upstream of every focal term
    const float *x_img = P.clas + (size_t)b * A * P.C;
    float *dx_img = GRAD ? P.dclas + (size_t)b * A * P.C : nullptr;
    const int32_t *m_img = P.matches + (size_t)b * A;
    __syncthreads();  // s_cat / s_box visible

    float acc_neg = 0.0f, acc_pos = 0.0f;
#pragma unroll 1
    for (int tile0 = cta0; tile0 < cta1; tile0 += RN_LOSS_TILE) {
        if (tile0 + RN_LOSS_TILE <= nvec)
            rn_loss_subtile<V, CVT, G2, GRAD, true, LOGITS>(P, x_img, dx_img, m_img, s_cat, CV, nvec, tile0, gl, acc_neg, acc_pos);
        else
            rn_loss_subtile<V, CVT, G2, GRAD, false, LOGITS>(P, x_img, dx_img, m_img, s_cat, CV, nvec, tile0, gl, acc_neg, acc_pos);
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

// One CTA, one warp per image: sums the image's CTA partials in a fixed order (float64), normalises like
// the reference (Vision.py:1530, :1566); thread 0 then accumulates over images in fp32 in image order
// (Vision.py:1640-1641) and combines (Vision.py:1643-1644).  (Folding this into the loss kernel
// with a last-CTA election -- __threadfence + ticket atomic per CTA -- was measured slower twice, also when
// only warp 0 stays for the election: COCO step 0.378 -> 0.393 ms, Pascal 78 -> 87 us; profiles/r01_summary.md.)
__global__ void __launch_bounds__(1024)
rn_loss_final_kernel(const float2 *__restrict__ partials, const int32_t *__restrict__ npos, int B, int tiles,
                     float w_reg, float w_clas, float bs, float *__restrict__ per_image /*[B][2]*/,
                     float *__restrict__ out3) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    rn_pdl_wait();  // launched with PDL behind the loss kernel: its partials must be complete and visible
    for (int b = warp; b < B; b += nwarps) {
        double cs = 0.0, rs = 0.0;
        const float2 *p = partials + (size_t)b * tiles;
#pragma unroll 4
        for (int t = lane; t < tiles; t += 32) {
            const float2 v = p[t];
            cs += (double)v.x;
            rs += (double)v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cs += __shfl_xor_sync(RN_FULL_MASK, cs, o);
            rs += __shfl_xor_sync(RN_FULL_MASK, rs, o);
        }
        if (lane == 0) {
            const int n = npos[b];
            const float n_norm = fmaxf((float)n, 1.0f);
            per_image[2 * b + 0] = n > 0 ? __fdiv_rn((float)rs, (float)(4 * n)) : 0.0f;  // reg loss of image b
            per_image[2 * b + 1] = __fdiv_rn((float)cs, n_norm);                          // clas loss of image b
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float reg_total = 0.f, clas_total = 0.f;
        for (int b = 0; b < B; ++b) {
            reg_total = __fadd_rn(reg_total, per_image[2 * b + 0]);
            clas_total = __fadd_rn(clas_total, per_image[2 * b + 1]);
        }
        const float reg_loss = __fdiv_rn(reg_total, bs), clas_loss = __fdiv_rn(clas_total, bs);
        out3[0] = __fadd_rn(__fmul_rn(w_reg, reg_loss), __fmul_rn(w_clas, clas_loss));
        out3[1] = reg_loss;
        out3[2] = clas_loss;
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

template <int V, int CVT, bool LOGITS>
static void rn_launch_loss(bool g2, bool grad, dim3 grid, size_t smem, cudaStream_t s, const RnLossParams &P,
                           const RnGeom &g) {
    if (g2 && grad) rn_launch_pdl(rn_loss_kernel<V, CVT, true, true, LOGITS>, grid, dim3(RN_THREADS), smem, s, P, g);
    else if (g2) rn_launch_pdl(rn_loss_kernel<V, CVT, true, false, LOGITS>, grid, dim3(RN_THREADS), smem, s, P, g);
    else if (grad) rn_launch_pdl(rn_loss_kernel<V, CVT, false, true, LOGITS>, grid, dim3(RN_THREADS), smem, s, P, g);
    else rn_launch_pdl(rn_loss_kernel<V, CVT, false, false, LOGITS>, grid, dim3(RN_THREADS), smem, s, P, g);
}

static int rn_loss_impl(bool logits, float *probs, const float *clas, const float *reg, const float *gt_boxes,
                        const int64_t *gt_cats, const int32_t *matches, const int32_t *npos, int B, int A, int C, int M,
                        int H, int W, const double *base, int K, const float *anchors, double alpha, double gamma,
                        double beta, int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                        size_t workspace_bytes, void *stream);

extern "C" int rn_loss(const float *clas, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
                       const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
                       const double *base, int K, const float *anchors, double alpha, double gamma, double beta,
                       int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                       size_t workspace_bytes, void *stream) {
    return rn_loss_impl(false, nullptr, clas, reg, gt_boxes, gt_cats, matches, npos, B, A, C, M, H, W, base, K, anchors, alpha,
                        gamma, beta, B_global, dclas, dreg, out3, workspace, workspace_bytes, stream);
}

extern "C" int rn_loss_logits(const float *logits, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
                              const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
                              const double *base, int K, const float *anchors, double alpha, double gamma,
                              double beta, int B_global, float *dlogits, float *dreg, float *probs_out, float *out3,
                              void *workspace, size_t workspace_bytes, void *stream) {
    return rn_loss_impl(true, probs_out, logits, reg, gt_boxes, gt_cats, matches, npos, B, A, C, M, H, W, base, K, anchors, alpha,
                        gamma, beta, B_global, dlogits, dreg, out3, workspace, workspace_bytes, stream);
}

static int rn_loss_impl(bool logits, float *probs, const float *clas, const float *reg, const float *gt_boxes,
                        const int64_t *gt_cats, const int32_t *matches, const int32_t *npos, int B, int A, int C, int M,
                        int H, int W, const double *base, int K, const float *anchors, double alpha, double gamma,
                        double beta, int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                        size_t workspace_bytes, void *stream) {
    if (B <= 0 || A <= 0 || C <= 0 || M < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss: B=%d A=%d C=%d M=%d", B, A, C, M);
    if (!clas || !reg || !matches || !npos || !out3 || (M > 0 && (!gt_boxes || !gt_cats)))
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
    P.matches = matches; P.npos = npos; P.table = reinterpret_cast<const float4 *>(anchors);
    P.dclas = dclas; P.dreg = dreg; P.probs = probs;
    P.B = B; P.A = A; P.C = C; P.CV = C / V; P.M = M; P.tiles = tiles; P.iters = rn_loss_iters(B, A, C);
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
    if (logits) {
        if (V == 4 && C == 80) rn_launch_loss<4, 20, true>(g2, grad, grid, smem, s, P, g);
        else if (V == 4 && C == 20) rn_launch_loss<4, 5, true>(g2, grad, grid, smem, s, P, g);
        else if (V == 4) rn_launch_loss<4, 0, true>(g2, grad, grid, smem, s, P, g);
        else rn_launch_loss<1, 0, true>(g2, grad, grid, smem, s, P, g);
    } else {
        if (V == 4 && C == 80) rn_launch_loss<4, 20, false>(g2, grad, grid, smem, s, P, g);
        else if (V == 4 && C == 20) rn_launch_loss<4, 5, false>(g2, grad, grid, smem, s, P, g);
        else if (V == 4) rn_launch_loss<4, 0, false>(g2, grad, grid, smem, s, P, g);
        else rn_launch_loss<1, 0, false>(g2, grad, grid, smem, s, P, g);
    }
    rc = rn_check_launch("rn_loss");
    if (rc) return rc;
    rn_launch_pdl(rn_loss_final_kernel, dim3(1), dim3(1024), 0, s, reinterpret_cast<const float2 *>(P.partials), npos, B,
                  tiles, w_reg, w_clas, bs, per_image, out3);
    return rn_check_launch("rn_loss_final");
}

extern "C" int rn_scale_grads(float *dclas, size_t n_clas, float *dreg, size_t n_reg, const float *grad_out,
                              void *stream) {
    if (!grad_out || (n_clas && !dclas) || (n_reg && !dreg)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_scale_grads: null pointer");
    if ((((uintptr_t)dclas) | ((uintptr_t)dreg)) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_scale_grads: 16-byte alignment required");
    rn_scale_kernel<<<148 * 8, RN_THREADS, 0, (cudaStream_t)stream>>>(dclas, n_clas, dreg, n_reg, grad_out);
    return rn_check_launch("rn_scale_grads");
}
