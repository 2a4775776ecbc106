// rn_loss_math.cuh -- element math of the focal loss (shared by the flat [B,A,C] kernel in rn_loss.cu and the
// NCHW level-tensor kernel in rn_loss_levels.cu) and the final-reduction kernel both are followed by.
#pragma once
#include "rn_common.cuh"

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

#ifndef RN_SIGMOID_NEWTON
#define RN_SIGMOID_NEWTON 0  // 1: refine MUFU.RCP by one Newton step (3 more packed instructions per pair)
#endif
// sigmoid of two logits, arithmetic packed two-wide: 2^(z * -log2 e) with the product's rounding error
// carried in a correction term, MUFU.EX2, 1 + e, MUFU.RCP.  Within 6 ulp of the correctly rounded value (MUFU.EX2 is
// a 2-ulp approximation, MUFU.RCP a 1-ulp one; tests/test_gpu_assign_loss.py checks the bound); ~5 issue slots per
// element instead of the ~17 of expf + an IEEE divide.  The logits kernels are limited by instruction issue (issue
// active 69 %, math-pipe throttle and not-selected stalls 2.1 / 2.2 per issue, profiles/r02_summary.md): leaving out the
// Newton refinement of the reciprocal (0.5 ulp) takes the flat logits step from 0.397 to 0.380 ms and the level-tensor one
// from 0.402 to 0.382 ms.
__device__ __forceinline__ void rn_sigmoid_pair(float z0, float z1, float &y0, float &y1) {
    // Signs are folded into the constants so that no negation (two LOP3 per packed value) is needed:
    //   t_hi = z * c,  c = float32(-log2 e);   s = -(z*c - t_hi) - z*c_lo = -(t_lo)   (two FFMA2 with -c, -c_lo)
    //   2^(t_hi + t_lo) = eb + (eb * -ln 2) * s
    const rn_f2 z = rn_pack(z0, z1);
    const rn_f2 t_hi = rn_mul2(z, rn_splat(-1.4426950216293335f));
    rn_f2 s = rn_fma2(z, rn_splat(1.4426950216293335f), t_hi);            // -(exact residual of the product)
    s = rn_fma2(z, rn_splat(1.925963033500011e-08f), s);                  // - z * (-log2 e - float32(-log2 e))
    float a0, a1;
    rn_unpack(t_hi, a0, a1);
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    const rn_f2 eb = rn_pack(e0, e1);
    const rn_f2 e = rn_fma2(rn_mul2(eb, rn_splat(-0.6931471805599453f)), s, eb);  // 2^(t_hi + t_lo)
    const rn_f2 one = rn_splat(1.0f);
    const rn_f2 d = rn_add2(e, one);
    float d0, d1;
    rn_unpack(d, d0, d1);
    const rn_f2 r0 = rn_pack(rn_rcp_approx(d0), rn_rcp_approx(d1));
#if RN_SIGMOID_NEWTON
    const rn_f2 dn = rn_fma2(e, rn_splat(-1.0f), rn_splat(-1.0f));        // -(1 + e), same rounding as d
    const rn_f2 err = rn_fma2(dn, r0, one);                               // 1 - d * r0
    rn_unpack(rn_fma2(r0, err, r0), y0, y1);
#else
    rn_unpack(r0, y0, y1);
#endif
}

// Two background (target 0) class elements with gamma == 2: same mathematics as
// rn_focal_elem<false, true, GRAD>, arithmetic packed two-wide.  acc2 accumulates pw * (-2 log q); ga2 holds the two
// elements' gradient scales.  CHAIN (logits variants): x0, x1 are sigmoid outputs y and the returned gradient is the
// one w.r.t. the LOGIT.  For a pass-through element (p == y) 1 - y is the q computed here, so
//   dl/dz = [pw/q + u*l2] * (1 - y) * y = u * (u + q*l2) * y        (pw = u*u)
// needs no reciprocal at all; the others are zeroed by the clamp mask either way.
template <bool GRAD, bool CHAIN = false>
__device__ __forceinline__ void rn_focal_pair_neg(float x0, float x1, float lo, float hi, rn_f2 ga2, rn_f2 &acc2,
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
    rn_f2 g;
    if (CHAIN) g = rn_mul2(rn_mul2(u, rn_fma2(q, l2, u)), rn_pack(x0, x1));
    else g = rn_fma2(u, l2, rn_mul2(pw, rn_pack(rn_rcp_approx(q0), rn_rcp_approx(q1))));
    g = rn_mul2(g, ga2);
    rn_unpack(g, g0, g1);
    g0 = (p0 == x0) ? g0 : 0.0f;  // clamp backward: pass-through iff lo <= x <= hi (inclusive)
    g1 = (p1 == x1) ? g1 : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Parameters of the flat [B,A,C] loss kernels (rn_loss.cu, rn_loss_tma.cu)
// ------------------------------------------------------------------------------------------------
struct RnLossParams {
    const float *clas;
    const float *reg;
    const float4 *gt_boxes;
    const int64_t *gt_cats;
    const int32_t *matches;   // RnMatchI32: [B][A] from rn_assign
    const uint8_t *m8;        // RnMatchU8: the byte map of rn_loss_step ([B][A]; 0 background, 255 ignored, 1 + box)
    int32_t *matches_out;     // RnMatchU8 only, may be NULL: the dense int32 assignment, written by the row owners
    const int32_t *npos;
    const float4 *table;
    float *dclas;
    float *dreg;
    float *probs;     // LOGITS only, may be NULL: sigmoid(logits) as used by the kernel (for checking / reuse)
    float *partials;  // [B][tiles][2] : {sum of focal terms, sum of smooth-L1 terms}
    int B, A, C, CV, M, tiles, iters, prefetch, resident;
    float a_pos, a_neg, gamma, lo, hi;
    float wc_over_bs, wr_over_bs;  // beta / B_global, (1-beta) / B_global   (Vision.py:1644)
};

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



// How the assignment of an anchor row reaches the loss kernels; load() returns the canonical value (>= 0 index of the
// matched ground-truth box, RN_MATCH_NEG, RN_MATCH_IGNORE).
struct RnMatchI32 {  // matches [B,A] int32 written by rn_assign
    typedef int32_t T;
    static __device__ __forceinline__ int load(const int32_t *p, int i) { return __ldg(p + i); }
    static __device__ __forceinline__ const int32_t *base(const RnLossParams &P);
};
struct RnMatchU8 {  // byte codes of the fused step: 0 background, 255 ignored, 1 + index otherwise
    typedef uint8_t T;
    static __device__ __forceinline__ int load(const uint8_t *p, int i) {
        // a plain (coherent) load: the bytes were written by other CTAs of the SAME grid, see rn_step.cu
        unsigned c;
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(c) : "l"(p + i));
        return c == 0u ? RN_MATCH_NEG : (c == 255u ? RN_MATCH_IGNORE : (int)c - 1);
    }
    static __device__ __forceinline__ const uint8_t *base(const RnLossParams &P);
};
struct RnMatchU8NC : RnMatchU8 {  // the same byte codes written by an EARLIER kernel (the byte-map chain): read-only path
    // read as a SIGNED byte the code is the canonical value + 1 (0 -> -1 background, 0xff -> -2 ignored, 1 + m -> m for
    // m <= 126), so the decode is one subtraction
    static __device__ __forceinline__ int load(const uint8_t *p, int i) {
        return (int)__ldg(reinterpret_cast<const signed char *>(p) + i) - 1;
    }
};
__device__ __forceinline__ const int32_t *RnMatchI32::base(const RnLossParams &P) { return P.matches; }
__device__ __forceinline__ const uint8_t *RnMatchU8::base(const RnLossParams &P) { return P.m8; }

// Smooth-L1 of one positive anchor (Vision.py:1532-1566): anchor `an`, its ground-truth box `tg`, predicted offsets `pr`.
// ge = (1-beta) / (B * 4 * npos): the mean() backward.  Adds the four loss terms to acc_reg, returns d loss / d reg.
__device__ __forceinline__ float4 rn_smooth_l1_row(float4 an, float4 tg, float4 pr, float ge, float &acc_reg) {
    const float knee = (float)(1.0 / 9.0), off = (float)(0.5 / 9.0);  // Vision.py:1565
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
    return make_float4(gg[0], gg[1], gg[2], gg[3]);
}

#ifndef RN_LOSS_U
#define RN_LOSS_U 8  // vectors per thread
#endif
#define RN_LOSS_TILE (RN_THREADS * RN_LOSS_U)


// One sub-tile of RN_LOSS_TILE vectors: U independent 128-bit loads per thread are issued first, then
// the element math, then the stores.  FULL = the sub-tile lies completely inside the image, so there is
// no per-vector bounds predicate; addresses are one 64-bit base per thread plus immediates.
// MT: how the assignment of a row is stored (RnMatchI32: the int32 matches of rn_assign; RnMatchU8: the byte codes of the
// fused step, rn_step.cu).  b = image (for probs_out), nvec = vectors per image, vend = end of the range this call may touch.
template <int V, int CVT, bool G2, bool GRAD, bool FULL, bool LOGITS, typename MT>
__device__ __forceinline__ void rn_loss_subtile(const RnLossParams &P, int b, const float *__restrict__ x_img,
                                                float *__restrict__ dx_img, const typename MT::T *__restrict__ m_img,
                                                const int *s_cat, int CV, int nvec, int vend, int tile0, float gl,
                                                float &acc_neg, float &acc_pos) {
    const int tid = threadIdx.x;
    const int v0 = tile0 + tid;
    const float *xp = x_img + (size_t)v0 * V;
    RnVec<V> xv[RN_LOSS_U];
    int mrow[RN_LOSS_U];
#pragma unroll
    for (int u = 0; u < RN_LOSS_U; ++u) {
        const int v = v0 + u * RN_THREADS;
        if (FULL) {
            xv[u].load(xp + (size_t)u * RN_THREADS * V);
            mrow[u] = MT::load(m_img, v / CV);
        } else {
            // Ragged end of a range: the loads stay UNCONDITIONAL (a vector past the end re-reads the last valid one), so all U
            // requests are in flight together as in the full sub-tile; the vector is then treated as an ignored row (its terms
            // are exactly zero) and its stores are predicated off.  A branch per load would serialise them.
            const bool ok = v < vend;
            const int vc = ok ? v : vend - 1;
            xv[u].load(x_img + (size_t)vc * V);
            const int mm = MT::load(m_img, vc / CV);
            mrow[u] = ok ? mm : RN_MATCH_IGNORE;
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
                const rn_f2 ga2 = rn_splat(ga);
                rn_focal_pair_neg<GRAD, LOGITS>(y[0], y[1], P.lo, P.hi, ga2, acc2, gv.at(0), gv.at(1));
                rn_focal_pair_neg<GRAD, LOGITS>(y[2], y[3], P.lo, P.hi, ga2, acc2, gv.at(2), gv.at(3));
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
        if (LOGITS && P.probs != nullptr && (FULL || v0 + u * RN_THREADS < vend)) {
            float *pp = P.probs + ((size_t)b * nvec + v0 + (size_t)u * RN_THREADS) * V;
#pragma unroll
            for (int e = 0; e < V; ++e) pp[e] = y[e];
        }
        if (LOGITS && GRAD && !(V == 4 && G2 && !slow)) {  // sigmoid backward: grad * (1 - y) * y (the packed path chains itself)
#pragma unroll
            for (int e = 0; e < V; ++e) gv.at(e) = (gv.at(e) * (1.0f - y[e])) * y[e];
        }
        acc_neg = fmaf(0.5f * a_row, part, acc_neg);
        if (GRAD && (FULL || v0 + u * RN_THREADS < vend)) gv.store(dp + (size_t)u * RN_THREADS * V);
    }
}


// One CTA, one warp per image: sums the image's CTA partials in a fixed order (float64), normalises like
// the reference (Vision.py:1530, :1566); thread 0 then accumulates over images in fp32 in image order
// (Vision.py:1640-1641) and combines (Vision.py:1643-1644).  (Folding this into the loss kernel
// with a last-CTA election -- __threadfence + ticket atomic per CTA -- was measured slower twice, also when
// only warp 0 stays for the election: COCO step 0.378 -> 0.393 ms, Pascal 78 -> 87 us; profiles/r01_summary.md.)
// The byte-map assignment of rn_loss_step as rn_loss_impl receives it (rn_loss.cu, rn_step.cu).
struct RnLossBytes {
    uint8_t *m8;
    const int32_t *clean_list;
    int32_t *clean_cnt;
    int32_t *npos_acc;
    int32_t *npos_out;
};
int rn_loss_impl(bool logits, float *probs, const float *clas, const float *reg, const float *gt_boxes,
                 const int64_t *gt_cats, const int32_t *matches, const int32_t *npos, const RnLossBytes *bytes, int B, int A,
                 int C, int M, int H, int W, const double *base, int K, const float *anchors, double alpha, double gamma,
                 double beta, int B_global, float *dclas, float *dreg, float *out3, void *workspace,
                 size_t workspace_bytes, void *stream);

struct RnFinalClean {  // rn_loss_step: leave the persistent byte map / counters as they were found (all of it may be NULL)
    uint8_t *m8;               // [B][A]
    const int32_t *clean_list; // [B][A] anchors whose byte was written
    int32_t *clean_cnt;        // [B]
    int32_t *npos_acc;         // [B] the counters `npos` points to
    int32_t *npos_out;         // [B] or NULL: where the caller wants the positive counts
    int A;
};
static __global__ void __launch_bounds__(1024)
rn_loss_final_kernel(const float2 *__restrict__ partials, const int32_t *__restrict__ npos, int B, int tiles,
                     float w_reg, float w_clas, float bs, float *__restrict__ per_image /*[B][2]*/,
                     float *__restrict__ out3, const RnFinalClean clean) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    __shared__ float2 s_pi[1024];  // per-image {reg loss, clas loss} (global `per_image` only for larger batches)
    float2 *pi = B <= 1024 ? s_pi : reinterpret_cast<float2 *>(per_image);
    rn_pdl_wait();  // launched with PDL behind the loss kernel: its partials must be complete and visible
    for (int b = warp; b < B; b += nwarps) {
        const int n = npos[b];  // issued ahead of the partial loads, consumed after them
        double cs = 0.0, rs = 0.0;
        const float2 *p = partials + (size_t)b * tiles;
#pragma unroll 1
        for (int t0 = lane; t0 < tiles; t0 += 32 * 16) {  // 16 independent loads per lane in flight, fixed summation order
            float2 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (t0 + 32 * u < tiles) ? __ldg(p + t0 + 32 * u) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                cs += (double)v[u].x;
                rs += (double)v[u].y;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cs += __shfl_xor_sync(RN_FULL_MASK, cs, o);
            rs += __shfl_xor_sync(RN_FULL_MASK, rs, o);
        }
        if (lane == 0) {
            const float n_norm = fmaxf((float)n, 1.0f);
            pi[b] = make_float2(n > 0 ? __fdiv_rn((float)rs, (float)(4 * n)) : 0.0f,  // reg loss of image b
                                __fdiv_rn((float)cs, n_norm));                         // clas loss of image b
        }
        if (clean.m8 && lane == 0 && clean.npos_out) clean.npos_out[b] = n;
    }
    __syncthreads();
    if (clean.m8) {  // the loss kernel is complete: zero the bytes the assignment wrote (all threads, image by image) ...
        for (int b = 0; b < B; ++b) {
            const int nc = clean.clean_cnt[b];
            const int32_t *lst = clean.clean_list + (size_t)b * clean.A;
            uint8_t *mb = clean.m8 + (size_t)b * clean.A;
#pragma unroll 4
            for (int i = threadIdx.x; i < nc; i += blockDim.x) mb[__ldg(lst + i)] = 0;
        }
        __syncthreads();  // ... and then the counters
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            clean.clean_cnt[b] = 0;
            clean.npos_acc[b] = 0;
        }
    }
    if (threadIdx.x == 0) {
        float reg_total = 0.f, clas_total = 0.f;
        for (int b = 0; b < B; ++b) {
            reg_total = __fadd_rn(reg_total, pi[b].x);
            clas_total = __fadd_rn(clas_total, pi[b].y);
        }
        const float reg_loss = __fdiv_rn(reg_total, bs), clas_loss = __fdiv_rn(clas_total, bs);
        out3[0] = __fadd_rn(__fmul_rn(w_reg, reg_loss), __fmul_rn(w_clas, clas_loss));
        out3[1] = reg_loss;
        out3[2] = clas_loss;
    }
}

