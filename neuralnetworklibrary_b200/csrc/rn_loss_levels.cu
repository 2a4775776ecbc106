// rn_loss_levels.cu -- the loss forward+backward of rn_loss.cu, reading the heads' NCHW level tensors directly
// (second half).
//
// The reference's heads end with  out.permute(0,2,3,1).contiguous().view(B,-1,4|C)  (retinanet.py:215-217,
// :289-295) and the model concatenates the five levels (Vision.py:1467-1468): one read+write pass over [B,A,C]
// for the permute and one for the cat, and the same again for their backward.  This kernel consumes the conv
// outputs as they are -- level l: clas_l [B, K*C, gh_l, gw_l], reg_l [B, K*4, gh_l, gw_l], channel = k*C + c
// resp. k*4 + j -- and writes the gradients in the same layout, so those passes disappear from the model.
//
// Data movement.  Anchor row (level l, cell p = iy*gw+ix, slot k) is a = off_l + p*K + k; its class c lives at
// clas_l[((b*K + k)*C + c)*P_l + p], P_l = gh_l*gw_l.  A thread owns V consecutive cells of one (b, k) -- V = the
// largest of {4, 2, 1} dividing P_l, so every plane base stays V*4-byte aligned -- reads their V assignments once
// (a stride-K gather from `matches`, amortised over all classes) and then walks the class planes with 8
// independent V*4-byte loads in flight, exactly the element math of rn_loss.cu, V*4-byte streaming stores for the
// gradient.  A warp therefore reads/writes 32*V*4 contiguous bytes per plane.  Algorithmic bytes are unchanged:
// 8*A*(C+4) per image with gradients.  No tensor cores: there is no contraction on this path.
#include <stdlib.h>
#include <string.h>

#include "rn_loss_math.cuh"

// Tuning (measured on B200, COCO B=16, A/B runs inside one gpurun call; profiles/r01_summary.md, profiles/r02_summary.md):
// both variants wait for memory most of the time (stall_long_sb ~60 % of the samples), so the CTAs are small (128
// threads: warps drift apart over a 40-class chunk and registers are only released per CTA), every load of a block is
// fenced ahead of its first consumer, and each block's loads are followed by an L2 prefetch of the NEXT block of planes
// (round 2: 0.419 -> 0.377 ms; look-ahead of 2/3/5/10 blocks is slower again).  With the prefetch in place a block of 4
// planes is enough: probabilities 4 planes, 6 CTAs/SM 0.364 ms (8 planes/6 CTAs 0.385, 8/4 0.366, 4/8 0.370, 2/8 0.388);
// logits 4 planes, 8 CTAs/SM 0.403 ms with the gradient chained through the sigmoid inside rn_focal_pair_neg (8/6 0.406,
// 2/8 0.420, 8/4 0.429, 256-thread CTAs 0.410).  A second register buffer per thread (RN_LVL_DB=1: the next block's loads
// issued ahead of the current block's arithmetic) does not help on top of the L2 prefetch: 0.363 vs 0.357 ms
// (probabilities), 0.400-0.465 vs 0.384 ms (logits, which then spill or lose occupancy).
// The lambdas of the body are force-inlined: left to nvcc's inliner, whose decision depends on the rest of the module,
// the same source produced class loops 10 % apart (two builds that differed only in the OTHER variant's macros).
#define RN_LAMBDA_INLINE __attribute__((always_inline))
#define RN_LVL_CHUNK_Q 8  // class chunks are multiples of this
#ifndef RN_LVL_THREADS_PROB
#define RN_LVL_THREADS_PROB 128
#endif
#ifndef RN_LVL_THREADS_LOGIT
#define RN_LVL_THREADS_LOGIT 128
#endif
#define RN_LVL_THREADS_MAX 256
#ifndef RN_LVL_PF_DIST
#define RN_LVL_PF_DIST 1  // blocks of look-ahead of the L2 prefetch (2, 3, 5, 10 measured slower, see load_block)
#endif
#ifndef RN_LVL_DB
#define RN_LVL_DB 0  // 1: two register buffers per thread (measured: see the tuning note)
#endif
#ifndef RN_LVL_U_PROB
#define RN_LVL_U_PROB 4
#endif
#ifndef RN_LVL_CTAS_PROB
#define RN_LVL_CTAS_PROB 6
#endif
#ifndef RN_LVL_U_LOGIT
#define RN_LVL_U_LOGIT 4
#endif
#ifndef RN_LVL_CTAS_LOGIT
#define RN_LVL_CTAS_LOGIT 8
#endif

struct RnLvlParams {
    const float *clas[RN_NUM_LEVELS];
    const float *reg[RN_NUM_LEVELS];
    float *dclas[RN_NUM_LEVELS];
    float *dreg[RN_NUM_LEVELS];
    float *probs[RN_NUM_LEVELS];  // LOGITS only, may be NULL
    const float4 *gt_boxes;
    const int64_t *gt_cats;
    const int32_t *matches;
    const int32_t *npos;
    float *partials;  // [B][grid.x][2]
    int B, A, C, M, K;
    int P[RN_NUM_LEVELS];         // cells per level
    int V[RN_NUM_LEVELS];         // cells per thread (4, 2 or 1)
    int tile0[RN_NUM_LEVELS + 1]; // first row-tile of each level in the per-image tile list
    int cchunk, nchunks;          // classes per CTA, chunks per row-tile
    int prefetch;                 // L2 prefetch RN_LVL_PF_DIST blocks of class planes ahead
    float a_pos, a_neg, gamma, lo, hi;
    float wc_over_bs, wr_over_bs;
};

template <int V>
struct RnLv;
template <>
struct RnLv<4> {
    float4 d;
    __device__ __forceinline__ void load(const float *p) { d = rn_ldg_stream(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ void store(float *p) const { rn_stg_stream(reinterpret_cast<float4 *>(p), d); }
    __device__ __forceinline__ float &at(int e) { return e == 0 ? d.x : (e == 1 ? d.y : (e == 2 ? d.z : d.w)); }
    __device__ __forceinline__ void keep() { rn_keep_live(d); }  // scheduling fence, see rn_common.cuh
};
template <>
struct RnLv<2> {
    float2 d;
    __device__ __forceinline__ void load(const float *p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(d.x), "=f"(d.y) : "l"(p));
    }
    __device__ __forceinline__ void store(float *p) const {
        asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(d.x), "f"(d.y) : "memory");
    }
    __device__ __forceinline__ float &at(int e) { return e == 0 ? d.x : d.y; }
    __device__ __forceinline__ void keep() { asm volatile("" : "+f"(d.x), "+f"(d.y)); }
};
template <>
struct RnLv<1> {
    float d;
    __device__ __forceinline__ void load(const float *p) {
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(d) : "l"(p));
    }
    __device__ __forceinline__ void store(float *p) const { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(d) : "memory"); }
    __device__ __forceinline__ float &at(int) { return d; }
    __device__ __forceinline__ void keep() { asm volatile("" : "+f"(d)); }
};

// One class element with a run-time target (the rare blocks that contain a positive anchor's class).
template <bool G2, bool GRAD>
__device__ __forceinline__ float rn_focal_elem_rt(bool pos, float x, const RnLvlParams &P, float ga_neg, float ga_pos,
                                                  float &acc_neg, float &acc_pos) {
    float t = 0.0f, g;
    if (pos) {
        g = rn_focal_elem<true, G2, GRAD>(x, P.lo, P.hi, P.gamma, ga_pos, t);
        acc_pos = fmaf(0.5f * P.a_pos, t, acc_pos);
    } else {
        g = rn_focal_elem<false, G2, GRAD>(x, P.lo, P.hi, P.gamma, ga_neg, t);
        acc_neg += t;
    }
    return g;
}

// The work of one thread: V cells of one (image, level, slot), classes [c_begin, c_end).
template <int V, int U, bool G2, bool GRAD, bool LOGITS>
__device__ __forceinline__ void rn_lvl_body(const RnLvlParams &P, const RnGeom &g, int b, int l, int row_tile, int chunk,
                                            float4 *s_box, int *s_cat, float &acc_neg_w, float &acc_pos, float &acc_reg) {
    constexpr int NPK = (V >= 2) ? V / 2 : 1;  // packed accumulators
    const int Pl = P.P[l], K = P.K, C = P.C;
    const int rows = K * Pl;
    const int r0 = (row_tile * (int)blockDim.x + threadIdx.x) * V;
    const bool valid = r0 < rows;
    const int k = valid ? r0 / Pl : 0;
    const int p = valid ? r0 - k * Pl : 0;
    const int c_begin = chunk * P.cchunk;
    const int c_end = min(C, c_begin + P.cchunk);
    const size_t plane0 = ((size_t)b * K + k) * C;  // first class plane of (b, k)
    const float *xp = P.clas[l] + plane0 * Pl + p;

    // The first block of class planes is requested before anything else: it does not depend on the ground truth,
    // on the assignment kernel this launch overlaps with (PDL) or on the CTA barrier below.
    RnLv<V> xv[U];
    // L2 prefetch of the block AFTER the one being loaded (one request per 128-byte line of the warp's contiguous run, plus the
    // warp's last lane for a run that is not line aligned): the kernel waits on memory most of the time (long_scoreboard 6.7
    // per issue), and a block's planes are ~2 us of work away -- the next block's loads then hit L2.
    const int lane_pf = threadIdx.x & 31;
    const bool pf_lane = ((lane_pf * V * 4) % 128 == 0) || lane_pf == 31;
    auto load_block = [&](RnLv<V>(&dst)[U], int c0) RN_LAMBDA_INLINE {
        const float *xq = xp + (size_t)c0 * Pl;
#pragma unroll
        for (int u = 0; u < U; ++u) dst[u].load(xq + u * Pl);
        if (P.prefetch && pf_lane && c0 + (RN_LVL_PF_DIST + 1) * U <= c_end) {
#pragma unroll
            for (int u = 0; u < U; ++u) asm volatile("prefetch.global.L2 [%0];" ::"l"(xq + (size_t)(RN_LVL_PF_DIST * U + u) * Pl));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dst[u].keep();  // every load of the block is issued before its first consumer
    };
    int c = c_begin;
    bool have = valid && (c + U <= c_end);
    if (have) {
        load_block(xv, c);
        // the blocks between the first one and the look-ahead distance
#pragma unroll
        for (int d = 1; d < RN_LVL_PF_DIST; ++d)
            if (P.prefetch && pf_lane && c + (d + 1) * U <= c_end) {
#pragma unroll
                for (int u = 0; u < U; ++u) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + (size_t)(c + d * U + u) * Pl));
            }
    }

    if (threadIdx.x < 32) rn_compact_gt(P.gt_boxes + (size_t)b * P.M, P.gt_cats + (size_t)b * P.M, P.M, s_box, nullptr, s_cat);
    rn_pdl_wait();  // launched with PDL right behind rn_assign: wait for its matches / npos
    const int n_pos = P.npos[b];
    const float n_norm = fmaxf((float)n_pos, 1.0f);    // clamp(min=1), Vision.py:1530
    const float gl = __fdiv_rn(P.wc_over_bs, n_norm);  // upstream of every focal term
    __syncthreads();                                   // s_cat / s_box visible
    if (!valid) return;

    const int32_t *mp = P.matches + (size_t)b * P.A + g.off[l] + p * K + k;  // this thread's V assignments, stride K
    const float ga_pos = P.a_pos * gl;
    float *dp = GRAD ? P.dclas[l] + plane0 * Pl + p : nullptr;
    float *pp = (LOGITS && P.probs[l]) ? P.probs[l] + plane0 * Pl + p : nullptr;
    auto next_block = [&]() RN_LAMBDA_INLINE {
        c += U;
        have = c + U <= c_end;
        if (have) load_block(xv, c);
    };

    if constexpr (G2) {
        // gamma == 2 (the reference's default): every element is first treated as background, arithmetic packed
        // two-wide, with nothing but the packed gradient scales and sums live across the class loop (registers: three
        // CTAs per SM).  The thread's few positive elements (Vision.py:1588-1593) are corrected afterwards.
        rn_f2 ga2[NPK], acc2[NPK];
        {
            float ga[V];
#pragma unroll
            for (int e = 0; e < V; ++e) ga[e] = ((__ldg(mp + e * K) == RN_MATCH_IGNORE) ? 0.0f : P.a_neg) * gl;
#pragma unroll
            for (int h = 0; h < NPK; ++h) {
                ga2[h] = (V >= 2) ? rn_pack(ga[2 * h], ga[(2 * h + 1) % V]) : rn_splat(ga[0]);
                acc2[h] = 0ull;
            }
        }
        // two background elements: probabilities (in place), gradients, packed sum
        auto two = [&](float &y0, float &y1, float &g0, float &g1, rn_f2 ga, rn_f2 &acc) RN_LAMBDA_INLINE {
            if (LOGITS) rn_sigmoid_pair(y0, y1, y0, y1);
            g0 = g1 = 0.0f;
            rn_focal_pair_neg<GRAD, LOGITS>(y0, y1, P.lo, P.hi, ga, acc, g0, g1);  // LOGITS: chained through the sigmoid
        };
        auto plane = [&](RnLv<V> &x, size_t off, rn_f2 *acc) RN_LAMBDA_INLINE {  // one class plane, V >= 2
            RnLv<V> gv;
#pragma unroll
            for (int h = 0; h < NPK; ++h) two(x.at(2 * h), x.at((2 * h + 1) % V), gv.at(2 * h), gv.at((2 * h + 1) % V), ga2[h], acc[h]);
            if (LOGITS && pp) x.store(pp + off);
            if (GRAD) gv.store(dp + off);
        };
        auto block = [&](RnLv<V>(&xb)[U], size_t off) RN_LAMBDA_INLINE {  // U class planes held in registers
            if (V >= 2) {
#pragma unroll
                for (int u = 0; u < U; ++u) plane(xb[u], off + u * Pl, acc2);
            } else {  // V == 1: pair two class planes of the same cell
#pragma unroll
                for (int u = 0; u < U; u += 2) {
                    RnLv<V> g0, g1;
                    two(xb[u].at(0), xb[u + 1].at(0), g0.at(0), g1.at(0), ga2[0], acc2[0]);
                    if (LOGITS && pp) {
                        xb[u].store(pp + off + u * Pl);
                        xb[u + 1].store(pp + off + (u + 1) * Pl);
                    }
                    if (GRAD) {
                        g0.store(dp + off + u * Pl);
                        g1.store(dp + off + (u + 1) * Pl);
                    }
                }
            }
        };
        if constexpr (RN_LVL_DB) {
            // register double buffering: the next block's loads are issued before the current block's arithmetic
            RnLv<V> xw[U];
            auto step = [&](RnLv<V>(&cur)[U], RnLv<V>(&nxt)[U]) RN_LAMBDA_INLINE {
                const size_t off = (size_t)c * Pl;
                const bool more = c + 2 * U <= c_end;
                if (more) load_block(nxt, c + U);
                block(cur, off);
                c += U;
                have = more;
            };
#pragma unroll 1
            while (have) {
                step(xv, xw);
                if (!have) break;
                step(xw, xv);
            }
        } else {
#pragma unroll 1
            while (have) {
                block(xv, (size_t)c * Pl);
                next_block();
            }
        }
        float rem1 = 0.0f;  // V == 1 only: remainder planes of the single cell
#pragma unroll 1
        for (; c < c_end; ++c) {  // remainder planes (chunk length not a multiple of U)
            RnLv<V> x1;
            x1.load(xp + (size_t)c * Pl);
            if (V >= 2) {
                plane(x1, (size_t)c * Pl, acc2);
            } else {
                RnLv<V> g0;
                float ydup = x1.at(0), gdup, s0, s1;
                rn_f2 t = 0ull;
                two(x1.at(0), ydup, g0.at(0), gdup, ga2[0], t);
                rn_unpack(t, s0, s1);
                rem1 += s0;
                if (LOGITS && pp) x1.store(pp + (size_t)c * Pl);
                if (GRAD) g0.store(dp + (size_t)c * Pl);
            }
        }
        // weight the per-cell sums, 0.5 * (1 - alpha) * sum(pw * -2 log q) (zero for ignored anchors), and redo the
        // positive elements of this thread: the matched class of a positive anchor has target 1.
        float cell_sum[V];
#pragma unroll
        for (int h = 0; h < NPK; ++h) {
            float s0, s1;
            rn_unpack(acc2[h], s0, s1);
            if (V >= 2) {
                cell_sum[2 * h] = s0;
                cell_sum[(2 * h + 1) % V] = s1;
            } else {
                cell_sum[0] = (s0 + s1) + rem1;
            }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int m = __ldg(mp + e * K);
            const float a_row = (m == RN_MATCH_IGNORE) ? 0.0f : P.a_neg;
            if (m >= 0) {
                const int pc = s_cat[m];
                if (pc >= c_begin && pc < c_end) {
                    const size_t off = (size_t)pc * Pl + e;
                    float y = __ldg(xp + off), ydup = y, gneg, gdup, s0, s1;
                    rn_f2 t = 0ull;
                    two(y, ydup, gneg, gdup, rn_splat(a_row * gl), t);  // exactly what the class loop added for it
                    rn_unpack(t, s0, s1);
                    cell_sum[e] -= s0;
                    float tp = 0.0f;
                    float gq = rn_focal_elem<true, true, GRAD>(y, P.lo, P.hi, P.gamma, ga_pos, tp);
                    acc_pos = fmaf(0.5f * P.a_pos, tp, acc_pos);
                    if (LOGITS && GRAD) gq = (gq * (1.0f - y)) * y;
                    if (GRAD) dp[off] = gq;
                }
            }
            acc_neg_w = fmaf(0.5f * a_row, cell_sum[e], acc_neg_w);
        }
    } else {
        // general gamma: scalar element math with a run-time target
        auto elems = [&](RnLv<V> &x, int cls, size_t off) RN_LAMBDA_INLINE {
            RnLv<V> gv;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const int m = __ldg(mp + e * K);
                const float a_row = (m == RN_MATCH_IGNORE) ? 0.0f : P.a_neg;
                float y = x.at(e);
                if (LOGITS) {
                    float dummy;
                    rn_sigmoid_pair(y, y, y, dummy);
                }
                float t = 0.0f;
                float gq = rn_focal_elem_rt<false, GRAD>(m >= 0 && s_cat[m] == cls, y, P, a_row * gl, ga_pos, t, acc_pos);
                acc_neg_w = fmaf(0.5f * a_row, t, acc_neg_w);
                if (LOGITS && GRAD) gq = (gq * (1.0f - y)) * y;
                gv.at(e) = gq;
                if (LOGITS && pp) x.at(e) = y;
            }
            if (LOGITS && pp) x.store(pp + off);
            if (GRAD) gv.store(dp + off);
        };
#pragma unroll 1
        while (have) {
#pragma unroll
            for (int u = 0; u < U; ++u) elems(xv[u], c + u, (size_t)(c + u) * Pl);
            next_block();
        }
#pragma unroll 1
        for (; c < c_end; ++c) {
            RnLv<V> x1;
            x1.load(xp + (size_t)c * Pl);
            elems(x1, c, (size_t)c * Pl);
        }
    }

    // ---- regression rows of this thread: smooth L1 (Vision.py:1532-1566); done by the CTA of class chunk 0 ----
    if (chunk != 0) return;
    const size_t rplane0 = ((size_t)b * K + k) * 4;
    const float *rp = P.reg[l] + rplane0 * Pl + p;
    RnLv<V> gj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < V; ++e) gj[j].at(e) = 0.0f;
    int m[V];
    bool anypos = false;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        m[e] = __ldg(mp + e * K);
        anypos |= m[e] >= 0;
    }
    if (anypos) {
        const float numel = (float)(4 * n_pos);
        const float ge = n_pos > 0 ? __fdiv_rn(P.wr_over_bs, numel) : 0.0f;  // mean() backward
        const float knee = (float)(1.0 / 9.0), off = (float)(0.5 / 9.0);       // Vision.py:1565
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (m[e] < 0) continue;
            const float4 an = rn_anchor_from_param(g, nullptr, g.off[l] + (p + e) * K + k);
            const float4 tg = s_box[m[e]];
            const float aw = __fsub_rn(an.z, an.x), ah = __fsub_rn(an.w, an.y);
            const float acx = __fadd_rn(an.x, __fmul_rn(0.5f, aw)), acy = __fadd_rn(an.y, __fmul_rn(0.5f, ah));
            float tw = __fsub_rn(tg.z, tg.x), th = __fsub_rn(tg.w, tg.y);
            const float tcx = __fadd_rn(tg.x, __fmul_rn(0.5f, tw)), tcy = __fadd_rn(tg.y, __fmul_rn(0.5f, th));
            tw = fmaxf(tw, 1.0f);  // Vision.py:1553-1554
            th = fmaxf(th, 1.0f);
            float ts[4];
            ts[0] = __fdiv_rn(__fdiv_rn(__fsub_rn(tcx, acx), aw), 0.1f);  // Vision.py:1556, :1562
            ts[1] = __fdiv_rn(__fdiv_rn(__fsub_rn(tcy, acy), ah), 0.1f);
            ts[2] = __fdiv_rn(logf(__fdiv_rn(tw, aw)), 0.2f);             // Vision.py:1558
            ts[3] = __fdiv_rn(logf(__fdiv_rn(th, ah)), 0.2f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float pv = __ldg(rp + (size_t)j * Pl + e);
                const float d = __fsub_rn(ts[j], pv);
                const float diff = fabsf(d);
                float lj, gd;
                if (diff < knee) {
                    lj = __fmul_rn(4.5f, __fmul_rn(diff, diff));
                    gd = __fmul_rn(__fmul_rn(ge, 4.5f), __fmul_rn(2.0f, diff));
                } else {
                    lj = __fsub_rn(diff, off);
                    gd = ge;
                }
                acc_reg += lj;
                gj[j].at(e) = d > 0.0f ? -gd : (d < 0.0f ? gd : 0.0f);  // -sign(t - p) * gd
            }
        }
    }
    if (GRAD) {
        float *drp = P.dreg[l] + rplane0 * Pl + p;
#pragma unroll
        for (int j = 0; j < 4; ++j) gj[j].store(drp + (size_t)j * Pl);
    }
}

// grid = (tiles per image x class chunks, B).  blockIdx.x -> (row tile, class chunk), row tile -> level.
template <bool G2, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(LOGITS ? RN_LVL_THREADS_LOGIT : RN_LVL_THREADS_PROB, LOGITS ? RN_LVL_CTAS_LOGIT : RN_LVL_CTAS_PROB)
rn_loss_levels_kernel(const __grid_constant__ RnLvlParams P, const __grid_constant__ RnGeom g) {
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *s_box = reinterpret_cast<float4 *>(smem);
    int *s_cat = reinterpret_cast<int *>(s_box + P.M);
    __shared__ float s_red[2][RN_LVL_THREADS_MAX / 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int chunk = blockIdx.x % P.nchunks;
    const int tile = blockIdx.x / P.nchunks;
    const int l = (tile >= P.tile0[1]) + (tile >= P.tile0[2]) + (tile >= P.tile0[3]) + (tile >= P.tile0[4]);
    const int row_tile = tile - P.tile0[l];

    float acc_neg = 0.0f, acc_pos = 0.0f, acc_reg = 0.0f;
    const int V = P.V[l];
    constexpr int U = LOGITS ? RN_LVL_U_LOGIT : RN_LVL_U_PROB;
    if (V == 4) rn_lvl_body<4, U, G2, GRAD, LOGITS>(P, g, b, l, row_tile, chunk, s_box, s_cat, acc_neg, acc_pos, acc_reg);
    else if (V == 2) rn_lvl_body<2, U, G2, GRAD, LOGITS>(P, g, b, l, row_tile, chunk, s_box, s_cat, acc_neg, acc_pos, acc_reg);
    else rn_lvl_body<1, U, G2, GRAD, LOGITS>(P, g, b, l, row_tile, chunk, s_box, s_cat, acc_neg, acc_pos, acc_reg);

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
        for (int w = 0; w < (int)blockDim.x / 32; ++w) {
            cs += s_red[0][w];
            rs += s_red[1][w];
        }
        reinterpret_cast<float2 *>(P.partials)[(size_t)b * gridDim.x + blockIdx.x] = make_float2(cs, rs);
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct RnLvlPlan {
    int P[RN_NUM_LEVELS], V[RN_NUM_LEVELS], tile0[RN_NUM_LEVELS + 1];
    int A, cchunk, nchunks, grid_x, threads;
};

static void rn_lvl_plan(RnLvlPlan *pl, int B, int H, int W, int K, int C, bool logits) {
    const int threads = logits ? RN_LVL_THREADS_LOGIT : RN_LVL_THREADS_PROB;
    const int ctas = logits ? RN_LVL_CTAS_LOGIT : RN_LVL_CTAS_PROB;
    pl->threads = threads;
    pl->A = 0;
    pl->tile0[0] = 0;
    for (int l = 0; l < RN_NUM_LEVELS; ++l) {
        const int s = 8 << l;
        const int gh = (H + s - 1) / s, gw = (W + s - 1) / s;  // retinanet.py:488
        const int P = gh * gw;
        pl->P[l] = P;
        pl->V[l] = (P % 4 == 0) ? 4 : ((P % 2 == 0) ? 2 : 1);
        pl->A += K * P;
        const int per_tile = threads * pl->V[l];
        pl->tile0[l + 1] = pl->tile0[l] + (K * P + per_tile - 1) / per_tile;
    }
    // classes per CTA: as many as possible (the stride-K gather of the assignments and the prologue are paid once
    // per CTA) while the grid keeps >= ~8 waves of 148 SMs x the resident CTAs, but never fewer than 40 classes per chunk
    // (Pascal, C = 20, B = 32: 1 chunk 65 us, 2 chunks 70 us, 3 chunks 74 us; COCO, C = 80, B = 16: 1 chunk 0.371 ms,
    // 2 chunks 0.357, 3 chunks 0.359, 5 chunks 0.371); chunks are balanced multiples of RN_LVL_CHUNK_Q.
    const int blocks = (C + RN_LVL_CHUNK_Q - 1) / RN_LVL_CHUNK_Q;
    const int nch_max = max(1, blocks / 5);
    int nch = 1;
    while (nch < nch_max && (long long)B * pl->tile0[RN_NUM_LEVELS] * nch < 8LL * 148 * ctas) ++nch;
    if (rn_opt(RN_OPT_LVL_NCHUNKS) > 0) nch = max(1, min(blocks, rn_opt(RN_OPT_LVL_NCHUNKS)));  // tuning override (rn_set_option)
    int cchunk = ((blocks + nch - 1) / nch) * RN_LVL_CHUNK_Q;
    if (cchunk > C) cchunk = C;
    pl->cchunk = cchunk;
    pl->nchunks = (C + cchunk - 1) / cchunk;
    pl->grid_x = pl->tile0[RN_NUM_LEVELS] * pl->nchunks;
}

extern "C" size_t rn_loss_levels_workspace_bytes(int B, int H, int W, int K, int C) {
    if (B <= 0 || H <= 0 || W <= 0 || K <= 0 || C <= 0) return 256;
    RnLvlPlan a, b;
    rn_lvl_plan(&a, B, H, W, K, C, false);
    rn_lvl_plan(&b, B, H, W, K, C, true);
    const size_t tiles = (size_t)(a.tile0[RN_NUM_LEVELS] > b.tile0[RN_NUM_LEVELS] ? a.tile0[RN_NUM_LEVELS] : b.tile0[RN_NUM_LEVELS]);
    // worst case over the chunking heuristic: one chunk per RN_LVL_CHUNK_Q classes
    const size_t max_chunks = (size_t)((C + RN_LVL_CHUNK_Q - 1) / RN_LVL_CHUNK_Q);
    const size_t partials = sizeof(float2) * (size_t)B * tiles * max_chunks;
    const size_t per_image = sizeof(float) * 2 * (size_t)B;
    return ((partials + 255) / 256) * 256 + ((per_image + 255) / 256) * 256;
}

template <bool LOGITS>
static void rn_launch_levels(bool g2, bool grad, dim3 grid, int threads, size_t smem, cudaStream_t s, const RnLvlParams &P, const RnGeom &g) {
    if (g2 && grad) rn_launch_pdl(rn_loss_levels_kernel<true, true, LOGITS>, grid, dim3(threads), smem, s, P, g);
    else if (g2) rn_launch_pdl(rn_loss_levels_kernel<true, false, LOGITS>, grid, dim3(threads), smem, s, P, g);
    else if (grad) rn_launch_pdl(rn_loss_levels_kernel<false, true, LOGITS>, grid, dim3(threads), smem, s, P, g);
    else rn_launch_pdl(rn_loss_levels_kernel<false, false, LOGITS>, grid, dim3(threads), smem, s, P, g);
}

extern "C" int rn_loss_levels(const float *const *clas_levels, const float *const *reg_levels, int from_logits,
                              const float *gt_boxes, const int64_t *gt_cats, const int32_t *matches, const int32_t *npos,
                              int B, int C, int M, int H, int W, const double *base, int K, double alpha, double gamma,
                              double beta, int B_global, float *const *dclas_levels, float *const *dreg_levels,
                              float *const *probs_levels, float *out3, void *workspace, size_t workspace_bytes,
                              void *stream) {
    if (B <= 0 || C <= 0 || M < 0 || H <= 0 || W <= 0 || K <= 0 || K > RN_MAX_K)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: B=%d C=%d M=%d H=%d W=%d K=%d", B, C, M, H, W, K);
    if (!clas_levels || !reg_levels || !matches || !npos || !out3 || !base || (M > 0 && (!gt_boxes || !gt_cats)))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: null pointer");
    if ((dclas_levels == nullptr) != (dreg_levels == nullptr))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: dclas_levels and dreg_levels must both be given or both be NULL");
    if (probs_levels && !from_logits) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: probs_levels needs from_logits");
    if (B_global < B) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: B_global=%d < B=%d", B_global, B);
    if (((uintptr_t)gt_boxes) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: gt_boxes must be 16-byte aligned");
    if (workspace_bytes < rn_loss_levels_workspace_bytes(B, H, W, K, C) || !workspace || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_loss_levels: workspace needs %zu bytes, 256-byte aligned",
                            rn_loss_levels_workspace_bytes(B, H, W, K, C));
    RnLvlPlan pl;
    rn_lvl_plan(&pl, B, H, W, K, C, from_logits != 0);
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, nullptr, pl.A);
    if (rc) return rc;

    RnLvlParams P;
    const bool grad = dclas_levels != nullptr;
    for (int l = 0; l < RN_NUM_LEVELS; ++l) {
        P.clas[l] = clas_levels[l];
        P.reg[l] = reg_levels[l];
        P.dclas[l] = grad ? dclas_levels[l] : nullptr;
        P.dreg[l] = grad ? dreg_levels[l] : nullptr;
        P.probs[l] = probs_levels ? probs_levels[l] : nullptr;
        if (!P.clas[l] || !P.reg[l] || (grad && (!P.dclas[l] || !P.dreg[l])))
            return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: null level pointer (level %d)", l);
        if ((((uintptr_t)P.clas[l]) | ((uintptr_t)P.reg[l]) | ((uintptr_t)P.dclas[l]) | ((uintptr_t)P.dreg[l]) |
             ((uintptr_t)P.probs[l])) & 15)
            return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: level tensors must be 16-byte aligned (level %d)", l);
        P.P[l] = pl.P[l];
        P.V[l] = pl.V[l];
        P.tile0[l] = pl.tile0[l];
    }
    P.tile0[RN_NUM_LEVELS] = pl.tile0[RN_NUM_LEVELS];
    P.gt_boxes = reinterpret_cast<const float4 *>(gt_boxes); P.gt_cats = gt_cats;
    P.matches = matches; P.npos = npos;
    P.B = B; P.A = pl.A; P.C = C; P.M = M; P.K = K;
    P.cchunk = pl.cchunk; P.nchunks = pl.nchunks;
    P.prefetch = rn_opt(RN_OPT_LOSS_PREFETCH) != 1;  // option value 1 = no look-ahead (A/B measurements)
    P.a_pos = (float)alpha; P.a_neg = (float)(1.0 - alpha);  // Vision.py:1526
    P.gamma = (float)gamma;
    P.lo = (float)1e-4; P.hi = (float)(1.0 - 1e-4);          // Vision.py:1524
    const float bs = (float)B_global;
    const float w_reg = (float)(1.0 - beta), w_clas = (float)beta;  // Vision.py:1644
    P.wc_over_bs = w_clas / bs;
    P.wr_over_bs = w_reg / bs;
    unsigned char *wsb = reinterpret_cast<unsigned char *>(workspace);
    const size_t ws_total = rn_loss_levels_workspace_bytes(B, H, W, K, C);
    P.partials = reinterpret_cast<float *>(wsb);
    float *per_image = reinterpret_cast<float *>(wsb + ws_total - ((sizeof(float) * 2 * (size_t)B + 255) / 256) * 256);

    const size_t smem = (size_t)M * (sizeof(float4) + sizeof(int));
    if (smem > 48 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_levels: M=%d too large", M);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(pl.grid_x, B);
    const bool g2 = (gamma == 2.0);
    if (from_logits) rn_launch_levels<true>(g2, grad, grid, pl.threads, smem, s, P, g);
    else rn_launch_levels<false>(g2, grad, grid, pl.threads, smem, s, P, g);
    rc = rn_check_launch("rn_loss_levels");
    if (rc) return rc;
    RnFinalClean clean;
    memset(&clean, 0, sizeof(clean));
    rn_launch_pdl(rn_loss_final_kernel, dim3(1), dim3(1024), 0, s, reinterpret_cast<const float2 *>(P.partials), npos, B,
                  pl.grid_x, w_reg, w_clas, bs, per_image, out3, clean);
    return rn_check_launch("rn_loss_levels_final");
}
