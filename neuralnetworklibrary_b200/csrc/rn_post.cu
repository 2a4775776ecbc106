// rn_post.cu -- inference post-processing: class max, threshold, decode, clip, top-k, bitmask NMS.
//
// Replaces (reference file:line): BBoxPredictor.__call__ retinanet.py:732-812 and nms
// retinanet.py:523-711 with rel_thresh/inc/dup = None (those three optional stages act on the <= top_k
// survivors and stay on the host, see neuralnetworklibrary_b200/retinanet.py).
//
// Pipeline (all on the caller's stream, no host synchronisation):
//   K3a rn_post_scan_kernel   HBM-bound: streams clas [B,A,C] once with 128-bit loads, L lanes per
//                             anchor row, warp-shuffle (max, first-argmax); rows over the threshold
//                             are decoded + clipped (reg row is read only for them) and appended as
//                             64-bit keys {sortable score | ~anchor} with one warp-aggregated atomic.
//   K3b rn_post_select_kernel one CTA per image: exact top-k of the keys (radix select on 11-bit
//                             digits until the survivors fit in shared memory, then a bitonic sort).
//                             Keys are unique, so the order (score desc, anchor asc) is total.
//   K3c rn_post_nms_kernel    one CTA per image: re-derives class/box of the <= top_k selected anchors,
//                             builds the class-aware suppression bitmask (IoU in strict fp32), one warp
//                             sweeps it visiting only kept boxes, first max_keep survivors are written.
// Algorithmic bytes per image: 4*A*(C+4) (clas + reg read); everything after K3a touches O(top_k^2).
#include <string.h>

#include "rn_common.cuh"

#define RN_SORT_N RN_MAX_TOP_K  // keys that fit the shared-memory sort
#define RN_SEL_THREADS 1024
#define RN_DIGIT_BITS 11
#define RN_BINS (1 << RN_DIGIT_BITS)

struct RnDecode {
    float mean[4], std[4];
    float img_w, img_h;
};

// Decode + clip one anchor: retinanet.py:750-753 (anchor centre form), :772-785 (shift), :790-793
// (the four one-sided clamps).  Every op is individually rounded like the reference's tensor ops.
__device__ __forceinline__ float4 rn_decode(float4 an, float4 rg, const RnDecode &d) {
    const float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
    const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
    const float dx = __fadd_rn(__fmul_rn(rg.x, d.std[0]), d.mean[0]);
    const float dy = __fadd_rn(__fmul_rn(rg.y, d.std[1]), d.mean[1]);
    const float dw = __fadd_rn(__fmul_rn(rg.z, d.std[2]), d.mean[2]);
    const float dh = __fadd_rn(__fmul_rn(rg.w, d.std[3]), d.mean[3]);
    const float pcx = __fadd_rn(cx, __fmul_rn(w, dx)), pcy = __fadd_rn(cy, __fmul_rn(h, dy));
    const float pw = __fmul_rn(w, expf(dw)), ph = __fmul_rn(h, expf(dh));
    float4 b;
    b.x = fmaxf(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), 0.0f);
    b.y = fmaxf(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), 0.0f);
    b.z = fminf(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), d.img_w);
    b.w = fminf(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), d.img_h);
    return b;
}

__device__ __forceinline__ bool rn_box_nonempty(float4 b) {  // retinanet.py:796-798
    return (__fsub_rn(b.z, b.x) > 0.0f) && (__fsub_rn(b.w, b.y) > 0.0f);
}

__device__ __forceinline__ void rn_argmax_update(float v, int c, float &best, int &bc) {
    if (v > best) {  // strict: lowest class index wins ties (torch.max(dim=1), retinanet.py:759)
        best = v;
        bc = c;
    }
}

// ------------------------------------------------------------------------------------------------
// K3a
// ------------------------------------------------------------------------------------------------
#define RN_SCAN_ROWS 1024  // anchor rows per CTA

// Candidate key: [sortable score : 32][~anchor : 24][class : 8] when the anchor index fits 24 bits and
// the class 8 bits (PACK), else [sortable score : 32][~anchor : 32] and the class is re-derived later.
// Either way keys are unique and their descending order is (score desc, anchor asc).
__device__ __forceinline__ unsigned long long rn_make_key(float score, int a, int cls, bool pack) {
    const unsigned long long hi = (unsigned long long)rn_float_sortable(score) << 32;
    if (pack) return hi | ((unsigned long long)(0xffffffu - (unsigned)a) << 8) | (unsigned long long)(cls & 0xff);
    return hi | (unsigned long long)(0xffffffffu - (unsigned)a);
}
__device__ __forceinline__ int rn_key_anchor(unsigned long long key, bool pack) {
    return pack ? (int)(0xffffffu - (unsigned)((key >> 8) & 0xffffffull)) : (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
}

// V = floats per load (4 or 1); L = lanes per anchor row (4 or 1); NV = loads per lane per row when it
// is a compile-time constant (C == V*L*NV), 0 = runtime loop; R = rows per lane group in flight.
// All R*NV loads of a step are issued before the first compare, so every warp keeps R*NV independent
// 128-bit requests in flight (the first version had a runtime-trip-count loop: one request in flight
// per warp and 27 % of DRAM peak, profiles/r01_first_pass.md).
// Resident CTAs per SM the scan kernel is compiled for.  The number matters less than its presence: with a bare
// __launch_bounds__(256) ptxas settled on 34 registers and spread the 10 independent 128-bit loads of a step over 100
// instructions (2-4 in flight); told the register budget (5 CTAs: 47 registers) it issues them back to back.
// COCO B=64 post-processing: 0.784 -> 0.650 ms (3 / 4 / 5 / 6 CTAs: 0.659 / 0.657 / 0.650 / 0.671).
#ifndef RN_SCAN_CTAS
#define RN_SCAN_CTAS 5
#endif
template <int V, int L, int NV, int R>
__global__ void __launch_bounds__(RN_THREADS, RN_SCAN_CTAS)
rn_post_scan_kernel(const float *__restrict__ clas, const float *__restrict__ reg, int C,
                    const __grid_constant__ RnGeom g, const float4 *__restrict__ table,
                    const __grid_constant__ RnDecode dec, float thresh, int pack,
                    unsigned long long *__restrict__ keys, int32_t *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: base doubles | candidate keys of this CTA [RN_SCAN_ROWS]
    double *s_base = reinterpret_cast<double *>(smem);
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem + sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4);
    __shared__ int s_count, s_pos;
    if (!table) rn_stage_base(g, s_base);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();

    const int b = blockIdx.y, A = g.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % L;  // position inside the row group
    constexpr int GROUPS = 32 / L;
    constexpr int ROWS_PER_WARP = GROUPS * R;
    constexpr int STEPS = RN_SCAN_ROWS / ((RN_THREADS / 32) * ROWS_PER_WARP);
    constexpr int ROWS_PER_CTA = RN_SCAN_ROWS;
    const int CV = C / V;
    const float *clas_b = clas + (size_t)b * A * C;

#pragma unroll 1
    for (int step = 0; step < STEPS; ++step) {
        const int a0 = blockIdx.x * ROWS_PER_CTA + (step * (RN_THREADS / 32) + warp) * ROWS_PER_WARP + lane / L;
        float best[R];
        int bc[R];
        if (NV > 0 && V == 4) {
            float4 x[R][NV > 0 ? NV : 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int a = a0 + r * GROUPS;
                const float4 *row4 = reinterpret_cast<const float4 *>(clas_b + (size_t)min(a, A - 1) * C);
#pragma unroll
                for (int j = 0; j < NV; ++j) x[r][j] = (L == 1) ? __ldg(row4 + sub + j * L) : rn_ldg_stream(row4 + sub + j * L);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NV; ++j) rn_keep_live(x[r][j]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                best[r] = -INFINITY;
                bc[r] = 0;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const int c0 = 4 * (sub + j * L);
                    rn_argmax_update(x[r][j].x, c0 + 0, best[r], bc[r]);
                    rn_argmax_update(x[r][j].y, c0 + 1, best[r], bc[r]);
                    rn_argmax_update(x[r][j].z, c0 + 2, best[r], bc[r]);
                    rn_argmax_update(x[r][j].w, c0 + 3, best[r], bc[r]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int a = a0 + r * GROUPS;
                best[r] = -INFINITY;
                bc[r] = 0;
                if (a < A) {
                    const float *row = clas_b + (size_t)a * C;
                    if (V == 4) {
                        const float4 *row4 = reinterpret_cast<const float4 *>(row);
#pragma unroll 4
                        for (int j = sub; j < CV; j += L) {
                            const float4 xx = __ldg(row4 + j);
                            rn_argmax_update(xx.x, 4 * j + 0, best[r], bc[r]);
                            rn_argmax_update(xx.y, 4 * j + 1, best[r], bc[r]);
                            rn_argmax_update(xx.z, 4 * j + 2, best[r], bc[r]);
                            rn_argmax_update(xx.w, 4 * j + 3, best[r], bc[r]);
                        }
                    } else {
#pragma unroll 4
                        for (int j = sub; j < CV; j += L) rn_argmax_update(__ldg(row + j), j, best[r], bc[r]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = a0 + r * GROUPS;
            if (L > 1) {  // combine the L lanes of a row: larger value, then lower class index
#pragma unroll
                for (int o = L / 2; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(RN_FULL_MASK, best[r], o);
                    const int oc = __shfl_xor_sync(RN_FULL_MASK, bc[r], o);
                    if (ov > best[r] || (ov == best[r] && oc < bc[r])) {
                        best[r] = ov;
                        bc[r] = oc;
                    }
                }
            }
            bool ok = (a < A) && (sub == 0) && (best[r] > thresh);  // strict, retinanet.py:760
            if (ok) {
                const float4 an = rn_anchor(g, s_base, table, a);
                const float4 rg = __ldg(reinterpret_cast<const float4 *>(reg) + (size_t)b * A + a);
                ok = rn_box_nonempty(rn_decode(an, rg, dec));
            }
            // candidates are collected per CTA in shared memory (one shared atomic per warp) ...
            const unsigned m = __ballot_sync(RN_FULL_MASK, ok);
            if (m) {
                const int leader = __ffs(m) - 1;
                int basepos = 0;
                if (lane == leader) basepos = atomicAdd(&s_count, __popc(m));
                basepos = __shfl_sync(RN_FULL_MASK, basepos, leader);
                if (ok) s_keys[basepos + __popc(m & ((1u << lane) - 1u))] = rn_make_key(best[r], a, bc[r], pack != 0);
            }
        }
    }
    rn_pdl_trigger();  // the select kernel may be scheduled as this grid's last CTAs retire
    // ... and appended to the image's list with ONE global atomic per CTA (per-warp global atomics on a
    // single per-image counter serialised in L2 and held the first version to 27 % of DRAM peak).
    __syncthreads();
    const int n = s_count;
    if (threadIdx.x == 0 && n > 0) s_pos = atomicAdd(counts + b, n);
    __syncthreads();
    if (n > 0) {
        unsigned long long *dst = keys + (size_t)b * A + s_pos;
        for (int i = threadIdx.x; i < n; i += RN_THREADS) dst[i] = s_keys[i];
    }
}

// ------------------------------------------------------------------------------------------------
// K3a on the heads' NCHW level tensors (rn_postproc_levels; SURVEY.md section 8f row 1 for inference): clas_l is
// [B, K*C, gh_l, gw_l] with channel = k*C + c, reg_l [B, K*4, gh_l, gw_l] with channel = k*4 + j -- the conv outputs
// without the sigmoid / permute / view / cat of retinanet.py:215-217, :258, :286-295 and Vision.py:1467-1468.
// ------------------------------------------------------------------------------------------------
// The four regression values of anchor a = off_l + cell*K + k of image b.
__device__ __forceinline__ float4 rn_reg_from_levels(const RnGeom &g, const float *const *reg_lv, int b, int a) {
    const int l = (a >= g.off[1]) + (a >= g.off[2]) + (a >= g.off[3]) + (a >= g.off[4]);
    const int local = a - g.off[l];
    const int cell = local / g.K, k = local - cell * g.K;
    const size_t Pl = (size_t)g.gw[l] * g.gh[l];
    const float *r = reg_lv[l] + (((size_t)b * g.K + k) * 4) * Pl + cell;
    return make_float4(__ldg(r), __ldg(r + Pl), __ldg(r + 2 * Pl), __ldg(r + 3 * Pl));
}

struct RnScanLvParams {
    const float *clas[RN_NUM_LEVELS];
    const float *reg[RN_NUM_LEVELS];
    int P[RN_NUM_LEVELS], V[RN_NUM_LEVELS], tile0[RN_NUM_LEVELS + 1];
    int C, K, pack;
    float thresh;
    unsigned long long *keys;
    int32_t *counts;
};

#define RN_SCANLV_THREADS 256
// resident CTAs per SM the level-tensor scan is compiled for (as RN_SCAN_CTAS: ptxas needs the register budget to schedule the
// loads early): unconstrained 70-80 registers 0.659-0.665 / 0.671-0.678 ms (probabilities / logits, COCO B=64), 3 CTAs 0.659 /
// 0.650, 4 CTAs 0.644 / 0.650, 5 CTAs 0.646 / 0.656
#ifndef RN_SCANLV_CTAS
#define RN_SCANLV_CTAS 4
#endif
#define RN_SCANLV_U 8  // class planes in flight per thread

// accurate sigmoid, the operations of torch's CUDA kernel (expf + IEEE divide): the scores equal those of
// torch.sigmoid followed by the flat path bit for bit
__device__ __forceinline__ float rn_sigmoid_exact(float z) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-z))); }

template <int V>
struct RnPv;
template <>
struct RnPv<4> {
    float4 d;
    __device__ __forceinline__ void load(const float *p) { d = rn_ldg_stream(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ float at(int e) const { return e == 0 ? d.x : (e == 1 ? d.y : (e == 2 ? d.z : d.w)); }
    __device__ __forceinline__ void keep() { rn_keep_live(d); }
};
template <>
struct RnPv<2> {
    float2 d;
    __device__ __forceinline__ void load(const float *p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(d.x), "=f"(d.y) : "l"(p));
    }
    __device__ __forceinline__ float at(int e) const { return e == 0 ? d.x : d.y; }
    __device__ __forceinline__ void keep() { asm volatile("" : "+f"(d.x), "+f"(d.y)); }
};
template <>
struct RnPv<1> {
    float d;
    __device__ __forceinline__ void load(const float *p) {
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(d) : "l"(p));
    }
    __device__ __forceinline__ float at(int) const { return d; }
    __device__ __forceinline__ void keep() { asm volatile("" : "+f"(d)); }
};

// One thread: V consecutive cells of one (image, level, anchor slot); running (max, first argmax) over the C class
// planes, 8 independent V*4-byte loads in flight.  LOGITS: the class tensors hold logits; the score is sigmoid(max
// logit) -- one sigmoid per cell, not per element -- and ties are resolved on the PROBABILITIES like the reference's
// clas.max(dim=1) (retinanet.py:759), see below.
template <int V, bool LOGITS>
__device__ __forceinline__ void rn_scan_lv_body(const RnScanLvParams &S, const RnGeom &g, const RnDecode &dec, int b, int l,
                                                int row_tile, unsigned long long *s_keys, int *s_count) {
    const int Pl = S.P[l], K = S.K, C = S.C;
    const int r0 = (row_tile * RN_SCANLV_THREADS + threadIdx.x) * V;
    const bool valid = r0 < K * Pl;
    const int k = valid ? r0 / Pl : 0, p = valid ? r0 - k * Pl : 0;
    const float *xp = S.clas[l] + ((size_t)b * K + k) * C * Pl + p;
    float best[V], prev[V];  // running max and (LOGITS) the running max before the last take-over
    int bc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        best[e] = -INFINITY;
        prev[e] = -INFINITY;
        bc[e] = 0;
    }
    auto update = [&](float z, int c, int e) {
        if (z > best[e]) {  // strict: lowest class index wins ties (torch.max(dim=1), retinanet.py:759)
            if (LOGITS) prev[e] = best[e];
            best[e] = z;
            bc[e] = c;
        }
    };
    if (valid) {
        int c = 0;
#pragma unroll 1
        for (; c + RN_SCANLV_U <= C; c += RN_SCANLV_U) {
            RnPv<V> x[RN_SCANLV_U];
            const float *xq = xp + (size_t)c * Pl;
#pragma unroll
            for (int u = 0; u < RN_SCANLV_U; ++u) x[u].load(xq + u * Pl);
#pragma unroll
            for (int u = 0; u < RN_SCANLV_U; ++u) x[u].keep();
#pragma unroll
            for (int u = 0; u < RN_SCANLV_U; ++u)
#pragma unroll
                for (int e = 0; e < V; ++e) update(x[u].at(e), c + u, e);
        }
#pragma unroll 1
        for (; c < C; ++c) {
            RnPv<V> x;
            x.load(xp + (size_t)c * Pl);
#pragma unroll
            for (int e = 0; e < V; ++e) update(x.at(e), c, e);
        }
    }
    // threshold, decode + clip, empty-box filter (retinanet.py:760-798); candidates are collected per CTA
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        float score = best[e];
        if (LOGITS && valid) {
            // The reference takes the max over PROBABILITIES (retinanet.py:759 after the head's sigmoid): the class is the
            // first index whose probability equals the largest one.  sigmoid is monotonic, so that is the logit argmax
            // unless an earlier, smaller logit rounds to the same probability; the largest earlier logit (prev) tells:
            // only if it ties is the row scanned again (rare: saturated or nearly equal logits).
            score = rn_sigmoid_exact(best[e]);
            if (score > S.thresh && prev[e] > -INFINITY && rn_sigmoid_exact(prev[e]) == score) {
                for (int c = 0; c < bc[e]; ++c)
                    if (rn_sigmoid_exact(__ldg(xp + (size_t)c * Pl + e)) == score) {
                        bc[e] = c;
                        break;
                    }
            }
        }
        const int a = g.off[l] + (p + e) * K + k;
        bool ok = valid && (score > S.thresh);  // strict, retinanet.py:760
        if (ok) {
            const float4 an = rn_anchor_from_param(g, nullptr, a);
            const float4 rg = rn_reg_from_levels(g, S.reg, b, a);
            ok = rn_box_nonempty(rn_decode(an, rg, dec));
        }
        const unsigned m = __ballot_sync(RN_FULL_MASK, ok);
        if (m) {
            const int leader = __ffs(m) - 1;
            int basepos = 0;
            if (lane == leader) basepos = atomicAdd(s_count, __popc(m));
            basepos = __shfl_sync(RN_FULL_MASK, basepos, leader);
            if (ok) s_keys[basepos + __popc(m & ((1u << lane) - 1u))] = rn_make_key(score, a, bc[e], S.pack != 0);
        }
    }
}

template <bool LOGITS>
__global__ void __launch_bounds__(RN_SCANLV_THREADS, RN_SCANLV_CTAS)
rn_post_scan_levels_kernel(const __grid_constant__ RnScanLvParams S, const __grid_constant__ RnGeom g,
                           const __grid_constant__ RnDecode dec) {
    __shared__ unsigned long long s_keys[RN_SCANLV_THREADS * 4];
    __shared__ int s_count, s_pos;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int b = blockIdx.y, tile = blockIdx.x;
    const int l = (tile >= S.tile0[1]) + (tile >= S.tile0[2]) + (tile >= S.tile0[3]) + (tile >= S.tile0[4]);
    const int row_tile = tile - S.tile0[l];
    const int V = S.V[l];
    if (V == 4) rn_scan_lv_body<4, LOGITS>(S, g, dec, b, l, row_tile, s_keys, &s_count);
    else if (V == 2) rn_scan_lv_body<2, LOGITS>(S, g, dec, b, l, row_tile, s_keys, &s_count);
    else rn_scan_lv_body<1, LOGITS>(S, g, dec, b, l, row_tile, s_keys, &s_count);
    rn_pdl_trigger();  // the select kernel may be scheduled as this grid's last CTAs retire
    __syncthreads();
    const int n = s_count;
    if (threadIdx.x == 0 && n > 0) s_pos = atomicAdd(S.counts + b, n);
    __syncthreads();
    if (n > 0) {
        unsigned long long *dst = S.keys + (size_t)b * g.A + s_pos;
        for (int i = threadIdx.x; i < n; i += RN_SCANLV_THREADS) dst[i] = s_keys[i];
    }
}

// Keys for caller-provided scores (rn_nms).
__global__ void rn_make_keys_kernel(const float *__restrict__ scores, int n, unsigned long long *__restrict__ keys,
                                    int32_t *__restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = rn_make_key(scores[i], i, 0, false);
    if (i == 0) *count = n;
}

// ------------------------------------------------------------------------------------------------
// K3b: exact top-k of unique 64-bit keys, sorted descending
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rn_block_excl_scan_1024(int v, int *s_warp /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(RN_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(RN_FULL_MASK, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;  // exclusive prefix of warp totals
    }
    __syncthreads();
    const int r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(RN_SEL_THREADS)
rn_post_select_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ counts, int cap,
                      const int32_t *__restrict__ seg_off /* NULL: image b owns keys[b*cap ..); else keys[seg_off[b] .. seg_off[b+1]) */,
                      int top_k, unsigned score_lo, unsigned score_hi /* range of the keys' upper 32 bits when known, else 0, 0 */,
                      unsigned long long *__restrict__ sel, int32_t *__restrict__ nsel) {
    __shared__ unsigned long long s_keys[RN_SORT_N];
    __shared__ int s_hist[RN_BINS];
    __shared__ int s_warp[32];
    __shared__ int s_digit, s_above, s_bincnt, s_n;

    const int b = blockIdx.x, tid = threadIdx.x;
    rn_pdl_wait();     // launched with PDL behind the scan kernel
    rn_pdl_trigger();  // the NMS kernel may become resident behind us; it waits for this grid to complete
    const unsigned long long *kb = seg_off ? keys + seg_off[b] : keys + (size_t)b * cap;
    const int n = seg_off ? seg_off[b + 1] - seg_off[b] : min(counts[b], cap);
    const int K = min(top_k, n);
    if (tid == 0) nsel[b] = K;
    if (K == 0) return;

    // Sort as few keys as possible: the radix passes narrow the candidates down to `limit` keys (a
    // power of two >= K), so the bitonic network is sized by top_k, not by the candidate count.
    int limit = 1024;
    while (limit < K) limit <<= 1;
    int total;  // keys staged in s_keys
    // rank search over s_hist (bins in descending order, thread t owns bins 2t and 2t+1 from the top): the bin that holds the
    // need-th largest key, the number of keys above it and its own count -> s_digit / s_above / s_bincnt
    auto find_bin = [&](int need) {
        const int d0 = RN_BINS - 1 - 2 * tid, d1 = d0 - 1;
        const int c0 = s_hist[d0], c1 = s_hist[d1];
        const int ex = rn_block_excl_scan_1024(c0 + c1, s_warp);
        if (ex < need && need <= ex + c0) {
            s_digit = d0; s_above = ex; s_bincnt = c0;
        } else if (ex + c0 < need && need <= ex + c0 + c1) {
            s_digit = d1; s_above = ex + c0; s_bincnt = c1;
        }
        __syncthreads();
    };
    // every key of the image, four independent loads per thread in flight (a pass is a chain of L2 round trips otherwise:
    // 17 k keys per image over 1024 threads = 17 dependent-latency iterations per pass)
    auto for_each_key = [&](auto &&fn) {
        for (int i0 = tid; i0 < n; i0 += 4 * RN_SEL_THREADS) {
            unsigned long long k4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) k4[u] = (i0 + u * RN_SEL_THREADS < n) ? __ldg(kb + i0 + u * RN_SEL_THREADS) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * RN_SEL_THREADS < n) fn(k4[u]);
        }
    };
    bool staged = false;
    if (n <= limit) {
        for (int i = tid; i < n; i += RN_SEL_THREADS) s_keys[i] = kb[i];
        total = n;
        staged = true;
    } else if (score_hi > score_lo) {
        // Fast path when the caller knows the range of the keys' upper halves (post-processing: the sortable bits of the
        // score threshold and of 1.0): ONE histogram over a monotone binning of that range -- 2048 bins across
        // [score_lo, score_hi], values outside clamp into the end bins -- instead of radix digits from bit 63 down, whose
        // first 11 bits (sign, exponent, two mantissa bits) put 17 k scores into ~20 bins: three passes and 17 k
        // shared-memory atomics on a handful of addresses.  Any monotone binning keeps the selection exact; if the bin of the
        // K-th key does not narrow the candidates to `limit`, the generic passes below start over.
        const unsigned range = score_hi - score_lo;
        const int sh = max(0, 32 - __clz(range) - RN_DIGIT_BITS);
        auto bin_of = [&](unsigned long long k) {
            const unsigned sc = (unsigned)(k >> 32);
            const unsigned v = sc > score_lo ? sc - score_lo : 0u;
            return (int)min(v >> sh, (unsigned)(RN_BINS - 1));
        };
        for (int i = tid; i < RN_BINS; i += RN_SEL_THREADS) s_hist[i] = 0;
        __syncthreads();
        for_each_key([&](unsigned long long k) { atomicAdd(&s_hist[bin_of(k)], 1); });
        __syncthreads();
        find_bin(K);
        const int digit = s_digit, fits = s_above + s_bincnt <= limit;
        if (tid == 0) s_n = 0;
        __syncthreads();
        if (fits) {
            for_each_key([&](unsigned long long k) {
                if (bin_of(k) >= digit) s_keys[atomicAdd(&s_n, 1)] = k;
            });
            __syncthreads();
            total = s_n;
            staged = true;
        }
    }
    if (!staged) {
        // radix select from the most significant digit down until (#certain + #in-bin) fits
        unsigned long long prefix = 0;
        int done = 0, need = K, above_total = 0;
        while (true) {
            const int db = min(RN_DIGIT_BITS, 64 - done);
            const int shift = 64 - done - db;
            for (int i = tid; i < RN_BINS; i += RN_SEL_THREADS) s_hist[i] = 0;
            __syncthreads();
            for_each_key([&](unsigned long long k) {
                if (done == 0 || (k >> (64 - done)) == prefix) atomicAdd(&s_hist[(int)((k >> shift) & ((1u << db) - 1u))], 1);
            });
            __syncthreads();
            find_bin(need);
            above_total += s_above;
            need -= s_above;
            prefix = (prefix << db) | (unsigned long long)s_digit;
            done += db;
            const int bincnt = s_bincnt;
            __syncthreads();
            if (above_total + bincnt <= limit || done == 64) break;
        }
        if (tid == 0) s_n = 0;
        __syncthreads();
        for_each_key([&](unsigned long long k) {
            const unsigned long long top = (done == 64) ? k : (k >> (64 - done));
            if (top >= prefix) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < RN_SORT_N) s_keys[pos] = k;
            }
        });
        __syncthreads();
        total = min(s_n, RN_SORT_N);
    }
    int P = 1;
    while (P < total) P <<= 1;
    for (int i = total + tid; i < P; i += RN_SEL_THREADS) s_keys[i] = 0ull;  // sorts to the end
    __syncthreads();
    if (P <= RN_SEL_THREADS) {
        // One key per thread, held in a register: the compare-exchange steps with a partner inside the warp (j < 32) are two
        // shuffles and no barrier; only the 15 of the 55 steps of a 1024-key network whose partner lives in another warp go
        // through shared memory (double-buffered: one barrier per step).  The network in shared memory alone was 46 % of the
        // kernel's stall samples (profiles/r02_summary.md).
        unsigned long long *s_alt = s_keys + RN_SEL_THREADS;   // second buffer (RN_SORT_N >= 2 * RN_SEL_THREADS)
        unsigned long long x = tid < P ? s_keys[tid] : 0ull;
        int buf = 1;  // s_keys holds the input: the first cross-warp step writes s_alt
        for (int k = 2; k <= P; k <<= 1) {
            const bool desc = (tid & k) == 0;
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long y;
                if (j >= 32) {
                    unsigned long long *dst = buf ? s_alt : s_keys;
                    dst[tid] = x;
                    __syncthreads();
                    y = dst[tid ^ j];
                    buf ^= 1;
                } else {
                    y = __shfl_xor_sync(RN_FULL_MASK, x, j);
                }
                const bool keep_max = desc == ((tid & j) == 0);
                x = keep_max ? (x > y ? x : y) : (x < y ? x : y);
            }
        }
        if (tid < K) sel[(size_t)b * top_k + tid] = x;
        return;
    }
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += RN_SEL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = s_keys[i], y = s_keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) {
                        s_keys[i] = y;
                        s_keys[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < K; i += RN_SEL_THREADS) sel[(size_t)b * top_k + i] = s_keys[i];
}

// ------------------------------------------------------------------------------------------------
// K3c: gather + class-aware greedy NMS in rank chunks of 64 (bitmask form)
// ------------------------------------------------------------------------------------------------
struct RnNmsParams {
    const unsigned long long *sel;  // [B][top_k] sorted keys
    const int32_t *nsel;            // [B]
    // FROM_BOXES = false: re-derive from the activations
    const float *clas, *reg;
    const float4 *table;
    int A, C, pack;
    // level-tensor variant (rn_postproc_levels): reg_lv[l] is [B, K*4, gh_l, gw_l]; class always packed in the key
    const float *reg_lv[RN_NUM_LEVELS];
    int levels;
    // FROM_BOXES = true
    const float4 *boxes_in;
    const int64_t *classes_in;
    const int32_t *in_off;  // ragged batch (rn_nms_batch): image b owns inputs [in_off[b], in_off[b+1]); NULL: one image
    int top_k, max_keep;
    float max_overlap;
    // outputs (each may be NULL)
    float4 *out_boxes;
    int64_t *out_classes;
    float *out_scores;
    int32_t *out_idx;
    int32_t *out_counts;
    int32_t *out_ncand;           // optional copy of the candidate counts (what nms received)
    const int32_t *cand_counts;
};

// The greedy loop of the reference (retinanet.py:590-602) keeps the best remaining box and deletes every
// remaining box of the same class with IoU > max_overlap.  Equivalent bitmask form, processed in rank
// chunks of 64: (1) all threads test the chunk against the boxes kept so far and against itself
// ((kept + 64) x 64 pairs, IoU in strict fp32), OR-ing suppression bits into shared memory; (2) one thread
// resolves the chunk with 64-bit bit operations; (3) stop as soon as max_keep boxes are kept (later
// survivors cannot reach the output, retinanet.py:702-704).  Work is O(K * (kept + 64)) pairs instead of
// the K^2/2 of a full mask, and the default max_boxes = 20 usually ends after a few chunks.
template <bool FROM_BOXES>
__global__ void __launch_bounds__(RN_SEL_THREADS)
rn_post_nms_kernel(const __grid_constant__ RnNmsParams P, const __grid_constant__ RnGeom g,
                   const __grid_constant__ RnDecode dec) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: base doubles | boxes float4[top_k] | area float[top_k] | cls int[top_k] | keep int[top_k]
    double *s_base = reinterpret_cast<double *>(smem);
    float4 *s_box = reinterpret_cast<float4 *>(smem + sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4);
    float *s_area = reinterpret_cast<float *>(s_box + P.top_k);
    int *s_cls = reinterpret_cast<int *>(s_area + P.top_k);
    int *s_keep = s_cls + P.top_k;
    __shared__ unsigned long long s_sup, s_intra[64];
    __shared__ int s_nk;

    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    rn_pdl_wait();  // launched with PDL behind the select kernel
    const int K = P.nsel[b];
    const unsigned long long *sel = P.sel + (size_t)b * P.top_k;
    const bool pack = !FROM_BOXES && P.pack;
    if (!FROM_BOXES && !P.table) rn_stage_base(g, s_base);
    __syncthreads();

    for (int t = tid; t < K; t += nthr) {
        const unsigned long long key = sel[t];
        const int a = rn_key_anchor(key, pack);
        float4 box;
        int cls;
        if (FROM_BOXES) {
            const int ga = P.in_off ? P.in_off[b] + a : a;
            box = P.boxes_in[ga];
            cls = (int)P.classes_in[ga];
        } else {
            if (pack) {
                cls = (int)(key & 0xffull);
            } else {
                const float *row = P.clas + ((size_t)b * P.A + a) * P.C;
                float best = -INFINITY;
                cls = 0;
                for (int c = 0; c < P.C; ++c) rn_argmax_update(__ldg(row + c), c, best, cls);
            }
            const float4 an = rn_anchor(g, s_base, P.table, a);
            float4 rg;
            if (P.levels) {
                rg = rn_reg_from_levels(g, P.reg_lv, b, a);
            } else {
                rg = __ldg(reinterpret_cast<const float4 *>(P.reg) + (size_t)b * P.A + a);
            }
            box = rn_decode(an, rg, dec);
        }
        s_box[t] = box;
        s_area[t] = rn_area(box);
        s_cls[t] = cls;
    }
    if (tid == 0) s_nk = 0;
    __syncthreads();

    int nk = 0;
    for (int c0 = 0; c0 < K && nk < P.max_keep; c0 += 64) {
        const int nc = min(64, K - c0);
        if (tid < 64) s_intra[tid] = 0ull;
        if (tid == 64) s_sup = 0ull;
        __syncthreads();
        const int items = (nk + nc) * 64;
        for (int item = tid; item < items; item += nthr) {
            const int e = item & 63, r = item >> 6;
            if (e >= nc) continue;
            int i;
            if (r < nk) {
                i = s_keep[r];
            } else {
                if (e <= r - nk) continue;
                i = c0 + r - nk;
            }
            const int j = c0 + e;
            // rows are score-descending: i is the better box (retinanet.py:591-594)
            if (s_cls[i] == s_cls[j] && rn_iou(s_box[i], s_area[i], s_box[j], s_area[j]) > P.max_overlap) {
                if (r < nk) atomicOr(&s_sup, 1ull << e);
                else atomicOr(&s_intra[r - nk], 1ull << e);
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long alive = ~s_sup & ((nc == 64) ? ~0ull : ((1ull << nc) - 1ull));
            int k = nk;
            while (alive && k < P.max_keep) {
                const int bit = __ffsll((long long)alive) - 1;
                s_keep[k++] = c0 + bit;
                alive &= ~s_intra[bit];
                alive &= ~(1ull << bit);
            }
            s_nk = k;
        }
        __syncthreads();
        nk = s_nk;
    }

    if (tid == 0 && P.out_counts) P.out_counts[b] = nk;
    if (tid == 0 && P.out_ncand) P.out_ncand[b] = FROM_BOXES && P.in_off ? P.in_off[b + 1] - P.in_off[b] : P.cand_counts[b];
    for (int t = tid; t < nk; t += nthr) {
        const int i = s_keep[t];
        const unsigned long long key = sel[i];
        const size_t o = (size_t)b * P.max_keep + t;
        if (P.out_boxes) P.out_boxes[o] = s_box[i];
        if (P.out_classes) P.out_classes[o] = (int64_t)s_cls[i];
        if (P.out_scores) P.out_scores[o] = rn_sortable_float((uint32_t)(key >> 32));
        if (P.out_idx) P.out_idx[o] = (int32_t)rn_key_anchor(key, pack);
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline size_t rn_up256(size_t x) { return (x + 255) / 256 * 256; }

struct RnPostWs {
    size_t counts, nsel, keys, sel, total;
};
static RnPostWs rn_post_layout(int B, long long cap, int top_k) {
    RnPostWs w;
    size_t o = 0;
    w.counts = o; o += rn_up256(sizeof(int32_t) * (size_t)B);
    w.nsel = o;   o += rn_up256(sizeof(int32_t) * (size_t)B);
    w.keys = o;   o += rn_up256(sizeof(unsigned long long) * (size_t)B * (size_t)cap);
    w.sel = o;    o += rn_up256(sizeof(unsigned long long) * (size_t)B * (size_t)top_k);
    w.total = o;
    return w;
}

extern "C" size_t rn_postproc_workspace_bytes(int B, int A, int top_k) {
    if (B <= 0 || A <= 0 || top_k <= 0) return 256;
    return rn_post_layout(B, A, top_k).total;
}
extern "C" size_t rn_nms_workspace_bytes(int n, int top_k) {
    if (n <= 0 || top_k <= 0) return 256;
    return rn_post_layout(1, n, top_k).total;
}

// host copy of rn_float_sortable
static unsigned rn_host_sortable(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// score_range: the scores behind the keys lie in (thresh, 1] (post-processing of probabilities); lets the select kernel bin
// them in one pass.  Only a hint: keys outside the range are still selected exactly.
static int rn_launch_select_nms(bool from_boxes, int B, int cap, RnNmsParams &P, const RnGeom &g, const RnDecode &dec,
                                unsigned char *ws, const RnPostWs &L, cudaStream_t s, const int32_t *seg_off = nullptr,
                                const float *score_thresh = nullptr) {
    unsigned score_lo = 0, score_hi = 0;
    if (score_thresh && *score_thresh == *score_thresh && *score_thresh < 1.0f) {
        score_lo = rn_host_sortable(*score_thresh);
        score_hi = rn_host_sortable(1.0f);
    }
    rn_launch_pdl(rn_post_select_kernel, dim3(B), dim3(RN_SEL_THREADS), 0, s,
                  reinterpret_cast<const unsigned long long *>(ws + L.keys), reinterpret_cast<const int32_t *>(ws + L.counts), cap,
                  seg_off, P.top_k, score_lo, score_hi, reinterpret_cast<unsigned long long *>(ws + L.sel),
                  reinterpret_cast<int32_t *>(ws + L.nsel));
    int rc = rn_check_launch("rn_post_select");
    if (rc) return rc;
    P.sel = reinterpret_cast<unsigned long long *>(ws + L.sel);
    P.nsel = reinterpret_cast<int32_t *>(ws + L.nsel);
    P.cand_counts = reinterpret_cast<const int32_t *>(ws + L.counts);
    const size_t smem = sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4 +
                        (size_t)P.top_k * (sizeof(float4) + sizeof(float) + 2 * sizeof(int));
    // the opt-in for > 48 KB of dynamic shared memory is made once per process (and device), for the largest top_k
    static bool s_attr_done[2][64] = {};
    int devi = 0;
    cudaGetDevice(&devi);
    if (devi >= 0 && devi < 64 && !s_attr_done[from_boxes ? 1 : 0][devi]) {
        const int max_smem = (int)(sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4 +
                                   (size_t)RN_MAX_TOP_K * (sizeof(float4) + sizeof(float) + 2 * sizeof(int)));
        cudaError_t e = from_boxes
            ? cudaFuncSetAttribute(rn_post_nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)
            : cudaFuncSetAttribute(rn_post_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_post_nms smem: %s", cudaGetErrorString(e));
        s_attr_done[from_boxes ? 1 : 0][devi] = true;
    }
    if (from_boxes) rn_launch_pdl(rn_post_nms_kernel<true>, dim3(B), dim3(RN_SEL_THREADS), smem, s, P, g, dec);
    else rn_launch_pdl(rn_post_nms_kernel<false>, dim3(B), dim3(RN_SEL_THREADS), smem, s, P, g, dec);
    return rn_check_launch("rn_post_nms");
}

extern "C" int rn_postproc(const float *clas, const float *reg, int B, int A, int C, int H, int W, const double *base,
                           int K, const float *anchors, const float *mean, const float *std, float thresh,
                           float max_overlap, int top_k, int max_keep, float *boxes, int64_t *classes, float *scores,
                           int32_t *anchor_idx, int32_t *counts, int32_t *n_candidates, void *workspace,
                           size_t workspace_bytes, void *stream) {
    if (B <= 0 || A <= 0 || C <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc: B=%d A=%d C=%d", B, A, C);
    if (!clas || !reg || !mean || !std || !boxes || !classes || !scores || !counts)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc: null pointer");
    if (top_k < 1 || top_k > RN_MAX_TOP_K) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc: top_k=%d outside [1,%d]", top_k, RN_MAX_TOP_K);
    if (max_keep < 1 || max_keep > top_k) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc: max_keep=%d outside [1,top_k=%d]", max_keep, top_k);
    if ((((uintptr_t)reg) | ((uintptr_t)anchors) | ((uintptr_t)boxes)) & 15)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc: reg/anchors/boxes must be 16-byte aligned");
    const bool vec = (C % 4 == 0) && ((((uintptr_t)clas) & 15) == 0);
    const RnPostWs L = rn_post_layout(B, A, top_k);
    if (!workspace || workspace_bytes < L.total || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_postproc: workspace needs %zu bytes, 256-byte aligned", L.total);
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, anchors, A);
    if (rc) return rc;
    RnDecode dec;
    for (int i = 0; i < 4; ++i) {
        dec.mean[i] = mean[i];
        dec.std[i] = std[i];
    }
    dec.img_w = (float)W;  // clamp(max=width), retinanet.py:792
    dec.img_h = (float)H;  // clamp(max=height), retinanet.py:793

    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    int32_t *d_counts = reinterpret_cast<int32_t *>(ws + L.counts);
    cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)B, s);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_postproc memset: %s", cudaGetErrorString(e));

    const size_t smem = sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4 + sizeof(unsigned long long) * RN_SCAN_ROWS;
    unsigned long long *d_keys = reinterpret_cast<unsigned long long *>(ws + L.keys);
    const float4 *table = reinterpret_cast<const float4 *>(anchors);
    const int pack = (A <= (1 << 24) && C <= 256) ? 1 : 0;
#define RN_SCAN_LAUNCH(V, L, NV, R)                                                                                  \
    do {                                                                                                             \
        const int rows = RN_SCAN_ROWS;                                                                               \
        rn_post_scan_kernel<V, L, NV, R><<<dim3((A + rows - 1) / rows, B), RN_THREADS, smem, s>>>(                   \
            clas, reg, C, g, table, dec, thresh, pack, d_keys, d_counts);                                            \
    } while (0)
    if (vec && C == 80) RN_SCAN_LAUNCH(4, 4, 5, 2);        // COCO: 20 vectors per row = 4 lanes x 5 loads
    else if (vec && C == 20) RN_SCAN_LAUNCH(4, 1, 5, 2);   // Pascal: 5 vectors per row, one lane per row
    else if (vec && (C / 4) % 4 == 0) RN_SCAN_LAUNCH(4, 4, 0, 1);
    else if (vec) RN_SCAN_LAUNCH(4, 1, 0, 1);
    else RN_SCAN_LAUNCH(1, 1, 0, 1);
#undef RN_SCAN_LAUNCH
    rc = rn_check_launch("rn_post_scan");
    if (rc) return rc;

    RnNmsParams P;
    memset(&P, 0, sizeof(P));
    P.clas = clas; P.reg = reg; P.table = table; P.A = A; P.C = C; P.pack = pack;
    P.top_k = top_k; P.max_keep = max_keep; P.max_overlap = max_overlap;
    P.out_boxes = reinterpret_cast<float4 *>(boxes); P.out_classes = classes; P.out_scores = scores;
    P.out_idx = anchor_idx; P.out_counts = counts; P.out_ncand = n_candidates;
    return rn_launch_select_nms(false, B, A, P, g, dec, ws, L, s, nullptr, &thresh);
}

extern "C" int rn_postproc_levels(const float *const *clas_levels, const float *const *reg_levels, int from_logits, int B,
                                  int C, int H, int W, const double *base, int K, const float *mean, const float *std,
                                  float thresh, float max_overlap, int top_k, int max_keep, float *boxes, int64_t *classes,
                                  float *scores, int32_t *anchor_idx, int32_t *counts, int32_t *n_candidates,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || K > RN_MAX_K)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: B=%d C=%d H=%d W=%d K=%d", B, C, H, W, K);
    if (!clas_levels || !reg_levels || !base || !mean || !std || !boxes || !classes || !scores || !counts)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: null pointer");
    if (top_k < 1 || top_k > RN_MAX_TOP_K) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: top_k=%d outside [1,%d]", top_k, RN_MAX_TOP_K);
    if (max_keep < 1 || max_keep > top_k) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: max_keep=%d outside [1,top_k=%d]", max_keep, top_k);
    if (((uintptr_t)boxes) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: boxes must be 16-byte aligned");
    const int A = rn_num_anchors(H, W, K);
    if (A > (1 << 24) || C > 256)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: needs A <= 2^24 and C <= 256 (A=%d C=%d)", A, C);
    const RnPostWs L = rn_post_layout(B, A, top_k);
    if (!workspace || workspace_bytes < L.total || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_postproc_levels: workspace needs %zu bytes, 256-byte aligned", L.total);
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, nullptr, A);
    if (rc) return rc;
    RnDecode dec;
    for (int i = 0; i < 4; ++i) {
        dec.mean[i] = mean[i];
        dec.std[i] = std[i];
    }
    dec.img_w = (float)W;  // clamp(max=width), retinanet.py:792
    dec.img_h = (float)H;  // clamp(max=height), retinanet.py:793

    RnScanLvParams S;
    RnNmsParams P;
    memset(&P, 0, sizeof(P));
    S.tile0[0] = 0;
    for (int l = 0; l < RN_NUM_LEVELS; ++l) {
        if (!clas_levels[l] || !reg_levels[l]) return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: null level pointer (level %d)", l);
        if ((((uintptr_t)clas_levels[l]) | ((uintptr_t)reg_levels[l])) & 15)
            return rn_set_error(RN_ERR_INVALID_ARG, "rn_postproc_levels: level tensors must be 16-byte aligned (level %d)", l);
        const int Pl = g.gw[l] * g.gh[l];
        S.clas[l] = clas_levels[l];
        S.reg[l] = reg_levels[l];
        P.reg_lv[l] = reg_levels[l];
        S.P[l] = Pl;
        S.V[l] = (Pl % 4 == 0) ? 4 : ((Pl % 2 == 0) ? 2 : 1);
        const int per_tile = RN_SCANLV_THREADS * S.V[l];
        S.tile0[l + 1] = S.tile0[l] + (K * Pl + per_tile - 1) / per_tile;
    }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    int32_t *d_counts = reinterpret_cast<int32_t *>(ws + L.counts);
    cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)B, s);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_postproc_levels memset: %s", cudaGetErrorString(e));
    S.C = C; S.K = K; S.pack = 1; S.thresh = thresh;
    S.keys = reinterpret_cast<unsigned long long *>(ws + L.keys);
    S.counts = d_counts;
    const dim3 grid(S.tile0[RN_NUM_LEVELS], B);
    if (from_logits) rn_post_scan_levels_kernel<true><<<grid, RN_SCANLV_THREADS, 0, s>>>(S, g, dec);
    else rn_post_scan_levels_kernel<false><<<grid, RN_SCANLV_THREADS, 0, s>>>(S, g, dec);
    rc = rn_check_launch("rn_post_scan_levels");
    if (rc) return rc;

    P.levels = 1; P.A = A; P.C = C; P.pack = 1;
    P.top_k = top_k; P.max_keep = max_keep; P.max_overlap = max_overlap;
    P.out_boxes = reinterpret_cast<float4 *>(boxes); P.out_classes = classes; P.out_scores = scores;
    P.out_idx = anchor_idx; P.out_counts = counts; P.out_ncand = n_candidates;
    return rn_launch_select_nms(false, B, A, P, g, dec, ws, L, s, nullptr, &thresh);
}

extern "C" int rn_nms(const float *boxes, const int64_t *classes, const float *scores, int n, float max_overlap,
                      int top_k, int max_keep, int32_t *keep_idx, int32_t *count, void *workspace,
                      size_t workspace_bytes, void *stream) {
    if (n < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: n=%d", n);
    if (!keep_idx || !count) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: null output");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {  // retinanet.py:570
        cudaError_t e0 = cudaMemsetAsync(count, 0, sizeof(int32_t), s);
        return e0 == cudaSuccess ? RN_OK : rn_set_error(RN_ERR_CUDA, "rn_nms memset: %s", cudaGetErrorString(e0));
    }
    if (!boxes || !classes || !scores) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: null input");
    if (((uintptr_t)boxes) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: boxes must be 16-byte aligned");
    if (top_k < 1 || top_k > RN_MAX_TOP_K) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: top_k=%d outside [1,%d]", top_k, RN_MAX_TOP_K);
    if (max_keep < 1 || max_keep > top_k) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms: max_keep=%d outside [1,top_k=%d]", max_keep, top_k);
    const RnPostWs L = rn_post_layout(1, n, top_k);
    if (!workspace || workspace_bytes < L.total || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_nms: workspace needs %zu bytes, 256-byte aligned", L.total);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    rn_make_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(scores, n, reinterpret_cast<unsigned long long *>(ws + L.keys),
                                                        reinterpret_cast<int32_t *>(ws + L.counts));
    int rc = rn_check_launch("rn_make_keys");
    if (rc) return rc;
    RnGeom g;
    memset(&g, 0, sizeof(g));
    RnDecode dec;
    memset(&dec, 0, sizeof(dec));
    RnNmsParams P;
    memset(&P, 0, sizeof(P));
    P.boxes_in = reinterpret_cast<const float4 *>(boxes);
    P.classes_in = classes;
    P.top_k = top_k; P.max_keep = max_keep; P.max_overlap = max_overlap;
    P.out_idx = keep_idx; P.out_counts = count;
    return rn_launch_select_nms(true, 1, n, P, g, dec, ws, L, s);
}

// ------------------------------------------------------------------------------------------------
// nms() for MANY images in one launch (the merge step of ImageLearner.TTA_bbox, Vision.py:2104-2119, calls nms once per
// image on the concatenated predictions of its five passes): the images' boxes / classes / scores are concatenated, offsets
// [L+1] gives the image boundaries.  Optionally the boxes are first mapped back to the original image -- the un-transform of
// Vision.py:2091-2097 -- per SEGMENT (= the predictions of one pass for one image): x -= col_jit, y -= row_jit, all * factor,
// then a horizontal flip about `cols`; float64 like NumPy's arithmetic on a float32 array and int64 / float64 scalars, rounded
// to float32 once (the TEN() of Vision.py:2112).
// ------------------------------------------------------------------------------------------------
__global__ void rn_nms_batch_keys_kernel(const float *__restrict__ scores, const int32_t *__restrict__ offsets, int L, int n,
                                         unsigned long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = L;  // image of element i: largest b with offsets[b] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid;
        else hi = mid;
    }
    keys[i] = rn_make_key(scores[i], i - offsets[lo], 0, false);
}

__global__ void rn_tta_untransform_kernel(const float4 *__restrict__ boxes, const int32_t *__restrict__ seg_off, int S,
                                          const double *__restrict__ seg_par /*[S][5]: col_jit, row_jit, factor, flip, cols*/,
                                          int n, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = S;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= i) lo = mid;
        else hi = mid;
    }
    const double *p = seg_par + 5 * (size_t)lo;
    const float4 b = boxes[i];
    double x1 = __dmul_rn(p[2], __dsub_rn((double)b.x, p[0]));  // (1/(rand_scale*scale)) * (x - col_jit), Vision.py:2093-2094
    double y1 = __dmul_rn(p[2], __dsub_rn((double)b.y, p[1]));
    double x2 = __dmul_rn(p[2], __dsub_rn((double)b.z, p[0]));
    double y2 = __dmul_rn(p[2], __dsub_rn((double)b.w, p[1]));
    if (p[3] != 0.0) {  // Vision.py:2095-2096
        const double nx1 = __dsub_rn(p[4], x2), nx2 = __dsub_rn(p[4], x1);
        x1 = nx1;
        x2 = nx2;
    }
    out[i] = make_float4(__double2float_rn(x1), __double2float_rn(y1), __double2float_rn(x2), __double2float_rn(y2));
}

struct RnNmsBatchWs {
    size_t post, boxes, total;
};
static RnNmsBatchWs rn_nms_batch_layout(int n_total, int L, int top_k) {
    RnNmsBatchWs w;
    // keys are stored ragged (n_total of them); the select / nms scratch is sized per image
    RnPostWs p = rn_post_layout(L, 1, top_k);
    w.post = 0;
    size_t o = p.total + rn_up256(sizeof(unsigned long long) * (size_t)(n_total > 0 ? n_total : 1));
    w.boxes = o; o += rn_up256(sizeof(float4) * (size_t)(n_total > 0 ? n_total : 1));
    w.total = o;
    return w;
}

extern "C" size_t rn_nms_batch_workspace_bytes(int n_total, int L, int top_k) {
    if (L <= 0 || top_k <= 0) return 256;
    return rn_nms_batch_layout(n_total, L, top_k).total;
}

extern "C" int rn_nms_batch(const float *boxes, const int64_t *classes, const float *scores, const int32_t *offsets, int L,
                            int n_total, const int32_t *seg_off, const double *seg_par, int S, float max_overlap, int top_k,
                            int max_keep, float *out_boxes, int64_t *out_classes, float *out_scores, int32_t *out_idx,
                            int32_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
    if (L <= 0 || n_total < 0 || S < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: L=%d n_total=%d S=%d", L, n_total, S);
    if (!offsets || !counts || !out_boxes || !out_classes || !out_scores) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: null pointer");
    if (n_total > 0 && (!boxes || !classes || !scores)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: null input");
    if (S > 0 && (!seg_off || !seg_par)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: null segment table");
    if ((((uintptr_t)boxes) | ((uintptr_t)out_boxes)) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: boxes must be 16-byte aligned");
    if (top_k < 1 || top_k > RN_MAX_TOP_K) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: top_k=%d outside [1,%d]", top_k, RN_MAX_TOP_K);
    if (max_keep < 1 || max_keep > top_k) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: max_keep=%d outside [1,top_k=%d]", max_keep, top_k);
    if (L > 65535) return rn_set_error(RN_ERR_INVALID_ARG, "rn_nms_batch: L=%d too large", L);
    const RnNmsBatchWs W = rn_nms_batch_layout(n_total, L, top_k);
    if (!workspace || workspace_bytes < W.total || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_nms_batch: workspace needs %zu bytes, 256-byte aligned", W.total);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    const RnPostWs P0 = rn_post_layout(L, 1, top_k);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws + P0.total);
    const float4 *in_boxes = reinterpret_cast<const float4 *>(boxes);
    if (n_total > 0) {
        if (S > 0) {
            float4 *tb = reinterpret_cast<float4 *>(ws + W.boxes);
            rn_tta_untransform_kernel<<<(n_total + 255) / 256, 256, 0, s>>>(in_boxes, seg_off, S, seg_par, n_total, tb);
            in_boxes = tb;
        }
        rn_nms_batch_keys_kernel<<<(n_total + 255) / 256, 256, 0, s>>>(scores, offsets, L, n_total, keys);
        int rc = rn_check_launch("rn_nms_batch keys");
        if (rc) return rc;
    }
    RnGeom g;
    memset(&g, 0, sizeof(g));
    RnDecode dec;
    memset(&dec, 0, sizeof(dec));
    RnNmsParams P;
    memset(&P, 0, sizeof(P));
    P.boxes_in = in_boxes;
    P.classes_in = classes;
    P.in_off = offsets;
    P.top_k = top_k; P.max_keep = max_keep; P.max_overlap = max_overlap;
    P.out_boxes = reinterpret_cast<float4 *>(out_boxes); P.out_classes = out_classes; P.out_scores = out_scores;
    P.out_idx = out_idx; P.out_counts = counts;
    // select / nms read their key segments through `offsets`; their scratch (sel, nsel) lives in the per-image layout
    RnPostWs Lw = P0;
    Lw.keys = P0.total;  // the ragged key array
    return rn_launch_select_nms(true, L, 0, P, g, dec, ws, Lw, s, offsets);
}

// ------------------------------------------------------------------------------------------------
// mAP matching: the per-image part of mAP1 (Vision.py:1716-1727) for every
// image, category and threshold of a validation set in one launch.  One thread per ground-truth box: among
// the image's predictions of the box's category (in prediction order) find the first maximal IoU
// (jaccard(targets, preds).max(dim=1), fp32, targets as Boxes1) and flag that prediction as correct for every
// threshold the IoU exceeds (strict >).  Several boxes may flag the same prediction; the store is idempotent.
// ------------------------------------------------------------------------------------------------
__global__ void rn_map_match_kernel(const float4 *__restrict__ pred_boxes, const int32_t *__restrict__ pred_cls,
                                    const int32_t *__restrict__ pred_off, const float4 *__restrict__ targ_boxes,
                                    const int32_t *__restrict__ targ_cls, const int32_t *__restrict__ targ_img, int NT,
                                    int NP, const float *__restrict__ thresholds, int T,
                                    unsigned char *__restrict__ is_correct /*[T][NP]*/) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NT) return;
    const float4 tb = targ_boxes[j];
    const float ta = rn_area(tb);
    const int c = targ_cls[j], img = targ_img[j];
    const int lo = pred_off[img], hi = pred_off[img + 1];
    float best = -INFINITY;
    int bi = -1;
    for (int p = lo; p < hi; ++p) {
        if (pred_cls[p] != c) continue;
        const float4 pb = pred_boxes[p];
        // Vision.jaccard(targets, preds): intersection clamped at 0, union (a_t + a_p) - inter, IEEE divide
        float iw = __fsub_rn(fminf(tb.z, pb.z), fmaxf(tb.x, pb.x));
        float ih = __fsub_rn(fminf(tb.w, pb.w), fmaxf(tb.y, pb.y));
        iw = iw > 0.0f ? iw : 0.0f;
        ih = ih > 0.0f ? ih : 0.0f;
        const float inter = __fmul_rn(iw, ih);
        const float v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ta, rn_area(pb)), inter));
        if (bi < 0 || v > best) {  // first maximal index
            best = v;
            bi = p;
        }
    }
    if (bi < 0) return;  // no prediction of this category in the image
    for (int t = 0; t < T; ++t)
        if (best > thresholds[t]) is_correct[(size_t)t * NP + bi] = 1;
}

extern "C" int rn_map_match(const float *pred_boxes, const int32_t *pred_cls, const int32_t *pred_off,
                            const float *targ_boxes, const int32_t *targ_cls, const int32_t *targ_img, int NT, int NP,
                            const float *thresholds, int T, unsigned char *is_correct, void *stream) {
    if (NT < 0 || NP < 0 || T <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_match: NT=%d NP=%d T=%d", NT, NP, T);
    cudaStream_t s = (cudaStream_t)stream;
    if (NP > 0) {
        if (!is_correct) return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_match: null output");
        cudaError_t e = cudaMemsetAsync(is_correct, 0, (size_t)T * NP, s);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_map_match memset: %s", cudaGetErrorString(e));
    }
    if (NT == 0 || NP == 0) return RN_OK;
    if (!pred_boxes || !pred_cls || !pred_off || !targ_boxes || !targ_cls || !targ_img || !thresholds)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_match: null pointer");
    if ((((uintptr_t)pred_boxes) | ((uintptr_t)targ_boxes)) & 15)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_match: boxes must be 16-byte aligned");
    rn_map_match_kernel<<<(NT + 127) / 128, 128, 0, s>>>(reinterpret_cast<const float4 *>(pred_boxes), pred_cls, pred_off,
                                                         reinterpret_cast<const float4 *>(targ_boxes), targ_cls, targ_img, NT,
                                                         NP, thresholds, T, is_correct);
    return rn_check_launch("rn_map_match");
}
