// rn_common.cuh -- shared device helpers for libretina_sm100.so (sm_100a only).
//
// Numerics contract: everything that feeds a bit-exact decision (anchor coordinates, IoU, decoded
// boxes) is written with explicit round-to-nearest intrinsics so the compiler can never contract a
// multiply-add; the library is additionally built with -fmad=false and fused multiply-adds appear
// only where fmaf() is spelled out (polynomial evaluation in the loss kernel, which is held to an
// rtol, not to bit equality).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/retina_b200.h"

#define RN_THREADS 256
#define RN_FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------
// Anchor geometry (host-built, passed by value as a __grid_constant__ kernel parameter)
// ------------------------------------------------------------------------------------------------
struct RnGeom {
    int H, W, K, A;
    int gw[RN_NUM_LEVELS];
    int gh[RN_NUM_LEVELS];
    int off[RN_NUM_LEVELS + 1];               // first anchor index of each level; off[5] = A
    int offc[RN_NUM_LEVELS + 1];              // first cell index of each level; offc[5] = number of cells
    double hw[RN_NUM_LEVELS], hh[RN_NUM_LEVELS];  // largest |x| / |y| extent of the level's base boxes
    float hwf[RN_NUM_LEVELS], hhf[RN_NUM_LEVELS]; // the same, rounded UP to float32
    double base[RN_NUM_LEVELS * RN_MAX_K * 4];  // size_l * anchor_set, float64 (retinanet.py:492)
};

// Error plumbing (rn_abi.cu)
int rn_set_error(int code, const char *fmt, ...);
int rn_check_launch(const char *what);
int rn_build_geom(RnGeom *g, int H, int W, const double *base, int K, const float *anchors, int A);

// Tuning / test switches (rn_set_option; rn_abi.cu).  All default to 0.
enum RnOption {
    RN_OPT_ASSIGN_DENSE = 0,   // != 0: rn_assign always takes the dense kernel (tests compare it with the sparse path)
    RN_OPT_ASSIGN_NO_BALANCE,  // != 0: dense kernel without the per-image work balancing
    RN_OPT_ASSIGN_WBASE,       // > 0: base weight of an image in the dense kernel's balancing (default 11)
    RN_OPT_LOSS_ITERS,         // > 0: sub-tiles per CTA of the flat loss kernel
    RN_OPT_LVL_NCHUNKS,        // > 0: class chunks per row tile of the level-tensor loss
    RN_OPT_STEP_FUSED,         // != 0: rn_loss_step runs as ONE persistent kernel (rn_step.cu) where its conditions hold
    RN_OPT_STEP_BYTEMAP,       // != 0: rn_loss_step takes its three-kernel byte-map chain instead of rn_assign + rn_loss
    RN_OPT_LOSS_PREFETCH,      // L2 prefetch of rn_loss_kernel before its PDL wait: 1 first sub-tile, 2 all (default), 3 + next wave
    RN_OPT_ASSIGN_PARTS,       // > 0: CTAs per ground-truth box of the sparse assignment kernel (default: by M * B, 1..4)
    RN_OPT_COUNT
};
int rn_opt(int id);

// Internal entry points shared by the translation units of the library (not part of the C ABI).
int rn_assign_bytes(const float *gt_boxes, const int64_t *gt_cats, int B, int M, const RnGeom &g, float pos_thr, float neg_thr,
                    uint8_t *m8, int32_t *npos_acc, int32_t *clean_list, int32_t *clean_cnt, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Loads / stores
// ------------------------------------------------------------------------------------------------
// Streaming 128-bit load: read-only path, do not allocate in L1 (each byte is touched once).
__device__ __forceinline__ float4 rn_ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// Scheduling fence for a loaded value: an empty volatile asm that "modifies" the registers.  Volatile
// asm statements keep their program order, so placing one of these per value AFTER a block of
// rn_ldg_stream calls forces every load of the block to be issued before the first consumer (ptxas
// otherwise interleaves consumers with later loads to save registers, leaving one or two requests in
// flight per warp).
__device__ __forceinline__ void rn_keep_live(float4 &v) {
    asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w));
}
// Streaming 128-bit store (evict-first in L2: the gradient is not re-read by this library).
__device__ __forceinline__ void rn_stg_stream(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining; rn_pdl_wait() blocks until
// the predecessor has completed and its writes are visible (a no-op for an ordinary launch), and
// rn_pdl_trigger() lets the successor's CTAs be scheduled as this grid's CTAs retire.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void rn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch helper: same as <<<grid, block, smem, stream>>> plus the PDL attribute.
template <typename... KArgs, typename... Args>
static inline cudaError_t rn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
// Warp / block reductions (fixed order => deterministic)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RN_FULL_MASK, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Anchors
// ------------------------------------------------------------------------------------------------
// Copies the float64 base table into shared memory (lanes of a warp read different k, which would
// serialise on the constant bank the kernel parameters live in).
__device__ __forceinline__ void rn_stage_base(const RnGeom &g, double *s_base) {
    const int n = RN_NUM_LEVELS * g.K * 4;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int l = i / (g.K * 4), r = i - l * (g.K * 4);
        s_base[i] = g.base[l * RN_MAX_K * 4 + r];
    }
}

// Anchor `a` of the level-major / row / column / slot ordering (retinanet.py:467-469, :491-495),
// computed as float32(float64 base + float64 shift) -- bit-identical to the reference's NumPy float64
// arithmetic followed by TEN()'s rounding (Core.py:61-62).
__device__ __forceinline__ float4 rn_gen_anchor(const RnGeom &g, const double *s_base, int a) {
    int l = (a >= g.off[1]) + (a >= g.off[2]) + (a >= g.off[3]) + (a >= g.off[4]);
    int local = a - g.off[l];
    int cell = local / g.K;
    int k = local - cell * g.K;
    int gw = g.gw[l];
    int iy = cell / gw;
    int ix = cell - iy * gw;
    double stride = (double)(8 << l);
    double sx = __dmul_rn((double)ix + 0.5, stride);  // retinanet.py:458 (exact)
    double sy = __dmul_rn((double)iy + 0.5, stride);  // retinanet.py:459
    const double *b = s_base + (l * g.K + k) * 4;
    float4 r;
    r.x = __double2float_rn(__dadd_rn(b[0], sx));
    r.y = __double2float_rn(__dadd_rn(b[1], sy));
    r.z = __double2float_rn(__dadd_rn(b[2], sx));
    r.w = __double2float_rn(__dadd_rn(b[3], sy));
    return r;
}

__device__ __forceinline__ float4 rn_anchor(const RnGeom &g, const double *s_base, const float4 *table, int a) {
    return table ? __ldg(table + a) : rn_gen_anchor(g, s_base, a);
}

// Same anchor, but the base table is read straight from the kernel parameter (constant bank, indexed
// load).  For kernels that need an anchor only for the rare positive rows and should not pay a
// shared-memory staging prologue in every CTA.
__device__ __forceinline__ float4 rn_anchor_from_param(const RnGeom &g, const float4 *table, int a) {
    if (table) return __ldg(table + a);
    int l = (a >= g.off[1]) + (a >= g.off[2]) + (a >= g.off[3]) + (a >= g.off[4]);
    int local = a - g.off[l];
    int cell = local / g.K;
    int k = local - cell * g.K;
    int gw = g.gw[l];
    int iy = cell / gw;
    int ix = cell - iy * gw;
    double stride = (double)(8 << l);
    double sx = __dmul_rn((double)ix + 0.5, stride);
    double sy = __dmul_rn((double)iy + 0.5, stride);
    const double *b = g.base + (l * RN_MAX_K + k) * 4;
    float4 r;
    r.x = __double2float_rn(__dadd_rn(b[0], sx));
    r.y = __double2float_rn(__dadd_rn(b[1], sy));
    r.z = __double2float_rn(__dadd_rn(b[2], sx));
    r.w = __double2float_rn(__dadd_rn(b[3], sy));
    return r;
}

// ------------------------------------------------------------------------------------------------
// IoU in strict fp32, every op rounded on its own (Vision.py:248-256, retinanet.py:506-521)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn_area(float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// IoU of two boxes given their areas; 0 when they do not overlap (the clamp(min=0) of the
// reference makes the intersection exactly 0 there, and 0/union = 0).
__device__ __forceinline__ float rn_iou(float4 p, float pa, float4 q, float qa) {
    float iw = __fsub_rn(fminf(p.z, q.z), fmaxf(p.x, q.x));
    float ih = __fsub_rn(fminf(p.w, q.w), fmaxf(p.y, q.y));
    if (!(iw > 0.0f && ih > 0.0f)) return 0.0f;
    float inter = __fmul_rn(iw, ih);
    float uni = __fsub_rn(__fadd_rn(pa, qa), inter);
    return __fdiv_rn(inter, uni);
}

// Warp-cooperative compaction of one image's ground truth into shared memory: rows with a negative
// category are padding (Vision.py:1637-1638).  Executed by one full warp; returns the count in every
// lane.  s_cat (int) and s_area may be NULL.
__device__ __forceinline__ int rn_compact_gt(const float4 *gt_boxes, const int64_t *gt_cats, int M,
                                              float4 *s_box, float *s_area, int *s_cat) {
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int j0 = 0; j0 < M; j0 += 32) {
        int j = j0 + lane;
        long long c = (j < M) ? gt_cats[j] : -1;
        bool valid = c >= 0;
        unsigned mask = __ballot_sync(RN_FULL_MASK, valid);
        if (valid) {
            int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            float4 bx = gt_boxes[j];
            s_box[pos] = bx;
            if (s_area) s_area[pos] = rn_area(bx);
            if (s_cat) s_cat[pos] = (int)c;
        }
        cnt += __popc(mask);
    }
    return cnt;
}

// float -> uint32 whose unsigned order equals the float order (for sort keys)
__device__ __forceinline__ uint32_t rn_float_sortable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float rn_sortable_float(uint32_t s) {
    uint32_t u = (s & 0x80000000u) ? (s & 0x7fffffffu) : ~s;
    return __uint_as_float(u);
}
