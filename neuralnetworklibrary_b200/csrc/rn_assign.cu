// rn_assign.cu -- anchor table, anchor x ground-truth IoU assignment, per-object max overlap.
//
// Replaces (reference file:line): AnchorGenerator.__call__ retinanet.py:485-495; Vision.jaccard
// Vision.py:234-256; match_anchors_objects Vision.py:1474-1511; the padding strip of
// SSD_loss.__call__ Vision.py:1637-1638; ComputeMaxOverlaps Vision.py:1666-1694.
//
// Kernel shape: one thread per anchor (APT anchors per thread for ILP), the image's ground truth
// compacted into shared memory once per CTA and read back as warp-wide broadcasts.  Anchors are
// generated on the fly (float64 add -> float32, bit-identical to the reference) or read from a
// caller-supplied table.  The IEEE divide only runs for overlapping pairs.  The work is compute
// only (reads O(M) bytes per CTA, writes 4 B per anchor).
#include "rn_common.cuh"

#define RN_ASSIGN_APT 2

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RN_THREADS)
rn_anchors_kernel(const __grid_constant__ RnGeom g, float4 *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_base = reinterpret_cast<double *>(smem);
    rn_stage_base(g, s_base);
    __syncthreads();
    int a = blockIdx.x * RN_THREADS + threadIdx.x;
    if (a < g.A) out[a] = rn_gen_anchor(g, s_base, a);
}

// ------------------------------------------------------------------------------------------------
// Spatial culling: a CTA owns 512 consecutive anchors (a few dozen neighbouring cells of one pyramid
// level).  Ground-truth boxes that do not touch the CTA's bounding box have IoU exactly 0 with every
// anchor of the CTA (the intersection width or height is <= 0), so they are dropped while the image's
// ground truth is compacted into shared memory; the survivors keep their index among the non-padding
// rows, in ascending order, so "first maximal index" is unchanged.
__global__ void __launch_bounds__(RN_THREADS)
rn_assign_kernel(const float4 *__restrict__ gt_boxes, const int64_t *__restrict__ gt_cats, int M,
                 const __grid_constant__ RnGeom g, const float4 *__restrict__ table, float pos_thr,
                 float neg_thr, int32_t *__restrict__ matches, int32_t *__restrict__ npos,
                 float *__restrict__ max_iou) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: base doubles | gt boxes float4[M] | gt areas float[M] | gt index int[M]
    double *s_base = reinterpret_cast<double *>(smem);
    float4 *s_box = reinterpret_cast<float4 *>(smem + sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4);
    float *s_area = reinterpret_cast<float *>(s_box + M);
    int *s_ci = reinterpret_cast<int *>(s_area + M);
    __shared__ float s_bb[4][RN_THREADS / 32];
    __shared__ int s_m, s_mvalid;
    __shared__ int s_cnt[RN_THREADS / 32];

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = g.A;
    const int a0 = blockIdx.x * (RN_THREADS * RN_ASSIGN_APT) + tid;
    if (!table) {
        rn_stage_base(g, s_base);
        __syncthreads();
    }

    float4 an[RN_ASSIGN_APT];
    float aa[RN_ASSIGN_APT], best[RN_ASSIGN_APT];
    int bi[RN_ASSIGN_APT];
    float bx1 = INFINITY, by1 = INFINITY, bx2 = -INFINITY, by2 = -INFINITY;
#pragma unroll
    for (int i = 0; i < RN_ASSIGN_APT; ++i) {
        const int a = a0 + i * RN_THREADS;
        if (a < A) {
            an[i] = rn_anchor(g, s_base, table, a);
            bx1 = fminf(bx1, an[i].x);
            by1 = fminf(by1, an[i].y);
            bx2 = fmaxf(bx2, an[i].z);
            by2 = fmaxf(by2, an[i].w);
        } else {
            an[i] = make_float4(0.f, 0.f, 1.f, 1.f);
        }
        aa[i] = rn_area(an[i]);
        best[i] = 0.0f;  // IoU >= 0 and torch.max returns index 0 for an all-zero column
        bi[i] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bx1 = fminf(bx1, __shfl_xor_sync(RN_FULL_MASK, bx1, o));
        by1 = fminf(by1, __shfl_xor_sync(RN_FULL_MASK, by1, o));
        bx2 = fmaxf(bx2, __shfl_xor_sync(RN_FULL_MASK, bx2, o));
        by2 = fmaxf(by2, __shfl_xor_sync(RN_FULL_MASK, by2, o));
    }
    if (lane == 0) {
        s_bb[0][warp] = bx1;
        s_bb[1][warp] = by1;
        s_bb[2][warp] = bx2;
        s_bb[3][warp] = by2;
    }
    __syncthreads();
    if (warp == 0) {
        bx1 = by1 = INFINITY;
        bx2 = by2 = -INFINITY;
#pragma unroll
        for (int w = 0; w < RN_THREADS / 32; ++w) {
            bx1 = fminf(bx1, s_bb[0][w]);
            by1 = fminf(by1, s_bb[1][w]);
            bx2 = fmaxf(bx2, s_bb[2][w]);
            by2 = fmaxf(by2, s_bb[3][w]);
        }
        const float4 *gb = gt_boxes + (size_t)b * M;
        const int64_t *gc = gt_cats + (size_t)b * M;
        int nvalid = 0, nkeep = 0;
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = (j < M) && (gc[j] >= 0);  // padding rows have a negative category
            const unsigned vmask = __ballot_sync(RN_FULL_MASK, valid);
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            bool keep = false;
            if (valid) {
                bx = gb[j];
                keep = (bx.z > bx1) && (bx.x < bx2) && (bx.w > by1) && (bx.y < by2);
            }
            const unsigned kmask = __ballot_sync(RN_FULL_MASK, keep);
            if (keep) {
                const int pos = nkeep + __popc(kmask & ((1u << lane) - 1u));
                s_box[pos] = bx;
                s_area[pos] = rn_area(bx);
                s_ci[pos] = nvalid + __popc(vmask & ((1u << lane) - 1u));
            }
            nvalid += __popc(vmask);
            nkeep += __popc(kmask);
        }
        if (lane == 0) {
            s_m = nkeep;
            s_mvalid = nvalid;
        }
    }
    __syncthreads();
    const int m = s_m, mvalid = s_mvalid;

    for (int j = 0; j < m; ++j) {
        const float4 gb = s_box[j];  // broadcast
#pragma unroll
        for (int i = 0; i < RN_ASSIGN_APT; ++i) {
            float iw = __fsub_rn(fminf(gb.z, an[i].z), fmaxf(gb.x, an[i].x));
            float ih = __fsub_rn(fminf(gb.w, an[i].w), fmaxf(gb.y, an[i].y));
            if (iw > 0.0f && ih > 0.0f) {
                float inter = __fmul_rn(iw, ih);
                float uni = __fsub_rn(__fadd_rn(s_area[j], aa[i]), inter);  // Vision.py:255
                float v = __fdiv_rn(inter, uni);
                if (v > best[i]) {  // strict: first maximal index wins (torch.max, Vision.py:1505)
                    best[i] = v;
                    bi[i] = s_ci[j];
                }
            }
        }
    }
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < RN_ASSIGN_APT; ++i) {
        int a = a0 + i * RN_THREADS;
        if (a < A) {
            int mt;
            if (mvalid == 0) mt = RN_MATCH_NEG;              // Vision.py:1498-1501
            else if (best[i] > pos_thr) mt = bi[i];          // Vision.py:1506, :1508-1509
            else if (best[i] < neg_thr) mt = RN_MATCH_NEG;   // Vision.py:1507
            else mt = RN_MATCH_IGNORE;
            matches[(size_t)b * A + a] = mt;
            if (max_iou) max_iou[(size_t)b * A + a] = best[i];
            cnt += (mt >= 0);
        }
    }
    cnt = __reduce_add_sync(RN_FULL_MASK, cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < RN_THREADS / 32; ++w) t += s_cnt[w];
        if (t) atomicAdd(npos + b, t);  // integer: order independent
    }
}

// ------------------------------------------------------------------------------------------------
// Per ground-truth row: max IoU over all anchors (jac.max(dim=1), Vision.py:1686-1687).
__global__ void rn_max_overlaps_init_kernel(const int64_t *__restrict__ gt_cats, int n, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gt_cats[i] >= 0 ? 0.0f : -1.0f;
}

__global__ void __launch_bounds__(RN_THREADS)
rn_max_overlaps_kernel(const float4 *__restrict__ gt_boxes, const int64_t *__restrict__ gt_cats, int M,
                       const __grid_constant__ RnGeom g, const float4 *__restrict__ table,
                       float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_base = reinterpret_cast<double *>(smem);
    int *s_max = reinterpret_cast<int *>(smem + sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4);
    const int b = blockIdx.y, tid = threadIdx.x;
    if (!table) rn_stage_base(g, s_base);
    for (int j = tid; j < M; j += RN_THREADS) s_max[j] = 0;
    __syncthreads();
    const int a = blockIdx.x * RN_THREADS + tid;
    const bool live = a < g.A;
    float4 an = live ? rn_anchor(g, s_base, table, a) : make_float4(0.f, 0.f, 0.f, 0.f);
    float aa = rn_area(an);
    for (int j = 0; j < M; ++j) {
        if (gt_cats[(size_t)b * M + j] < 0) continue;  // uniform across the CTA
        float4 gb = gt_boxes[(size_t)b * M + j];
        float v = live ? rn_iou(gb, rn_area(gb), an, aa) : 0.0f;
        // IoU >= 0, so the int order of the bit pattern is the float order
        int vm = __reduce_max_sync(RN_FULL_MASK, __float_as_int(v));
        if ((tid & 31) == 0 && vm > 0) atomicMax(s_max + j, vm);
    }
    __syncthreads();
    for (int j = tid; j < M; j += RN_THREADS)
        if (s_max[j] > 0) atomicMax(reinterpret_cast<int *>(out) + (size_t)b * M + j, s_max[j]);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static const size_t kBaseBytes = sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4;

extern "C" int rn_anchors(int H, int W, const double *base, int K, float *anchors_out, void *stream) {
    if (!anchors_out) return rn_set_error(RN_ERR_INVALID_ARG, "rn_anchors: null output");
    RnGeom g;
    int A = rn_num_anchors(H, W, K);
    int rc = rn_build_geom(&g, H, W, base, K, nullptr, A);
    if (rc) return rc;
    rn_anchors_kernel<<<(A + RN_THREADS - 1) / RN_THREADS, RN_THREADS, kBaseBytes, (cudaStream_t)stream>>>(
        g, reinterpret_cast<float4 *>(anchors_out));
    return rn_check_launch("rn_anchors");
}

extern "C" int rn_assign(const float *gt_boxes, const int64_t *gt_cats, int B, int M, int H, int W,
                         const double *base, int K, const float *anchors, int A, float pos_thr, float neg_thr,
                         int32_t *matches, int32_t *npos, float *max_iou, void *stream) {
    if (B < 0 || M < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_assign: B=%d M=%d", B, M);
    if (B == 0) return RN_OK;
    if (!matches || !npos || (M > 0 && (!gt_boxes || !gt_cats)))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_assign: null pointer");
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, anchors, A);
    if (rc) return rc;
    size_t smem = kBaseBytes + (size_t)M * (sizeof(float4) + sizeof(float) + sizeof(int));
    if (smem > 200 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_assign: M=%d too large", M);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(npos, 0, sizeof(int32_t) * (size_t)B, s);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign memset: %s", cudaGetErrorString(e));
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(rn_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign smem: %s", cudaGetErrorString(e));
    }
    const int per_cta = RN_THREADS * RN_ASSIGN_APT;
    dim3 grid((A + per_cta - 1) / per_cta, B);
    rn_assign_kernel<<<grid, RN_THREADS, smem, s>>>(reinterpret_cast<const float4 *>(gt_boxes), gt_cats, M, g,
                                                    reinterpret_cast<const float4 *>(anchors), pos_thr, neg_thr,
                                                    matches, npos, max_iou);
    return rn_check_launch("rn_assign");
}

extern "C" int rn_max_overlaps(const float *gt_boxes, const int64_t *gt_cats, int B, int M, int H, int W,
                               const double *base, int K, const float *anchors, int A, float *out, void *stream) {
    if (B <= 0 || M <= 0) return RN_OK;
    if (!gt_boxes || !gt_cats || !out) return rn_set_error(RN_ERR_INVALID_ARG, "rn_max_overlaps: null pointer");
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, anchors, A);
    if (rc) return rc;
    size_t smem = kBaseBytes + (size_t)M * sizeof(int);
    if (smem > 48 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_max_overlaps: M=%d too large", M);
    cudaStream_t s = (cudaStream_t)stream;
    int n = B * M;
    rn_max_overlaps_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(gt_cats, n, out);
    dim3 grid((A + RN_THREADS - 1) / RN_THREADS, B);
    rn_max_overlaps_kernel<<<grid, RN_THREADS, smem, s>>>(reinterpret_cast<const float4 *>(gt_boxes), gt_cats, M, g,
                                                          reinterpret_cast<const float4 *>(anchors), out);
    return rn_check_launch("rn_max_overlaps");
}
