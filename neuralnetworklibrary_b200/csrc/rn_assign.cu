// rn_assign.cu -- anchor table, anchor x ground-truth IoU assignment, per-object max overlap.
//
// Replaces (reference file:line): AnchorGenerator.__call__ retinanet.py:485-495; Vision.jaccard
// Vision.py:234-256; match_anchors_objects Vision.py:1474-1511; the padding strip of
// SSD_loss.__call__ Vision.py:1637-1638; ComputeMaxOverlaps Vision.py:1666-1694.
//
// Kernel shape: one thread per grid cell (its K anchors in registers), the image's ground truth
// compacted + spatially culled into shared memory once per CTA and read back as warp-wide broadcasts.
// Anchors are generated on the fly (float64 add -> float32, bit-identical to the reference) or read
// from a caller-supplied table.  The IEEE divide only runs for overlapping pairs.  The work is compute
// only (reads O(M) bytes per CTA, writes 4 B per anchor).
#include "rn_common.cuh"


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
// One thread per grid CELL (all K base boxes of the cell), RN_ASSIGN_CELLS cells per CTA.
//   * the cell is decoded once (level, iy, ix: one integer division) and its K anchors are
//     float32(base + shift) with the float64 add of the reference -- 4 DADD + 4 F2F each;
//   * the image's non-padding ground truth is compacted into shared memory by one warp and culled
//     against the CTA's bounding box; each thread additionally skips a box that does not touch the
//     conservative bounding box of its own cell (both tests are exact: such a box has an intersection
//     width or height <= 0, hence IoU exactly 0, with every anchor involved);
//   * survivors keep their index among the non-padding rows, in ascending order, so the strict `>`
//     still implements torch.max's "first maximal index" (Vision.py:1505);
//   * results go through shared memory so the 4 B/anchor stores are coalesced.
// In table mode (caller-supplied anchors) every anchor is its own "cell" (K = 1).
// KT: compile-time K (0 = runtime K, anchor-outer loop without the per-cell cull).
#define RN_ASSIGN_CELLS 128

template <int KT>
__global__ void __launch_bounds__(RN_ASSIGN_CELLS, 4)
rn_assign_kernel(const float4 *__restrict__ gt_boxes, const int64_t *__restrict__ gt_cats, int M,
                 const __grid_constant__ RnGeom g, const float4 *__restrict__ table, float pos_thr,
                 float neg_thr, int32_t *__restrict__ matches, int32_t *__restrict__ npos,
                 float *__restrict__ max_iou) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: base doubles | gt boxes float4[M] | gt areas float[M] | gt index int[M] | out int[CELLS*K] | iou float[CELLS*K]
    const int K = table ? 1 : (KT ? KT : g.K);
    double *s_base = reinterpret_cast<double *>(smem);
    float4 *s_box = reinterpret_cast<float4 *>(smem + sizeof(double) * RN_NUM_LEVELS * RN_MAX_K * 4);
    float *s_area = reinterpret_cast<float *>(s_box + M);
    int *s_ci = reinterpret_cast<int *>(s_area + M);
    int *s_out = s_ci + M;
    float *s_iou = reinterpret_cast<float *>(s_out + RN_ASSIGN_CELLS * K);
    __shared__ float s_bb[4][RN_ASSIGN_CELLS / 32];
    __shared__ int s_m, s_mvalid;
    __shared__ int s_cnt[RN_ASSIGN_CELLS / 32];

    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = table ? g.A : g.offc[RN_NUM_LEVELS];
    const int cell0 = blockIdx.x * RN_ASSIGN_CELLS;
    const int c = cell0 + tid;
    const bool live = c < ncell;

    // ---- warp 0 issues the loads of the first 4 x 32 ground-truth rows right away, so their latency
    // overlaps the base-table staging and the cell decode below (the compaction itself needs the CTA
    // bounding box and happens after the barrier) ----
    constexpr int PRE = 4;
    const float4 *gbp = gt_boxes + (size_t)b * M;
    const int64_t *gcp = gt_cats + (size_t)b * M;
    long long pcat[PRE];
    float4 pbox[PRE];
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < PRE; ++i) {
            const int j = i * 32 + lane;
            pcat[i] = (j < M) ? gcp[j] : -1;
            pbox[i] = (j < M) ? gbp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (!table) rn_stage_base(g, s_base);

    // ---- decode the cell, conservative bounding box of its anchors ----
    int l = 0;
    double sx = 0.0, sy = 0.0;
    float bx1 = INFINITY, by1 = INFINITY, bx2 = -INFINITY, by2 = -INFINITY;
    float4 tan = make_float4(0.f, 0.f, 1.f, 1.f);
    if (live) {
        if (table) {
            tan = __ldg(table + c);
            bx1 = tan.x; by1 = tan.y; bx2 = tan.z; by2 = tan.w;
        } else {
            l = (c >= g.offc[1]) + (c >= g.offc[2]) + (c >= g.offc[3]) + (c >= g.offc[4]);
            const int local = c - g.offc[l];
            const int gw = g.gw[l];
            const int iy = local / gw, ix = local - iy * gw;
            const double stride = (double)(8 << l);
            sx = __dmul_rn((double)ix + 0.5, stride);  // retinanet.py:458 (exact)
            sy = __dmul_rn((double)iy + 0.5, stride);  // retinanet.py:459
            bx1 = __double2float_rd(sx - g.hw[l]);     // rounded outwards: never inside any anchor of the cell
            bx2 = __double2float_ru(sx + g.hw[l]);
            by1 = __double2float_rd(sy - g.hh[l]);
            by2 = __double2float_ru(sy + g.hh[l]);
        }
    }
    // ---- CTA bounding box -> compact + cull the image's ground truth (warp 0) ----
    float cx1 = bx1, cy1 = by1, cx2 = bx2, cy2 = by2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx1 = fminf(cx1, __shfl_xor_sync(RN_FULL_MASK, cx1, o));
        cy1 = fminf(cy1, __shfl_xor_sync(RN_FULL_MASK, cy1, o));
        cx2 = fmaxf(cx2, __shfl_xor_sync(RN_FULL_MASK, cx2, o));
        cy2 = fmaxf(cy2, __shfl_xor_sync(RN_FULL_MASK, cy2, o));
    }
    if (lane == 0) {
        s_bb[0][warp] = cx1; s_bb[1][warp] = cy1; s_bb[2][warp] = cx2; s_bb[3][warp] = cy2;
    }
    __syncthreads();  // s_bb and the staged base table
    if (warp == 0) {
        cx1 = cy1 = INFINITY;
        cx2 = cy2 = -INFINITY;
#pragma unroll
        for (int w = 0; w < RN_ASSIGN_CELLS / 32; ++w) {
            cx1 = fminf(cx1, s_bb[0][w]); cy1 = fminf(cy1, s_bb[1][w]);
            cx2 = fmaxf(cx2, s_bb[2][w]); cy2 = fmaxf(cy2, s_bb[3][w]);
        }
        int nvalid = 0, nkeep = 0;
        auto take = [&](long long cat, float4 bx, bool inrange) {
            const bool valid = inrange && (cat >= 0);  // padding rows have a negative category
            const unsigned vmask = __ballot_sync(RN_FULL_MASK, valid);
            const bool keep = valid && (bx.z > cx1) && (bx.x < cx2) && (bx.w > cy1) && (bx.y < cy2);
            const unsigned kmask = __ballot_sync(RN_FULL_MASK, keep);
            if (keep) {
                const int pos = nkeep + __popc(kmask & ((1u << lane) - 1u));
                s_box[pos] = bx;
                s_area[pos] = rn_area(bx);
                s_ci[pos] = nvalid + __popc(vmask & ((1u << lane) - 1u));
            }
            nvalid += __popc(vmask);
            nkeep += __popc(kmask);
        };
#pragma unroll
        for (int i = 0; i < PRE; ++i)
            if (i * 32 < M) take(pcat[i], pbox[i], i * 32 + lane < M);
        for (int j0 = PRE * 32; j0 < M; j0 += 32) {  // beyond the prefetched rows
            const int j = j0 + lane;
            take((j < M) ? gcp[j] : -1, (j < M) ? gbp[j] : make_float4(0.f, 0.f, 0.f, 0.f), j < M);
        }
        if (lane == 0) {
            s_m = nkeep;
            s_mvalid = nvalid;
        }
    }
    __syncthreads();
    const int m = s_m, mvalid = s_mvalid;

    int cnt = 0;
    auto finish = [&](int k, float best, int bi) {
        int mt;
        if (mvalid == 0) mt = RN_MATCH_NEG;             // Vision.py:1498-1501
        else if (best > pos_thr) mt = bi;               // Vision.py:1506, :1508-1509
        else if (best < neg_thr) mt = RN_MATCH_NEG;     // Vision.py:1507
        else mt = RN_MATCH_IGNORE;
        s_out[tid * K + k] = mt;
        if (max_iou) s_iou[tid * K + k] = best;
        cnt += (mt >= 0);
    };
    auto pair = [&](const float4 &an, float aa, const float4 &gb, float ga, int ci, float &best, int &bi) {
        const float iw = __fsub_rn(fminf(gb.z, an.z), fmaxf(gb.x, an.x));
        const float ih = __fsub_rn(fminf(gb.w, an.w), fmaxf(gb.y, an.y));
        if (iw > 0.0f && ih > 0.0f) {
            const float inter = __fmul_rn(iw, ih);
            const float uni = __fsub_rn(__fadd_rn(ga, aa), inter);  // Vision.py:255
            const float v = __fdiv_rn(inter, uni);
            if (v > best) {  // strict: first maximal index wins (torch.max, Vision.py:1505)
                best = v;
                bi = ci;
            }
        }
    };

    if (live) {
        if (KT > 0 && !table) {
            // all K anchors of the cell in registers; ground truth outer, anchors inner
            float4 an[KT > 0 ? KT : 1];
            float aa[KT > 0 ? KT : 1], best[KT > 0 ? KT : 1];
            int bi[KT > 0 ? KT : 1];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                const double *bb = s_base + (l * KT + k) * 4;
                an[k].x = __double2float_rn(__dadd_rn(bb[0], sx));
                an[k].y = __double2float_rn(__dadd_rn(bb[1], sy));
                an[k].z = __double2float_rn(__dadd_rn(bb[2], sx));
                an[k].w = __double2float_rn(__dadd_rn(bb[3], sy));
                aa[k] = rn_area(an[k]);
                best[k] = 0.0f;  // IoU >= 0 and torch.max returns index 0 for an all-zero column
                bi[k] = 0;
            }
            for (int j = 0; j < m; ++j) {
                const float4 gb = s_box[j];  // broadcast
                if (gb.z > bx1 && gb.x < bx2 && gb.w > by1 && gb.y < by2) {
                    const float ga = s_area[j];
                    const int ci = s_ci[j];
#pragma unroll
                    for (int k = 0; k < KT; ++k) pair(an[k], aa[k], gb, ga, ci, best[k], bi[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) finish(k, best[k], bi[k]);
        } else {
            for (int k = 0; k < K; ++k) {
                float4 an = tan;
                if (!table) {
                    const double *bb = s_base + (l * K + k) * 4;
                    an.x = __double2float_rn(__dadd_rn(bb[0], sx));
                    an.y = __double2float_rn(__dadd_rn(bb[1], sy));
                    an.z = __double2float_rn(__dadd_rn(bb[2], sx));
                    an.w = __double2float_rn(__dadd_rn(bb[3], sy));
                }
                const float aa = rn_area(an);
                float best = 0.0f;
                int bi = 0;
                for (int j = 0; j < m; ++j) pair(an, aa, s_box[j], s_area[j], s_ci[j], best, bi);
                finish(k, best, bi);
            }
        }
    }
    cnt = __reduce_add_sync(RN_FULL_MASK, cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();

    // ---- coalesced copy-out of the CTA's contiguous anchor range ----
    const int A = g.A;
    const int a0 = cell0 * K;
    const int na = min(RN_ASSIGN_CELLS * K, A - a0);
    for (int i = tid; i < na; i += RN_ASSIGN_CELLS) {
        matches[(size_t)b * A + a0 + i] = s_out[i];
        if (max_iou) max_iou[(size_t)b * A + a0 + i] = s_iou[i];
    }
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < RN_ASSIGN_CELLS / 32; ++w) t += s_cnt[w];
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
    const int Kc = anchors ? 1 : K;
    const int ncell = anchors ? A : g.offc[RN_NUM_LEVELS];
    size_t smem = kBaseBytes + (size_t)M * (sizeof(float4) + sizeof(float) + sizeof(int)) +
                  (size_t)RN_ASSIGN_CELLS * Kc * (sizeof(int) + sizeof(float));
    if (smem > 200 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_assign: M=%d too large", M);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(npos, 0, sizeof(int32_t) * (size_t)B, s);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign memset: %s", cudaGetErrorString(e));
    const bool k9 = !anchors && K == 9;  // the reference's 3 ratios x 3 scales
    if (smem > 48 * 1024) {
        e = k9 ? cudaFuncSetAttribute(rn_assign_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
               : cudaFuncSetAttribute(rn_assign_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign smem: %s", cudaGetErrorString(e));
    }
    dim3 grid((ncell + RN_ASSIGN_CELLS - 1) / RN_ASSIGN_CELLS, B);
    const float4 *gb4 = reinterpret_cast<const float4 *>(gt_boxes), *tb4 = reinterpret_cast<const float4 *>(anchors);
    if (k9) rn_assign_kernel<9><<<grid, RN_ASSIGN_CELLS, smem, s>>>(gb4, gt_cats, M, g, tb4, pos_thr, neg_thr, matches, npos, max_iou);
    else rn_assign_kernel<0><<<grid, RN_ASSIGN_CELLS, smem, s>>>(gb4, gt_cats, M, g, tb4, pos_thr, neg_thr, matches, npos, max_iou);
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
