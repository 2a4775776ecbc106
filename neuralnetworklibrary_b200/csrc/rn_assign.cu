// rn_assign.cu -- anchor table, anchor x ground-truth IoU assignment, per-object max overlap.
//
// Replaces (reference file:line): AnchorGenerator.__call__ retinanet.py:485-495; Vision.jaccard
// Vision.py:234-256; match_anchors_objects Vision.py:1474-1511; the padding strip of
// SSD_loss.__call__ Vision.py:1637-1638; ComputeMaxOverlaps Vision.py:1666-1694.
//
// Kernel shape: SPLIT threads per grid cell (its K anchors in registers), persistent-style CTAs, the
// image's ground truth compacted into shared memory once per CTA and read back as warp-wide broadcasts.
// Anchors are generated on the fly (float64 add -> float32, bit-identical to the reference) or read
// from a caller-supplied table.  The IEEE divide only runs for overlapping pairs.  The work is compute
// only (reads O(M) bytes per CTA, writes 4 B per anchor).
#include <string.h>

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
// Assignment kernel.  SPLIT threads per grid CELL, each holding KT / SPLIT of the cell's K base boxes
// in registers; persistent CTAs, each working on ONE image, whose warps walk warp-tiles round-robin.
//   * prologue, once per CTA: the float64 base table is staged in shared memory and one warp compacts
//     the image's non-padding ground truth (rows with a negative category are padding,
//     Vision.py:1637-1638) into shared memory.  Earlier versions did this once per 128 cells; the kernel
//     was then bound by the latency of that serial prologue times the number of CTA waves (31-40 us for
//     COCO B=16 regardless of instruction count, profiles/r01_summary.md);
//   * per tile, no barrier at all: the cell is decoded (level, iy, ix: one integer division), its
//     anchors are float32(base + shift) with the float64 add of the reference (4 DADD + 4 F2F each), and
//     every ground-truth box that does not touch the cell's conservative bounding box is skipped -- an
//     exact test: such a box has an intersection width or height <= 0, hence IoU exactly 0, with every
//     anchor of the cell;
//   * boxes are visited in ascending index, so the strict `>` implements torch.max's "first maximal
//     index" (Vision.py:1505); the IEEE divide only runs for overlapping pairs;
//   * each thread stores its anchors' results directly: a warp writes one contiguous 4*32*KPT-byte run.
// In table mode (caller-supplied anchors) every anchor is its own "cell" (K = 1, KT = 0).
#define RN_ASSIGN_CELLS 64
#ifndef RN_ASSIGN_CTAS
#define RN_ASSIGN_CTAS 5  // resident CTAs per SM (register bound: 64 registers per thread)
#endif

template <int KT, int SPLIT>
__global__ void __launch_bounds__(RN_ASSIGN_CELLS * SPLIT, RN_ASSIGN_CTAS)
rn_assign_kernel(const float4 *__restrict__ gt_boxes, const int64_t *__restrict__ gt_cats, int M,
                 const __grid_constant__ RnGeom g, const float4 *__restrict__ table, float pos_thr,
                 float neg_thr, int32_t *__restrict__ matches, int32_t *__restrict__ npos,
                 float *__restrict__ max_iou, int B, int w_base, int w_box) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: gt boxes float4[M] | gt areas float[M].  The float64 base table is read straight from the
    // kernel parameters (constant bank): staging it in shared memory with per-lane indices serialises on
    // the constant cache and cost every CTA several microseconds (profiles/r01_summary.md).
    constexpr int NTHR = RN_ASSIGN_CELLS * SPLIT;
    constexpr int KPT = KT > 0 ? KT / SPLIT : 1;  // anchors per thread on the register path
    const int K = table ? 1 : (KT ? KT : g.K);
    float4 *s_box = reinterpret_cast<float4 *>(smem);
    float *s_area = reinterpret_cast<float *>(s_box + M);
    int *s_w = reinterpret_cast<int *>(s_area + M);  // [B] per-image work estimate (balanced mode only)
    __shared__ int s_mvalid;
    __shared__ int s_cnt[NTHR / 32];
    __shared__ int s_part[3];  // image, rank of this CTA inside the image, CTAs of the image

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = table ? g.A : g.offc[RN_NUM_LEVELS];
    const int A = g.A;

    rn_pdl_trigger();  // the loss kernel that follows may start filling SMs as this grid's CTAs retire
    // ---- which image does this CTA work on? ----
    // gridDim.y == B: a fixed share of gridDim.x CTAs per image.  gridDim.y == 1 (balanced mode, gridDim.x >= B): the
    // cost of an image grows with its number of ground-truth boxes (an image without objects is ~3x cheaper than one
    // with 20), so every CTA counts the boxes of all images -- B*M category reads, L2 hits -- and derives the same
    // partition: one CTA per image, the remaining ones in proportion to w_base + w_box * boxes.
    int b = blockIdx.y, rank = blockIdx.x, nshare = gridDim.x;
    if (gridDim.y == 1 && B > 1) {
        for (int i = tid; i < B; i += NTHR) s_w[i] = 0;
        __syncthreads();
        for (int i = tid; i < B * M; i += NTHR)
            if (gt_cats[i] >= 0) atomicAdd(s_w + i / M, w_box);
        __syncthreads();
        if (warp == 0) {
            long long total = 0;
            for (int b0 = 0; b0 < B; b0 += 32) {
                const int w = (b0 + lane < B) ? s_w[b0 + lane] + w_base : 0;
                total += __reduce_add_sync(RN_FULL_MASK, w);
            }
            const long long extra = (long long)gridDim.x - B;  // CTAs beyond the first of every image
            const long long e = (long long)blockIdx.x - B;
            long long carry = 0;
            for (int b0 = 0; b0 < B; b0 += 32) {
                const int bb = b0 + lane;
                const int w = (bb < B) ? s_w[bb] + w_base : 0;
                int incl = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(RN_FULL_MASK, incl, o);
                    if (lane >= o) incl += t;
                }
                const long long p0 = carry + incl - w, p1 = carry + incl;
                const long long st = extra * p0 / total, en = extra * p1 / total;
                if (bb < B) {
                    if ((long long)blockIdx.x == bb) {  // the image's first CTA
                        s_part[0] = bb;
                        s_part[1] = 0;
                        s_part[2] = 1 + (int)(en - st);
                    } else if (e >= st && e < en) {
                        s_part[0] = bb;
                        s_part[1] = 1 + (int)(e - st);
                        s_part[2] = 1 + (int)(en - st);
                    }
                }
                carry += __shfl_sync(RN_FULL_MASK, incl, 31);
            }
        }
        __syncthreads();
        b = s_part[0];
        rank = s_part[1];
        nshare = s_part[2];
    }
    if (warp == 0) {
        const int mv = rn_compact_gt(gt_boxes + (size_t)b * M, gt_cats + (size_t)b * M, M, s_box, s_area, nullptr);
        if (lane == 0) s_mvalid = mv;
    }
    __syncthreads();
    const int m = s_mvalid;

    int cnt = 0;
    auto classify = [&](float best, int bi) -> int {
        if (m == 0) return RN_MATCH_NEG;             // Vision.py:1498-1501
        if (best > pos_thr) return bi;               // Vision.py:1506, :1508-1509
        if (best < neg_thr) return RN_MATCH_NEG;     // Vision.py:1507
        return RN_MATCH_IGNORE;
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

    // Work unit = one warp-tile of 32 consecutive (cell, anchor-group) slots.  Warp-tiles are dealt
    // round-robin over all warps that work on this image: the tiles of the coarse pyramid levels, whose huge
    // anchors overlap every ground-truth box and cost ~20x a P3 tile, end up on different warps.  With
    // contiguous per-CTA ranges one CTA per image owned all of them and the whole kernel waited for it (SMs
    // 45 % idle, profiles/r01_summary.md).  Nothing inside the loop needs a CTA barrier.
    const int nslots = table ? ncell : ncell * SPLIT;
    const int nwt = (nslots + 31) >> 5;
    const int gwarps = nshare * (NTHR / 32);
    // Heaviest tiles first (they are the last in index order): the cheap P3 tiles then fill the tail evenly.
#pragma unroll 1
    for (int wt = nwt - 1 - (rank * (NTHR / 32) + warp); wt >= 0; wt -= gwarps) {
        const int slot = (wt << 5) + lane;
        const int c = table ? slot : slot / SPLIT;
        const int part = table ? 0 : slot - c * SPLIT;  // anchor group inside the cell
        const bool live = c < ncell;
        if (KT > 0 && !table) {
            // decode the cell; conservative bounding box of all its anchors (rounded outwards).  Lanes past
            // the last cell keep going with the last cell's geometry but an empty box: the warp-wide
            // ballot / shuffles below need every lane, and those lanes store nothing.
            const int cc = live ? c : ncell - 1;
            const int l = (cc >= g.offc[1]) + (cc >= g.offc[2]) + (cc >= g.offc[3]) + (cc >= g.offc[4]);
            const int local = cc - g.offc[l];
            const int gw = g.gw[l];
            const int iy = local / gw, ix = local - iy * gw;
            // The cell's conservative bounding box in fp32 only: the cell centre (i + 0.5) * stride is
            // exact in fp32, the level's largest half extent is rounded up, and the subtraction / addition
            // round outwards -- so the box contains every anchor of the cell, and no float64 is touched
            // unless a ground-truth box actually reaches the cell.
            const float strf = (float)(8 << l);
            const float sxf = ((float)ix + 0.5f) * strf, syf = ((float)iy + 0.5f) * strf;
            const float bx1 = live ? __fsub_rd(sxf, g.hwf[l]) : INFINITY;
            const float bx2 = live ? __fadd_ru(sxf, g.hwf[l]) : -INFINITY;
            const float by1 = live ? __fsub_rd(syf, g.hhf[l]) : INFINITY;
            const float by2 = live ? __fadd_ru(syf, g.hhf[l]) : -INFINITY;
            float4 an[KPT];
            float aa[KPT], best[KPT];
            int bi[KPT];
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                an[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                aa[i] = 0.0f;
                best[i] = 0.0f;  // IoU >= 0 and torch.max returns index 0 for an all-zero column
                bi[i] = 0;
            }
            bool have_anchors = false;  // generated lazily, on the first box that reaches the cell
            // Warp-level cull: lane j tests ground-truth box j against the bounding box of the warp's
            // ~11 neighbouring cells; only boxes that touch it are visited (ascending index), and each
            // thread still skips a box that misses its own cell.  Typically 1-3 of M boxes survive.
            float wx1 = bx1, wy1 = by1, wx2 = bx2, wy2 = by2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wx1 = fminf(wx1, __shfl_xor_sync(RN_FULL_MASK, wx1, o));
                wy1 = fminf(wy1, __shfl_xor_sync(RN_FULL_MASK, wy1, o));
                wx2 = fmaxf(wx2, __shfl_xor_sync(RN_FULL_MASK, wx2, o));
                wy2 = fmaxf(wy2, __shfl_xor_sync(RN_FULL_MASK, wy2, o));
            }
            for (int j0 = 0; j0 < m; j0 += 32) {
                bool touch = false;
                if (j0 + lane < m) {
                    const float4 gq = s_box[j0 + lane];
                    touch = gq.z > wx1 && gq.x < wx2 && gq.w > wy1 && gq.y < wy2;
                }
                unsigned cand = __ballot_sync(RN_FULL_MASK, touch);
                while (cand) {
                    const int j = j0 + __ffs(cand) - 1;
                    cand &= cand - 1;
                    const float4 gb = s_box[j];  // broadcast
                    if (gb.z > bx1 && gb.x < bx2 && gb.w > by1 && gb.y < by2) {
                        if (!have_anchors) {
                            have_anchors = true;
                            const double stride = (double)(8 << l);
                            const double sx = __dmul_rn((double)ix + 0.5, stride);  // retinanet.py:458 (exact)
                            const double sy = __dmul_rn((double)iy + 0.5, stride);  // retinanet.py:459
#pragma unroll
                            for (int i = 0; i < KPT; ++i) {
                                const double *bb = g.base + (l * RN_MAX_K + part * KPT + i) * 4;
                                an[i].x = __double2float_rn(__dadd_rn(bb[0], sx));
                                an[i].y = __double2float_rn(__dadd_rn(bb[1], sy));
                                an[i].z = __double2float_rn(__dadd_rn(bb[2], sx));
                                an[i].w = __double2float_rn(__dadd_rn(bb[3], sy));
                                aa[i] = rn_area(an[i]);
                            }
                        }
                        const float ga = s_area[j];
#pragma unroll
                        for (int i = 0; i < KPT; ++i) pair(an[i], aa[i], gb, ga, j, best[i], bi[i]);
                    }
                }
            }
            const size_t o = (size_t)b * A + (size_t)cc * KT + part * KPT;
#pragma unroll
            for (int i = 0; i < KPT && live; ++i) {
                const int mt = classify(best[i], bi[i]);
                matches[o + i] = mt;
                if (max_iou) max_iou[o + i] = best[i];
                cnt += (mt >= 0);
            }
        } else if (part == 0 && live) {
            int l = 0;
            double sx = 0.0, sy = 0.0;
            if (!table) {
                l = (c >= g.offc[1]) + (c >= g.offc[2]) + (c >= g.offc[3]) + (c >= g.offc[4]);
                const int local = c - g.offc[l];
                const int gw = g.gw[l];
                const int iy = local / gw, ix = local - iy * gw;
                const double stride = (double)(8 << l);
                sx = __dmul_rn((double)ix + 0.5, stride);
                sy = __dmul_rn((double)iy + 0.5, stride);
            }
            for (int k = 0; k < K; ++k) {
                float4 an;
                if (table) {
                    an = __ldg(table + c);
                } else {
                    const double *bb = g.base + (l * RN_MAX_K + k) * 4;
                    an.x = __double2float_rn(__dadd_rn(bb[0], sx));
                    an.y = __double2float_rn(__dadd_rn(bb[1], sy));
                    an.z = __double2float_rn(__dadd_rn(bb[2], sx));
                    an.w = __double2float_rn(__dadd_rn(bb[3], sy));
                }
                const float aa = rn_area(an);
                float best = 0.0f;
                int bi = 0;
                for (int j = 0; j < m; ++j) pair(an, aa, s_box[j], s_area[j], j, best, bi);
                const int mt = classify(best, bi);
                const size_t o = (size_t)b * A + (size_t)c * K + k;
                matches[o] = mt;
                if (max_iou) max_iou[o] = best;
                cnt += (mt >= 0);
            }
        }
    }
    cnt = __reduce_add_sync(RN_FULL_MASK, cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < NTHR / 32; ++w) t += s_cnt[w];
        if (t) atomicAdd(npos + b, t);  // integer: order independent
    }
}

// ------------------------------------------------------------------------------------------------
// Sparse assignment.  An anchor is background unless some ground-truth box reaches IoU >= neg_thr with it, and the
// anchors that can do so with ONE box are few: their area must be within [neg_thr, 1/neg_thr] of the box's and their
// centre close to the box's centre -- a few hundred of the 200 k anchors of a COCO-sized image.  So `matches` is filled
// with RN_MATCH_NEG by a memset and one CTA per ground-truth box enumerates that box's candidate window on every
// pyramid level; each candidate anchor is then evaluated EXACTLY like in the dense kernel (float32(float64 base + shift),
// strict-fp32 IoU against all boxes of the image in ascending order, first maximal index), and written by the CTA of
// its argmax box only, so every non-background anchor is written and counted exactly once without atomics on matches.
//
// Why the window is conservative.  IoU = inter/(Aa + Ag - inter) >= t  <=>  inter >= t/(1+t) * (Aa + Ag) =: need (which
// also requires min(A) >= t*max(A)); with inter = iw*ih and ih <= min(ha, hg) this gives iw >= need/min(ha, hg) =: iw_min,
// and since iw <= (wa+wg)/2 - |dcx| always, |dcx| <= (wa+wg)/2 - iw_min (same in y).  One window per (level, base box);
// the kernel uses t = 0.95*neg_thr in float64 and 0.01 cell of slack, which dwarfs every rounding involved (the float32
// rounding of the anchor coordinates is ~1e-5 px).  Anchors outside the window have IoU < neg_thr with this box, so it
// is neither their argmax above a threshold nor able to lift them out of "background".
// ------------------------------------------------------------------------------------------------
#define RN_SPARSE_THREADS 256

// Every anchor background, every positive count zero (replaces two memset nodes; the sparse kernel behind it is
// launched with PDL and computes its prologue while this one runs).
__global__ void __launch_bounds__(256)
rn_assign_fill_kernel(int32_t *__restrict__ matches, size_t n, int32_t *__restrict__ npos, int B) {
    rn_pdl_trigger();
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (i0 < (size_t)B) npos[i0] = 0;  // B <= gridDim.x * 256 is checked by the host
    if ((((uintptr_t)matches) & 15) == 0) {
        int4 *m4 = reinterpret_cast<int4 *>(matches);
        const int4 v = make_int4(RN_MATCH_NEG, RN_MATCH_NEG, RN_MATCH_NEG, RN_MATCH_NEG);
        for (size_t i = i0; i < n / 4; i += stride) m4[i] = v;
        for (size_t i = (n / 4) * 4 + i0; i < n; i += stride) matches[i] = RN_MATCH_NEG;
    } else {
        for (size_t i = i0; i < n; i += stride) matches[i] = RN_MATCH_NEG;
    }
}

// BYTES = false: writes the int32 `matches` array that rn_assign_fill_kernel has filled with RN_MATCH_NEG (rn_assign).
// BYTES = true (rn_loss_step): no fill kernel at all -- the output is a persistent BYTE map that is all-zero (= background)
// between launches: 1 + box index for a positive, 255 for an ignored anchor.  Every byte written is also recorded in the
// image's clean list, with which rn_loss_final_kernel zeroes the map again after the loss kernel has consumed it; the
// positive counts accumulate in a persistent zeroed counter array that the final kernel copies out and clears.
template <bool BYTES>
__global__ void __launch_bounds__(RN_SPARSE_THREADS)
rn_assign_sparse_kernel(const float4 *__restrict__ gt_boxes, const int64_t *__restrict__ gt_cats, int M,
                        const __grid_constant__ RnGeom g, float pos_thr, float neg_thr, int32_t *__restrict__ matches,
                        int32_t *__restrict__ npos, uint8_t *__restrict__ m8, int32_t *__restrict__ clean_list,
                        int32_t *__restrict__ clean_cnt, int parts) {
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *s_box = reinterpret_cast<float4 *>(smem);
    float *s_area = reinterpret_cast<float *>(s_box + M);
    __shared__ int s_info[2];  // valid boxes of the image, compacted index of this CTA's box
    __shared__ int s_cnt[RN_SPARSE_THREADS / 32];
    __shared__ int s_ix0[RN_NUM_LEVELS * RN_MAX_K], s_iy0[RN_NUM_LEVELS * RN_MAX_K], s_nw[RN_NUM_LEVELS * RN_MAX_K];
    __shared__ int s_pref[RN_NUM_LEVELS * RN_MAX_K + 1];
    __shared__ float s_ub[4][RN_NUM_LEVELS * RN_MAX_K];  // per window: bounding box of its anchors (x1, y1, x2, y2), rounded outwards
    __shared__ unsigned char s_list[128];                // compacted indices of the image's boxes that touch the windows (M <= 128)
    __shared__ int s_nl;
    // `parts` CTAs per box may share its candidates (interleaved in steps of 256); see rn_sparse_parts for why it stays 1
    const int b = blockIdx.y, row = blockIdx.x / parts, part = blockIdx.x - row * parts;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    rn_pdl_trigger();
    const int64_t *cats = gt_cats + (size_t)b * M;
    // every global load the prologue needs is issued before the first one is consumed (one L2 round trip instead of three):
    // this CTA's box, its category, and -- warp 0 -- the first 32 slots of the image for the compaction below
    const float4 me = gt_boxes[(size_t)b * M + row];
    const int64_t my_cat = cats[row];
    int64_t cat0 = -1;
    float4 box0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0 && lane < M) {
        cat0 = cats[lane];
        box0 = gt_boxes[(size_t)b * M + lane];
    }
    if (my_cat < 0) {  // padding row (Vision.py:1637-1638); uniform over the CTA
        // Every CTA of a PDL-launched grid must wait: the grid's completion is what the loss kernel behind it waits for,
        // and it only implies the completion of rn_assign_fill_kernel if no CTA retires without having waited for it.
        if (!BYTES) rn_pdl_wait();
        return;
    }
    if (warp == 0) {            // compact the image's boxes (as rn_compact_gt) and find this CTA's box among them
        const float4 *boxes = gt_boxes + (size_t)b * M;
        int cnt = 0, self = 0;
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = (j < M) && (j0 == 0 ? cat0 : cats[j]) >= 0;
            const unsigned mask = __ballot_sync(RN_FULL_MASK, valid);
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (valid) {
                const float4 bx = j0 == 0 ? box0 : boxes[j];
                s_box[pos] = bx;
                s_area[pos] = rn_area(bx);
                if (j == row) self = pos;
            }
            cnt += __popc(mask);
        }
        self = __reduce_max_sync(RN_FULL_MASK, self);
        if (lane == 0) {
            s_info[0] = cnt;
            s_info[1] = self;
        }
    }
    // ---- candidate windows: one per (level, base box), computed by warps 1.. while warp 0 compacts ----
    const double wg = (double)me.z - (double)me.x, hg = (double)me.w - (double)me.y;
    const int K = g.K, nseg = RN_NUM_LEVELS * K;
    for (int sg = tid - 32; sg >= 0 && sg < nseg; sg += RN_SPARSE_THREADS - 32) {
        const int l = sg / K, k = sg - l * K;
        int cx0 = 0, cy0 = 0, nw = 0, count = 0;
        float wb0 = INFINITY, wb1 = INFINITY, wb2 = -INFINITY, wb3 = -INFINITY;
        if (wg > 0.0 && hg > 0.0) {  // a degenerate box overlaps nothing
            const double Ag = wg * hg, cxg = 0.5 * ((double)me.x + (double)me.z), cyg = 0.5 * ((double)me.y + (double)me.w);
            const double tq = 0.95 * (double)neg_thr;
            const double *bb = g.base + (l * RN_MAX_K + k) * 4;
            const double wa = bb[2] - bb[0], ha = bb[3] - bb[1], Aa = wa * ha;
            if (fmin(Aa, Ag) >= tq * fmax(Aa, Ag)) {       // IoU <= min(A)/max(A)
                const double need = tq / (1.0 + tq) * (Aa + Ag);  // inter >= t/(1+t) * (Aa + Ag)
                const double dx = 0.5 * (wa + wg) - need / fmin(ha, hg);
                const double dy = 0.5 * (ha + hg) - need / fmin(wa, wg);
                if (dx >= 0.0 && dy >= 0.0) {
                    const double inv = 1.0 / (double)(8 << l);  // exact (power of two)
                    // cells whose centre (i + 0.5) * stride lies within the distance (+ 0.01 cell of slack)
                    // (clamped in float64 before the conversion: a box far outside the image must not overflow an int)
                    const double gwd = (double)g.gw[l], ghd = (double)g.gh[l];
                    const int ix0 = (int)fmin(gwd, fmax(0.0, ceil((cxg - dx) * inv - 0.51)));
                    const int ix1 = (int)fmin(gwd - 1.0, fmax(-1.0, floor((cxg + dx) * inv - 0.49)));
                    const int iy0 = (int)fmin(ghd, fmax(0.0, ceil((cyg - dy) * inv - 0.51)));
                    const int iy1 = (int)fmin(ghd - 1.0, fmax(-1.0, floor((cyg + dy) * inv - 0.49)));
                    if (ix1 >= ix0 && iy1 >= iy0) {
                        cx0 = ix0;
                        cy0 = iy0;
                        nw = ix1 - ix0 + 1;
                        count = nw * (iy1 - iy0 + 1);
                        const double st = (double)(8 << l);
                        wb0 = __double2float_rd(((double)ix0 + 0.5) * st + bb[0]);
                        wb1 = __double2float_rd(((double)iy0 + 0.5) * st + bb[1]);
                        wb2 = __double2float_ru(((double)ix1 + 0.5) * st + bb[2]);
                        wb3 = __double2float_ru(((double)iy1 + 0.5) * st + bb[3]);
                    }
                }
            }
        }
        s_ix0[sg] = cx0;
        s_iy0[sg] = cy0;
        s_nw[sg] = nw;
        s_pref[sg + 1] = count;
        s_ub[0][sg] = wb0;
        s_ub[1][sg] = wb1;
        s_ub[2][sg] = wb2;
        s_ub[3][sg] = wb3;
    }
    __syncthreads();
    const int m = s_info[0], self = s_info[1];
    // Cull (bounds the candidate loop by the boxes NEAR this one instead of all m: M = 100 valid boxes took 40 us): the
    // image's boxes that touch the bounding box of all candidate anchors, in ascending order.  Exact: a box outside it has an
    // intersection width or height <= 0, i.e. IoU exactly 0, with every candidate (the float32 anchor coordinates are roundings
    // of values inside the outward-rounded box).
    if (warp == 1) {
        float u0 = INFINITY, u1 = INFINITY, u2 = -INFINITY, u3 = -INFINITY;
        for (int sg = lane; sg < nseg; sg += 32) {
            u0 = fminf(u0, s_ub[0][sg]);
            u1 = fminf(u1, s_ub[1][sg]);
            u2 = fmaxf(u2, s_ub[2][sg]);
            u3 = fmaxf(u3, s_ub[3][sg]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u0 = fminf(u0, __shfl_xor_sync(RN_FULL_MASK, u0, o));
            u1 = fminf(u1, __shfl_xor_sync(RN_FULL_MASK, u1, o));
            u2 = fmaxf(u2, __shfl_xor_sync(RN_FULL_MASK, u2, o));
            u3 = fmaxf(u3, __shfl_xor_sync(RN_FULL_MASK, u3, o));
        }
        int nl = 0;
        for (int j0 = 0; j0 < m; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            if (j < m) {
                const float4 bx = s_box[j];
                hit = (j == self) || (bx.z > u0 && bx.x < u2 && bx.w > u1 && bx.y < u3);
            }
            const unsigned hm = __ballot_sync(RN_FULL_MASK, hit);
            if (hit) s_list[nl + __popc(hm & ((1u << lane) - 1u))] = (unsigned char)j;
            nl += __popc(hm);
        }
        if (lane == 0) s_nl = nl;
    }
    if (warp == 0) {  // inclusive scan of the window sizes (<= 80 entries)
        int carry = 0;
        for (int i0 = 0; i0 < nseg; i0 += 32) {
            const int i = i0 + lane;
            int v = (i < nseg) ? s_pref[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(RN_FULL_MASK, v, o);
                if (lane >= o) v += t;
            }
            if (i < nseg) s_pref[i + 1] = carry + v;
            carry += __shfl_sync(RN_FULL_MASK, v, 31);
        }
        if (lane == 0) s_pref[0] = 0;
    }
    __syncthreads();
    const int total = s_pref[nseg];
    const int nl = s_nl;
    int cnt = 0;
    if (!BYTES) rn_pdl_wait();  // launched with PDL behind rn_assign_fill_kernel: its background fill must be complete before we write
#pragma unroll 1
    for (int idx = part * RN_SPARSE_THREADS + tid; idx < total; idx += parts * RN_SPARSE_THREADS) {
        int seg = 0;  // largest seg with s_pref[seg] <= idx (binary search over <= 80 segments)
#pragma unroll
        for (int step = 64; step > 0; step >>= 1)
            if (seg + step < nseg && s_pref[seg + step] <= idx) seg += step;
        const int l = seg / K, k = seg - l * K;
        const int local = idx - s_pref[seg], nw = s_nw[seg];
        const int iy = s_iy0[seg] + local / nw, ix = s_ix0[seg] + local % nw;
        const double stride = (double)(8 << l);
        const double sx = __dmul_rn((double)ix + 0.5, stride);  // retinanet.py:458 (exact)
        const double sy = __dmul_rn((double)iy + 0.5, stride);  // retinanet.py:459
        const double *bb = g.base + (l * RN_MAX_K + k) * 4;
        float4 an;
        an.x = __double2float_rn(__dadd_rn(bb[0], sx));
        an.y = __double2float_rn(__dadd_rn(bb[1], sy));
        an.z = __double2float_rn(__dadd_rn(bb[2], sx));
        an.w = __double2float_rn(__dadd_rn(bb[3], sy));
        const float aa = rn_area(an);
        float best = 0.0f;  // IoU >= 0 and torch.max returns index 0 for an all-zero column
        int bi = 0;
        for (int q = 0; q < nl; ++q) {
            const int j = s_list[q];
            const float4 gb = s_box[j];
            const float iw = __fsub_rn(fminf(gb.z, an.z), fmaxf(gb.x, an.x));
            const float ih = __fsub_rn(fminf(gb.w, an.w), fmaxf(gb.y, an.y));
            if (iw > 0.0f && ih > 0.0f) {
                const float inter = __fmul_rn(iw, ih);
                const float uni = __fsub_rn(__fadd_rn(s_area[j], aa), inter);  // Vision.py:255
                const float v = __fdiv_rn(inter, uni);
                if (v > best) {  // strict: first maximal index wins (torch.max, Vision.py:1505)
                    best = v;
                    bi = j;
                }
            }
        }
        if (bi != self) continue;  // another box's CTA owns this anchor (or nothing overlaps it and self != 0)
        int mt = RN_MATCH_IGNORE;
        if (best > pos_thr) mt = bi;                 // Vision.py:1506, :1508-1509
        else if (best < neg_thr) continue;           // background: already there (Vision.py:1507)
        const int a = g.off[l] + (iy * g.gw[l] + ix) * K + k;
        if (BYTES) {
            m8[(size_t)b * g.A + a] = (uint8_t)(mt >= 0 ? 1 + mt : 255);
            clean_list[(size_t)b * g.A + atomicAdd(clean_cnt + b, 1)] = a;  // a few hundred per image; order is irrelevant
        } else {
            matches[(size_t)b * g.A + a] = mt;
        }
        cnt += (mt >= 0);
    }
    cnt = __reduce_add_sync(RN_FULL_MASK, cnt);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < RN_SPARSE_THREADS / 32; ++w) t += s_cnt[w];
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
// Target staging: the bounding-box half of AspectRatioCollater
// (Vision.py:770-785 scale + jitter, :798-809 padding with -1), on the device from one ragged upload.
// out[b, j] = float32((box * scale_b) * rand_scale + jitter) in float64 like NumPy, -1 beyond the image's count.
__global__ void rn_stage_targets_kernel(const double *__restrict__ boxes, const int64_t *__restrict__ cats,
                                        const int32_t *__restrict__ offsets, const double *__restrict__ scales,
                                        double rand_scale, double row_jit, double col_jit, int B, int M,
                                        float4 *__restrict__ out_boxes, int64_t *__restrict__ out_cats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * M) return;
    const int b = i / M, j = i - b * M;
    const int lo = offsets[b], n = offsets[b + 1] - lo;
    float4 ob = make_float4(-1.f, -1.f, -1.f, -1.f);
    long long oc = -1;
    if (j < n) {
        const double *src = boxes + 4 * (size_t)(lo + j);
        const double sc = scales[b];
        ob.x = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(src[0], sc), rand_scale), col_jit));
        ob.y = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(src[1], sc), rand_scale), row_jit));
        ob.z = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(src[2], sc), rand_scale), col_jit));
        ob.w = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(src[3], sc), rand_scale), row_jit));
        oc = cats[lo + j];
    }
    out_boxes[i] = ob;
    out_cats[i] = oc;
}

// ------------------------------------------------------------------------------------------------
// Image staging: the pixel half of AspectRatioCollater after its cv2.resize (Vision.py:775-777 jitter placement, :786 HWC -> CHW,
// :790-796 zero padding of every image to the batch's common 32-multiple size), from one ragged upload.
// out[b, c, y, x] = img_b[y - row_jit, x - col_jit, c] inside the image, 0 elsewhere.  Pure data movement: bit exact.
// One CTA per (output row y, image b): the source row (cols*C consecutive elements) is read with consecutive threads on
// consecutive addresses into a shared-memory row in pixel-major order, then every channel plane's row is written with
// 128-bit stores (Wp is a multiple of 32, so each plane row starts 128-byte aligned); the channel-stride shared-memory reads
// are conflict free for odd C.  First version: one strided 4-byte load per output element (HWC -> CHW uncoalesced).
// T = float: the reference's format.  T = uint8_t (extension, 4x fewer upload bytes): pixels are 0..255 and
// out = (float(x) / 255 - mean[c]) / std[c] in fp32, each operation rounded on its own.
struct RnStageNorm {
    float mean[8], std[8];
    int on;
};
template <typename T>
__global__ void __launch_bounds__(256)
rn_stage_images_kernel(const T *__restrict__ pixels, const int64_t *__restrict__ offsets, const int32_t *__restrict__ dims,
                       int C, int Hp, int Wp, int row_jit, int col_jit, const __grid_constant__ RnStageNorm norm,
                       float *__restrict__ out) {
    extern __shared__ __align__(16) float s_row[];  // [Wp][C]
    const int b = blockIdx.y, y = blockIdx.x, tid = threadIdx.x;
    const int rows = dims[2 * b], cols = dims[2 * b + 1];
    const int sy = y - row_jit;
    const bool row_in = sy >= 0 && sy < rows;
    const int n = C * Wp;
    const T *src = pixels + offsets[b] + (row_in ? (size_t)sy * cols * C : 0);
    const int lo = col_jit * C, hi = (col_jit + cols) * C;  // the image's columns in the padded row, in elements
    if (sizeof(T) == 1) {
        // bytes: zero the padding, then read the image's part of the row as aligned 32-bit words (4 pixels-channels per load;
        // a word that straddles the ends of the row only contributes its bytes inside it -- it lies inside the same 4-byte
        // aligned word of the allocation as a valid byte, so the read is always in bounds)
        for (int i = tid; i < n; i += 256)
            if (!row_in || i < lo || i >= hi) s_row[i] = 0.0f;
        if (row_in) {
            const int nb = hi - lo;
            const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 3);
            const uint32_t *w0 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(src) - mis);
            const int words = (mis + nb + 3) >> 2;
            for (int wi0 = tid; wi0 < words; wi0 += 4 * 256) {  // four independent word loads per thread in flight
                uint32_t w4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w4[u] = (wi0 + u * 256 < words) ? __ldg(w0 + wi0 + u * 256) : 0u;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int wi = wi0 + u * 256;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int idx = 4 * wi + k - mis;
                        if (wi < words && idx >= 0 && idx < nb) {
                            float v = (float)((w4[u] >> (8 * k)) & 0xffu);
                            if (norm.on) {
                                const int c = (lo + idx) % C;
                                v = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), norm.mean[c]), norm.std[c]);
                            }
                            s_row[lo + idx] = v;
                        }
                    }
                }
            }
        }
    } else {
        for (int i0 = tid; i0 < n; i0 += 8 * 256) {  // eight independent loads per thread in flight
            float v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 256;
                v8[u] = (row_in && i >= lo && i < hi) ? (float)__ldg(src + (i - lo)) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * 256 < n) s_row[i0 + u * 256] = v8[u];
        }
    }
    __syncthreads();
    const int W4 = Wp >> 2;
    for (int c = 0; c < C; ++c) {
        float4 *dst = reinterpret_cast<float4 *>(out + (((size_t)b * C + c) * Hp + y) * Wp);
        for (int x4 = tid; x4 < W4; x4 += 256) {
            const int x = 4 * x4;
            dst[x4] = make_float4(s_row[x * C + c], s_row[(x + 1) * C + c], s_row[(x + 2) * C + c], s_row[(x + 3) * C + c]);
        }
    }
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

// CTAs per ground-truth box of the sparse kernel (rn_set_option("assign_parts", n)); 1 by default.  Measured (whole step,
// graph replay): Pascal B=32 63.3 / 63.2 / 64.3 / 66.1 / 69.6 us and COCO B=16 341.4 / 340.6 / 342.2 / 344.6 / 348.1 us for
// 1 / 2 / 3 / 4 / 8 parts -- the kernel's time is its prologue (dependent loads, box compaction, float64 windows, two
// barriers), not the candidate loop, and every extra CTA repeats the prologue.
static int rn_sparse_parts(int M, int B) {
    (void)M; (void)B;
    const int p = rn_opt(RN_OPT_ASSIGN_PARTS);
    return p > 0 ? (p < 16 ? p : 16) : 1;
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
    const int ncell = anchors ? A : g.offc[RN_NUM_LEVELS];
    size_t smem = (size_t)M * (sizeof(float4) + sizeof(float)) + sizeof(int) * (size_t)B;  // boxes | areas | image weights
    if (smem > 200 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_assign: M=%d B=%d too large", M, B);
    cudaStream_t s = (cudaStream_t)stream;
    // Sparse path (see rn_assign_sparse_kernel): generated anchors, no max-IoU output, thresholds in the usual order.
    if (!anchors && !max_iou && M >= 1 && M <= 128 && neg_thr >= 0.2f && pos_thr >= neg_thr && !rn_opt(RN_OPT_ASSIGN_DENSE)) {
        const size_t n = (size_t)B * (size_t)A;
        const int fill_ctas = 148 * 4;
        if (B <= fill_ctas * 256) {
            rn_assign_fill_kernel<<<fill_ctas, 256, 0, s>>>(matches, n, npos, B);
            const size_t sm = (size_t)M * (sizeof(float4) + sizeof(float));
            const int parts = rn_sparse_parts(M, B);
            rn_launch_pdl(rn_assign_sparse_kernel<false>, dim3(M * parts, B), dim3(RN_SPARSE_THREADS), sm, s,
                          reinterpret_cast<const float4 *>(gt_boxes), gt_cats, M, g, pos_thr, neg_thr, matches, npos,
                          (uint8_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, parts);
            return rn_check_launch("rn_assign (sparse)");
        }
    }
    cudaError_t e = cudaMemsetAsync(npos, 0, sizeof(int32_t) * (size_t)B, s);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign memset: %s", cudaGetErrorString(e));
    const bool k9 = !anchors && K == 9;  // the reference's 3 ratios x 3 scales
    if (smem > 48 * 1024) {
        e = k9 ? cudaFuncSetAttribute(rn_assign_kernel<9, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
               : cudaFuncSetAttribute(rn_assign_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_assign smem: %s", cudaGetErrorString(e));
    }
    // One wave of persistent CTAs: five per SM in total (the resident limit).  Balanced mode (a 1-D grid whose CTAs
    // share themselves out over the images by work, see the kernel) for large images: COCO B=16 with 0..20 boxes per
    // image 26.8 -> 24.7 us; for Pascal-sized images the counting prologue costs more than the balance gains
    // (16.7 -> 18.9 us), so they keep the fixed share.  (Skipping the IEEE divide of weakly overlapping pairs changed
    // nothing: the kernel is bound by the dependent chain per candidate box, not by instruction count.)
    const int nthr = k9 ? RN_ASSIGN_CELLS * 3 : RN_ASSIGN_CELLS;
    const long long nwt = ((long long)ncell * (k9 ? 3 : 1) + 31) / 32;
    const long long max_ctas = (nwt + nthr / 32 - 1) / (nthr / 32);  // per image
    int w_base = 11, w_box = 1;  // measured (COCO shape: 11.7 us without boxes + 1.04 us per box and image)
    if (rn_opt(RN_OPT_ASSIGN_WBASE) > 0) w_base = rn_opt(RN_OPT_ASSIGN_WBASE);  // tuning override (rn_set_option)
    const bool no_balance = rn_opt(RN_OPT_ASSIGN_NO_BALANCE) != 0;
    const int total = RN_ASSIGN_CTAS * 148;
    dim3 grid;
    if (!no_balance && B > 1 && total >= 2 * B && (long long)total <= max_ctas * B && nwt >= 1024) {
        grid = dim3(total, 1);
    } else {
        int ctas = (total + B - 1) / B;
        if (ctas > max_ctas) ctas = (int)max_ctas;
        if (ctas < 1) ctas = 1;
        grid = dim3(ctas, B);
    }
    const float4 *gb4 = reinterpret_cast<const float4 *>(gt_boxes), *tb4 = reinterpret_cast<const float4 *>(anchors);
    if (k9) rn_assign_kernel<9, 3><<<grid, nthr, smem, s>>>(gb4, gt_cats, M, g, tb4, pos_thr, neg_thr, matches, npos, max_iou, B, w_base, w_box);
    else rn_assign_kernel<0, 1><<<grid, nthr, smem, s>>>(gb4, gt_cats, M, g, tb4, pos_thr, neg_thr, matches, npos, max_iou, B, w_base, w_box);
    return rn_check_launch("rn_assign");
}

#ifdef RN_EXPERIMENTAL
// The byte-map assignment of rn_loss_step (declared in rn_common.cuh): one launch, no fill.
int rn_assign_bytes(const float *gt_boxes, const int64_t *gt_cats, int B, int M, const RnGeom &g, float pos_thr, float neg_thr,
                    uint8_t *m8, int32_t *npos_acc, int32_t *clean_list, int32_t *clean_cnt, cudaStream_t s) {
    const size_t sm = (size_t)M * (sizeof(float4) + sizeof(float));
    const int parts = rn_sparse_parts(M, B);
    rn_assign_sparse_kernel<true><<<dim3(M * parts, B), RN_SPARSE_THREADS, sm, s>>>(
        reinterpret_cast<const float4 *>(gt_boxes), gt_cats, M, g, pos_thr, neg_thr, nullptr, npos_acc, m8, clean_list, clean_cnt,
        parts);
    return rn_check_launch("rn_assign (byte map)");
}
#endif

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

extern "C" int rn_stage_targets(const double *boxes, const int64_t *cats, const int32_t *offsets, const double *scales,
                                double rand_scale, int row_jit, int col_jit, int B, int M, float *out_boxes,
                                int64_t *out_cats, void *stream) {
    if (B <= 0 || M <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_targets: B=%d M=%d", B, M);
    if (!offsets || !scales || !out_boxes || !out_cats) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_targets: null pointer");
    if (((uintptr_t)out_boxes) & 15) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_targets: out_boxes must be 16-byte aligned");
    const int n = B * M;
    rn_stage_targets_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        boxes, cats, offsets, scales, rand_scale, (double)row_jit, (double)col_jit, B, M,
        reinterpret_cast<float4 *>(out_boxes), out_cats);
    return rn_check_launch("rn_stage_targets");
}

template <typename T>
static int rn_stage_images_impl(const T *pixels, const int64_t *offsets, const int32_t *dims, int B, int C, int Hp, int Wp,
                                int row_jit, int col_jit, const float *mean, const float *std, float *out, void *stream) {
    if (B <= 0 || C <= 0 || C > 8 || Hp <= 0 || Wp <= 0 || Hp > 65535 * 32 || B > 65535)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_images: B=%d C=%d Hp=%d Wp=%d", B, C, Hp, Wp);
    if (Wp % 4 || (((uintptr_t)out) & 15)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_images: Wp must be a multiple of 4 and out 16-byte aligned");
    if (!pixels || !offsets || !dims || !out) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_images: null pointer");
    if ((mean == nullptr) != (std == nullptr)) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_images: mean and std go together");
    const size_t smem = sizeof(float) * (size_t)C * (size_t)Wp;
    if (smem > 200 * 1024) return rn_set_error(RN_ERR_INVALID_ARG, "rn_stage_images: row of %d x %d floats does not fit shared memory", Wp, C);
    RnStageNorm norm;
    memset(&norm, 0, sizeof(norm));
    if (mean) {
        norm.on = 1;
        for (int c = 0; c < C; ++c) {
            norm.mean[c] = mean[c];
            norm.std[c] = std[c];
        }
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rn_stage_images_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_stage_images smem: %s", cudaGetErrorString(e));
    }
    rn_stage_images_kernel<T><<<dim3(Hp, B), 256, smem, (cudaStream_t)stream>>>(pixels, offsets, dims, C, Hp, Wp, row_jit, col_jit, norm, out);
    return rn_check_launch("rn_stage_images");
}

extern "C" int rn_stage_images(const float *pixels, const int64_t *offsets, const int32_t *dims, int B, int C, int Hp, int Wp,
                               int row_jit, int col_jit, float *out, void *stream) {
    return rn_stage_images_impl<float>(pixels, offsets, dims, B, C, Hp, Wp, row_jit, col_jit, nullptr, nullptr, out, stream);
}

extern "C" int rn_stage_images_u8(const unsigned char *pixels, const int64_t *offsets, const int32_t *dims, int B, int C, int Hp,
                                  int Wp, int row_jit, int col_jit, const float *mean, const float *std, float *out, void *stream) {
    return rn_stage_images_impl<unsigned char>(pixels, offsets, dims, B, C, Hp, Wp, row_jit, col_jit, mean, std, out, stream);
}
