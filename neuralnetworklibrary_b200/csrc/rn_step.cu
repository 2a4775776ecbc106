// rn_step.cu -- rn_loss_step, the training step of the path as one library call (anchor/GT assignment + focal/smooth-L1 loss
// forward AND backward + the final reduction): by default the kernels of rn_assign + rn_loss chained with programmatic
// dependent launch; with -DRN_EXPERIMENTAL also as ONE persistent kernel launch (the rest of this comment) or as a three-kernel
// byte-map chain.
//
// Replaces (reference file:line): match_anchors_objects Vision.py:1474-1511 (+ jaccard :234-256, the padding strip
// :1637-1638), ssd1 :1568-1605, focal_loss_retina :1513-1530, smoothL1_loss_retina :1532-1566, SSD_loss.__call__
// :1620-1644 and the autograd replay of all of it (General/Learner.py:514).
//
// Status of the persistent kernel: OPT-IN at build time (-DRN_EXPERIMENTAL) and at run time (rn_set_option("step_fused", 1));
// rn_loss_step launches the separate kernels by default, because they are still faster.  Measured on B200 (profiles/r02_summary.md; CUDA-graph replay, 2 rotating input sets):
//                                     separate kernels        this kernel
//     COCO  B=16  (2.17 GB)           0.352 ms  (0.94)        0.394 ms  (0.84)
//     Pascal B=32 (302 MB)            0.063 ms  (0.73)        0.082 ms  (0.56)
//     COCO  B=256 (34.7 GB)           5.36  ms  (0.99)        6.03  ms  (0.88)
// The per-CTA %globaltimer stamps (build with -DRN_EXPERIMENTAL -DRN_STEP_TIMING, profiles/step_timing.py) show where it goes for COCO B=16:
// phase A 8 us (median; 16 us for the last CTA), phase B 363 us with a 34 us spread between the first and the last CTA to run
// out of tickets, phase C 9 us.  So the three latency-bound pieces the single launch was meant to remove come back as phases
// of similar length, and the streaming phase is no faster than rn_loss_kernel's 340 us.
//
// Why one kernel was tried.  The separate-kernel chain (background fill, sparse assignment, streaming loss, final reduction)
// spends 17 us in its three small kernels -- pure latency -- which is 26 % of the Pascal-sized step (65.7 us) although the
// streaming kernel itself runs at 85 % of the HBM roofline (profiles/r01_launches.csv).  Design:
//   phase A  ground-truth boxes are handed out to WARPS by an atomic ticket, each box split into RN_STEP_PARTS tasks that
//            share its candidate anchors: the box's candidate windows per (level, base box), a cull of the image's other
//            boxes against the windows' bounding box, then every candidate anchor is evaluated exactly like in rn_assign
//            (float32(float64 base + shift), strict-fp32 IoU in ascending box order, first maximal index) and the owner
//            box writes ONE BYTE per non-background anchor into a per-image byte map that is all-zero (= background)
//            between launches.  No fill kernel, no [B,A] int32 array.  Before phase A every CTA draws its first chunk ticket
//            and issues L2 prefetches for that chunk of `clas`, so HBM already works while the assignment's dependent
//            latencies (~5 us) run.
//   phase B  the B*A anchor rows are cut into chunks of a few sub-tiles (2048 128-bit vectors each); the persistent CTAs
//            (3 per SM) draw chunk tickets in ascending order -- the memory window in flight moves through the tensors like
//            under the hardware CTA scheduler, and fast SMs simply draw more tickets.  The chunk size is chosen so that the
//            number of chunks is just under a whole number of rounds of the grid (no wave quantisation).  Per chunk: wait
//            (acquire) until the image's boxes are done, run the element math of rn_loss.cu on the chunk, then the owner
//            thread of each row computes its smooth-L1 term / d loss / d reg and re-zeroes the row's byte (the map is
//            self-cleaning: the kernel leaves it as it found it).  One partial pair per (chunk, image).
//   phase C  the last CTA to finish (one atomic per CTA) sums the partials per image in chunk order in float64, normalises
//            like the reference, writes {loss, reg_loss, clas_loss} and the positive counts, and resets the counters.
// Forward progress: phase B waits only on phase-A tasks, and a phase-A ticket is only ever held by a warp that is running
// and finishes it before its CTA waits for anything -- no CTA depends on a CTA that has not been scheduled yet.
// Determinism: the chunk grid is a pure function of the shapes and the device's SM count; a chunk's partial does not
// depend on which CTA computed it; all sums are fixed-order trees; no floating-point atomics.
// What the measurements taught (each item was worth 5-20 % of the step): (1) the streaming loop must be ONE code path -- a
// full and a ragged sub-tile variant of 16 KB each alternate out of the 32 KB instruction cache (stall_no_instruction 1.8
// per issue); (2) a function that is not inlined must receive the element-math scalars by value, a reference to the kernel
// parameters turns every use into a generic LD after each "memory"-clobbering store; (3) one ticket atomic per warp (3552 at
// once on one address) serialises longer than the work it hands out -- one per CTA and round; (4) 64-bit divisions in the
// final reduction cost more than its loads; (5) equal static slices per CTA finish 20 % apart (per-SM bandwidth is not
// uniform), tickets are needed for balance.
#include <string.h>

#include "rn_loss_math.cuh"

#define RN_STEP_CTAS 3                                // resident CTAs per SM (register bound, like rn_loss_kernel)
#define RN_STEP_WARPS (RN_THREADS / 32)
#define RN_STEP_PARTS 8                               // tasks per ground-truth box (its candidates are dealt round-robin)
#define RN_STEP_MAXSEG (RN_NUM_LEVELS * RN_MAX_K)     // candidate windows per box
#define RN_STEP_MAXM 128                              // ground-truth slots per image the fused step supports
#define RN_STEP_MAX_GRID 2048
#define RN_STEP_PF_BYTES (160 * 1024)                 // bytes of its first chunk a CTA prefetches into L2 before phase A

struct RnStepCtrl {
    int ticket_a;    // next phase-A task
    int ticket_b;    // next phase-B chunk
    int finished;    // CTAs that have left phase B
    int err;         // bit 0: a ground-truth category >= C was seen
    int tasks_done;  // phase-A tasks completed (all images)
    int pad[3];
};

struct RnStepWarpScratch {  // phase A, one per warp
    int ix0[RN_STEP_MAXSEG], iy0[RN_STEP_MAXSEG], nw[RN_STEP_MAXSEG], pref[RN_STEP_MAXSEG + 1];
    float4 cbox[RN_STEP_MAXM];
    float carea[RN_STEP_MAXM];
    unsigned char cidx[RN_STEP_MAXM];
};

struct RnStepParams {
    RnLossParams L;          // clas, reg, gt_boxes, gt_cats, dclas, dreg, probs, B, A, C, CV, M, scalars (matches/npos/partials unused)
    RnStepCtrl *ctrl;
    int *done;               // [B] phase-A tasks of the image that have completed
    int *npos_acc;           // [B] positives counted so far
    uint8_t *m8;             // [B][A] 0 background, 255 ignored, 1 + matched box
    double2 *partials;       // [nchunks][slots] {sum of focal terms, sum of smooth-L1 terms} per (chunk, image of the chunk)
    float *per_image;        // [B][2]
    float *out3;
    int32_t *npos_out;       // [B] or NULL
    int32_t *matches_out;    // [B][A] or NULL (inspection / tests)
    int total_rows;          // B*A
    int chunk_rows;          // rows per chunk
    int nchunks;
    int slots;               // partial slots per chunk = images a chunk can touch
    int pf_bytes;
    float pos_thr, neg_thr;
    float w_reg, w_clas, bs;
#ifdef RN_STEP_TIMING
    unsigned long long *dbg;  // [grid][8] %globaltimer stamps (profiles/step_timing.py)
#endif
};

#ifdef RN_EXPERIMENTAL  // the persistent kernel (device code below) and the byte-map chain: opt-in at build time
__device__ __forceinline__ int rn_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double rn_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RN_FULL_MASK, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Phase A: part `part` of RN_STEP_PARTS of the assignment contribution of ONE ground-truth box, by one warp (see
// rn_assign_sparse_kernel in rn_assign.cu for why the candidate windows are conservative; the arithmetic per candidate anchor
// is identical).  The parts of a box recompute its windows (cheap) and share out its candidates round-robin, so a huge box
// (a thousand candidates) does not become the one long task everything waits for.
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void rn_step_box_task(const RnStepParams &S, const RnGeom &g, int b, int row, int part,
                                                     RnStepWarpScratch &W) {
    const int lane = threadIdx.x & 31;
    const int M = S.L.M, K = g.K, nseg = RN_NUM_LEVELS * K;
    const int64_t *cats = S.L.gt_cats + (size_t)b * M;
    const float4 *boxes = S.L.gt_boxes + (size_t)b * M;
    const long long cat = cats[row];
    if (cat < 0) return;  // padding row (Vision.py:1637-1638); uniform over the warp
    if (cat >= S.L.C && lane == 0 && part == 0) atomicOr(&S.ctrl->err, 1);  // the reference raises IndexError (Vision.py:1593)
    const float4 me = boxes[row];
    const double wg = (double)me.z - (double)me.x, hg = (double)me.w - (double)me.y;
    const float neg_thr = S.neg_thr, pos_thr = S.pos_thr;
    // ---- candidate windows, one per (level, base box), and the bounding box of all their anchors ----
    double ux1 = INFINITY, uy1 = INFINITY, ux2 = -INFINITY, uy2 = -INFINITY;
    for (int sg = lane; sg < nseg; sg += 32) {
        const int l = sg / K, k = sg - l * K;
        int cx0 = 0, cy0 = 0, nw = 0, count = 0;
        if (wg > 0.0 && hg > 0.0) {  // a degenerate box overlaps nothing
            const double Ag = wg * hg, cxg = 0.5 * ((double)me.x + (double)me.z), cyg = 0.5 * ((double)me.y + (double)me.w);
            const double tq = 0.95 * (double)neg_thr;
            const double *bb = g.base + (l * RN_MAX_K + k) * 4;
            const double b0 = bb[0], b1 = bb[1], b2 = bb[2], b3 = bb[3];
            const double wa = b2 - b0, ha = b3 - b1, Aa = wa * ha;
            if (fmin(Aa, Ag) >= tq * fmax(Aa, Ag)) {              // IoU <= min(A)/max(A)
                const double need = tq / (1.0 + tq) * (Aa + Ag);  // inter >= t/(1+t) * (Aa + Ag)
                const double dx = 0.5 * (wa + wg) - need / fmin(ha, hg);
                const double dy = 0.5 * (ha + hg) - need / fmin(wa, wg);
                if (dx >= 0.0 && dy >= 0.0) {
                    const double stride = (double)(8 << l), inv = 1.0 / stride;  // exact (power of two)
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
                        ux1 = fmin(ux1, ((double)ix0 + 0.5) * stride + b0);
                        uy1 = fmin(uy1, ((double)iy0 + 0.5) * stride + b1);
                        ux2 = fmax(ux2, ((double)ix1 + 0.5) * stride + b2);
                        uy2 = fmax(uy2, ((double)iy1 + 0.5) * stride + b3);
                    }
                }
            }
        }
        W.ix0[sg] = cx0;
        W.iy0[sg] = cy0;
        W.nw[sg] = nw;
        W.pref[sg + 1] = count;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ux1 = fmin(ux1, __shfl_xor_sync(RN_FULL_MASK, ux1, o));
        uy1 = fmin(uy1, __shfl_xor_sync(RN_FULL_MASK, uy1, o));
        ux2 = fmax(ux2, __shfl_xor_sync(RN_FULL_MASK, ux2, o));
        uy2 = fmax(uy2, __shfl_xor_sync(RN_FULL_MASK, uy2, o));
    }
    __syncwarp();
    {   // inclusive scan of the window sizes (<= 80 entries)
        int carry = 0;
        for (int i0 = 0; i0 < nseg; i0 += 32) {
            const int i = i0 + lane;
            int v = (i < nseg) ? W.pref[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(RN_FULL_MASK, v, o);
                if (lane >= o) v += t;
            }
            if (i < nseg) W.pref[i + 1] = carry + v;
            carry += __shfl_sync(RN_FULL_MASK, v, 31);
        }
        if (lane == 0) W.pref[0] = 0;
    }
    __syncwarp();
    const int total = W.pref[nseg];
    if (total <= part * 32) return;  // nothing left for this part (parts take warp-sized groups of candidates round-robin)
    // ---- the image's boxes that touch the windows' bounding box, in ascending order (an exact cull: a box outside it has an
    // intersection width or height <= 0, i.e. IoU exactly 0, with every candidate anchor; the float32 anchor coordinates
    // are the roundings of values inside the float64 box, so rounding the box outwards keeps them inside) ----
    const float bx1 = __double2float_rd(ux1), by1 = __double2float_rd(uy1), bx2 = __double2float_ru(ux2), by2 = __double2float_ru(uy2);
    int nl = 0, nvalid = 0, self = 0;
    for (int j0 = 0; j0 < M; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = (j < M) && cats[j] >= 0;
        const unsigned vmask = __ballot_sync(RN_FULL_MASK, valid);
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        bool hit = false;
        if (valid) {
            bx = boxes[j];
            hit = (j == row) || (bx.z > bx1 && bx.x < bx2 && bx.w > by1 && bx.y < by2);
        }
        const unsigned hmask = __ballot_sync(RN_FULL_MASK, hit);
        const unsigned lt = (1u << lane) - 1u;
        if (hit) {
            const int pos = nl + __popc(hmask & lt);
            W.cbox[pos] = bx;
            W.carea[pos] = rn_area(bx);
            W.cidx[pos] = (unsigned char)(nvalid + __popc(vmask & lt));
        }
        if (row >= j0 && row < j0 + 32) self = nvalid + __popc(vmask & ((1u << (row - j0)) - 1u));
        nl += __popc(hmask);
        nvalid += __popc(vmask);
    }
    __syncwarp();
    // ---- candidates: groups of 32 consecutive candidates, group q belongs to part q % RN_STEP_PARTS ----
    uint8_t *m8 = S.m8 + (size_t)b * g.A;
    int cnt = 0;
#pragma unroll 1
    for (int idx = part * 32 + lane; idx < total; idx += 32 * RN_STEP_PARTS) {
        int seg = 0;  // largest seg with pref[seg] <= idx
#pragma unroll
        for (int step = 64; step > 0; step >>= 1)
            if (seg + step < nseg && W.pref[seg + step] <= idx) seg += step;
        const int l = seg / K, k = seg - l * K;
        const int local = idx - W.pref[seg], nw = W.nw[seg];
        const int iy = W.iy0[seg] + local / nw, ix = W.ix0[seg] + local % nw;
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
            const float4 gb = W.cbox[q];
            const float iw = __fsub_rn(fminf(gb.z, an.z), fmaxf(gb.x, an.x));
            const float ih = __fsub_rn(fminf(gb.w, an.w), fmaxf(gb.y, an.y));
            if (iw > 0.0f && ih > 0.0f) {
                const float inter = __fmul_rn(iw, ih);
                const float uni = __fsub_rn(__fadd_rn(W.carea[q], aa), inter);  // Vision.py:255
                const float v = __fdiv_rn(inter, uni);
                if (v > best) {  // strict: first maximal index wins (torch.max, Vision.py:1505)
                    best = v;
                    bi = W.cidx[q];
                }
            }
        }
        if (bi != self) continue;  // another box's task owns this anchor (or nothing overlaps it)
        unsigned code = 255u;                          // ignored: max IoU in [neg_thr, pos_thr]
        if (best > pos_thr) code = 1u + (unsigned)bi;  // Vision.py:1506, :1508-1509
        else if (best < neg_thr) continue;             // background: already there (Vision.py:1507)
        m8[g.off[l] + (iy * g.gw[l] + ix) * K + k] = (uint8_t)code;
        cnt += (code != 255u);
    }
    cnt = __reduce_add_sync(RN_FULL_MASK, cnt);
    if (lane == 0 && cnt) atomicAdd(S.npos_acc + b, cnt);  // integer: order independent
}

// ------------------------------------------------------------------------------------------------
// Phase B, hot part: the class vectors [v0, v1) of image b (whole rows).  Not inlined on purpose: the streaming loop then
// has the register file to itself (80 registers at 3 CTAs per SM, like rn_loss_kernel) instead of sharing it with the loop
// control of the persistent kernel around it.  Returns this thread's share of the focal sum.
// ------------------------------------------------------------------------------------------------
struct RnStepElem {  // the scalars of the element math, passed BY VALUE (registers): in a function that is not inlined a
    float lo, hi, a_neg, a_pos, gamma;  // reference to the kernel parameters is a generic pointer, and the "memory" clobber
    float *probs;                       // of every streaming store would make the compiler reload each field (~8 LD per vector)
};
template <int V, int CVT, bool G2, bool GRAD, bool LOGITS>
static __device__ __noinline__ float rn_step_chunk(RnStepElem E, int b, const float *__restrict__ x_img, float *__restrict__ dx_img,
                                                   const uint8_t *__restrict__ m_img, const int *s_cat, int CVr, int nvec,
                                                   int v0, int v1, float gl) {
    RnLossParams L;
    L.lo = E.lo; L.hi = E.hi; L.a_neg = E.a_neg; L.a_pos = E.a_pos; L.gamma = E.gamma; L.probs = E.probs;
    const int CV = CVT ? CVT : CVr;
    float acc_neg = 0.0f, acc_pos = 0.0f;
    // ONE code path for full and ragged sub-tiles (the bounds-checked one, whose loads are unconditional): a chunk of whole
    // rows ends in a ragged sub-tile, and alternating between two 16 KB loop bodies overflows the 32 KB instruction cache
    // (measured: stall_no_instruction 1.8 per issue and 0.47 instead of 0.35 ms for COCO B=16, profiles/r02_summary.md).
#pragma unroll 1
    for (int tile0 = v0; tile0 < v1; tile0 += RN_LOSS_TILE)
        rn_loss_subtile<V, CVT, G2, GRAD, false, LOGITS, RnMatchU8>(L, b, x_img, dx_img, m_img, s_cat, CV, nvec, v1, tile0, gl, acc_neg, acc_pos);
    return acc_neg + acc_pos;
}

// Phase B, per-row part of a chunk: smooth L1 of the positive rows (Vision.py:1532-1566), d loss / d reg for every row, the
// optional dense matches, and the self-cleaning of the byte map.  Returns this thread's share of the smooth-L1 sum.
template <bool GRAD>
static __device__ __forceinline__ float rn_step_rows(const RnStepParams &S, const RnGeom &g, int b, int c0, int c1, float ge,
                                                     const float4 *s_box) {
    const int A = S.L.A;
    uint8_t *m_img = S.m8 + (size_t)b * A;
    const float4 *reg4 = reinterpret_cast<const float4 *>(S.L.reg) + (size_t)b * A;
    float4 *dreg4 = GRAD ? reinterpret_cast<float4 *>(S.L.dreg) + (size_t)b * A : nullptr;
    int32_t *mout = S.matches_out ? S.matches_out + (size_t)b * A : nullptr;
    float acc_reg = 0.0f;
    for (int row = c0 + (int)threadIdx.x; row < c1; row += RN_THREADS) {
        unsigned code;
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(code) : "l"(m_img + row));
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (code != 0u) {
            m_img[row] = 0;  // back to "background" for the next launch
            if (code != 255u)
                g4 = rn_smooth_l1_row(rn_anchor_from_param(g, nullptr, row), s_box[code - 1u], __ldg(reg4 + row), ge, acc_reg);
        }
        if (GRAD) dreg4[row] = g4;
        if (mout) mout[row] = code == 0u ? RN_MATCH_NEG : (code == 255u ? RN_MATCH_IGNORE : (int)code - 1);
    }
    return acc_reg;
}

// ------------------------------------------------------------------------------------------------
// Phase C: executed by the last CTA only.
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void rn_step_final(const RnStepParams &S, float2 *s_img /* smem [<= cap] or NULL */) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = S.L.B, A = S.L.A, CR = S.chunk_rows, SL = S.slots;
    const bool err = *reinterpret_cast<volatile int *>(&S.ctrl->err) != 0;
    for (int b = warp; b < B; b += RN_THREADS / 32) {
        // chunks that hold rows of image b; only the first of them can have started in an earlier image (its slot for b is
        // the number of image boundaries before b inside it), every later one starts inside b (slot 0)
        const int r0 = b * A, r1 = r0 + A - 1;
        const int n_lo = r0 / CR, n_hi = r1 / CR;
        const int slot_lo = b - (n_lo * CR) / A;
        double cs = 0.0, rs = 0.0;
#pragma unroll 1
        for (int n0 = n_lo + lane; n0 <= n_hi; n0 += 32 * 8) {  // 8 independent loads per lane in flight, fixed summation order
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int n = n0 + 32 * u;
                v[u] = make_double2(0.0, 0.0);
                if (n <= n_hi) v[u] = __ldcg(S.partials + (size_t)n * SL + (n == n_lo ? slot_lo : 0));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                cs += v[u].x;
                rs += v[u].y;
            }
        }
        cs = rn_warp_sum_d(cs);
        rs = rn_warp_sum_d(rs);
        if (lane == 0) {
            const int n = __ldcg(S.npos_acc + b);
            const float n_norm = fmaxf((float)n, 1.0f);
            float2 pi;
            pi.x = n > 0 ? __fdiv_rn((float)rs, (float)(4 * n)) : 0.0f;  // reg loss of image b (mean over npos*4, Vision.py:1566)
            pi.y = __fdiv_rn((float)cs, n_norm);                          // clas loss of image b (Vision.py:1530)
            if (s_img) s_img[b] = pi;
            else reinterpret_cast<float2 *>(S.per_image)[b] = pi;
            if (S.npos_out) S.npos_out[b] = n;
            S.npos_acc[b] = 0;  // leave the state buffer as it was found
            S.done[b] = 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float reg_total = 0.f, clas_total = 0.f;
        for (int b = 0; b < B; ++b) {  // image order, fp32, like the reference's Python loop (Vision.py:1640-1641)
            const float2 pi = s_img ? s_img[b] : reinterpret_cast<float2 *>(S.per_image)[b];
            reg_total = __fadd_rn(reg_total, pi.x);
            clas_total = __fadd_rn(clas_total, pi.y);
        }
        const float reg_loss = __fdiv_rn(reg_total, S.bs), clas_loss = __fdiv_rn(clas_total, S.bs);
        float loss = __fadd_rn(__fmul_rn(S.w_reg, reg_loss), __fmul_rn(S.w_clas, clas_loss));
        if (err) loss = __int_as_float(0x7fc00000);  // a category >= C: the reference raises; here the loss is poisoned
        S.out3[0] = loss;
        S.out3[1] = reg_loss;
        S.out3[2] = clas_loss;
        S.ctrl->ticket_a = 0;
        S.ctrl->ticket_b = 0;
        S.ctrl->finished = 0;
        S.ctrl->err = 0;
        S.ctrl->tasks_done = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Per-image context of phase B (double buffered: the context of the next chunk's image is loaded while the current chunk
// finishes).
struct RnStepCtx {
    float4 box[RN_STEP_MAXM];  // compacted ground-truth boxes
    int cat[RN_STEP_MAXM];     // their categories
    int npos;                  // positives of the image
    int image;                 // which image this is (-1: none)
};

// Loads the context of image b into `ctx`: warp 0 compacts the ground truth, lane 0 of warp 1 fetches the positive count.
// Callers separate this from the readers with a CTA barrier.
__device__ __forceinline__ void rn_step_load_ctx(const RnStepParams &S, int b, RnStepCtx &ctx) {
    const int tid = threadIdx.x;
    const int M = S.L.M;
    if (tid < 32) rn_compact_gt(S.L.gt_boxes + (size_t)b * M, S.L.gt_cats + (size_t)b * M, M, ctx.box, nullptr, ctx.cat);
    if (tid == 32) {
        ctx.npos = *reinterpret_cast<volatile const int *>(S.npos_acc + b);
        ctx.image = b;
    }
}

template <int V, int CVT, bool G2, bool GRAD, bool LOGITS>
__global__ void __launch_bounds__(RN_THREADS, RN_STEP_CTAS)
rn_step_kernel(const __grid_constant__ RnStepParams S, const __grid_constant__ RnGeom g) {
    // phase A scratch and the phase B / C tables share this buffer (their lifetimes do not overlap)
    __shared__ __align__(16) unsigned char s_raw[sizeof(RnStepWarpScratch) * RN_STEP_WARPS];
    __shared__ double s_red[2][2][RN_STEP_WARPS];  // [parity of the chunk][focal, smooth-L1][warp]
    __shared__ int s_next, s_last, s_alldone;
    static_assert(sizeof(RnStepWarpScratch) * RN_STEP_WARPS >= 2 * sizeof(RnStepCtx), "phase B tables must fit");

    const RnLossParams &P = S.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = P.A, C = P.C, M = P.M;
    const int CV = CVT ? CVT : P.CV;
    const int nvec = A * CV;
    const int CR = S.chunk_rows;
    const int ntask = P.B * M * RN_STEP_PARTS;
#ifdef RN_STEP_TIMING
    auto stamp = [&](int i) {
        if (S.dbg && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            S.dbg[(size_t)blockIdx.x * 8 + i] = t;
        }
    };
#else
    auto stamp = [](int) {};
#endif
    stamp(0);

    // The first chunk ticket, and an L2 prefetch of that chunk of clas ([B*A, C] is contiguous over images): HBM works
    // while phase A runs.
    if (tid == 0) {
        s_next = atomicAdd(&S.ctrl->ticket_b, 1);
        s_alldone = 0;
    }
    __syncthreads();
    int n = s_next;
    if (n < S.nchunks) {
        const int lo = n * CR, hi = min(S.total_rows, lo + CR);
        const char *head = reinterpret_cast<const char *>(P.clas) + (size_t)lo * C * sizeof(float);
        size_t bytes = (size_t)(hi - lo) * C * sizeof(float);
        if (bytes > (size_t)S.pf_bytes) bytes = (size_t)S.pf_bytes;
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)RN_THREADS * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(head + o));
    }

    // ---- phase A ----
    // The CTA draws RN_STEP_WARPS consecutive tasks at a time (one atomic per CTA and round instead of one per warp: 3552 warps
    // hammering one counter took longer than the tasks themselves); with RN_STEP_PARTS == RN_STEP_WARPS that is one box.
    {
        RnStepWarpScratch &W = reinterpret_cast<RnStepWarpScratch *>(s_raw)[warp];
        for (;;) {
            if (tid == 0) s_last = atomicAdd(&S.ctrl->ticket_a, RN_STEP_WARPS);
            __syncthreads();
            const int t0 = s_last;
            if (t0 >= ntask) break;
            const int t = t0 + warp;
            if (t < ntask) {
                const int part = t % RN_STEP_PARTS, box = t / RN_STEP_PARTS;
                const int b = box / M;
                rn_step_box_task(S, g, b, box - b * M, part, W);
            }
            __syncthreads();  // the round's tasks have issued all their stores (and s_last may be overwritten)
            if (tid == 0) {
                __threadfence();  // release: the bytes and the positive counts of this round before the counters
                const int nt = min(RN_STEP_WARPS, ntask - t0);
                int first = t0;
                while (first < t0 + nt) {  // the round's tasks, grouped by image (usually one group)
                    const int b = (first / RN_STEP_PARTS) / M;
                    const int end = min(t0 + nt, (b + 1) * M * RN_STEP_PARTS);
                    atomicAdd(S.done + b, end - first);
                    first = end;
                }
                atomicAdd(&S.ctrl->tasks_done, nt);
            }
        }
    }
    stamp(1);

    // ---- phase B ----
    RnStepCtx *ctx = reinterpret_cast<RnStepCtx *>(s_raw);  // [2]
    RnStepElem E;
    E.lo = P.lo; E.hi = P.hi; E.a_neg = P.a_neg; E.a_pos = P.a_pos; E.gamma = P.gamma; E.probs = P.probs;
    __syncthreads();  // phase A scratch is dead
    if (tid == 32) {
        ctx[0].image = -1;
        ctx[1].image = -1;
    }
    int cur = 0;      // context buffer of the image being processed
    int parity = 0;   // of the chunk, for s_red

    // Blocks until the phase-A tasks of image b are complete and visible (thread 0 polls, the barrier publishes).  Once all
    // tasks of the launch are known to be complete no image is polled any more (`alldone` is the same in every thread: it is
    // only ever read from shared memory right after the barrier that follows its one write).
    bool alldone = false;
    auto wait_image = [&](int b) {
        if (!alldone) {
            if (tid == 0) {
                if (rn_ld_acquire(&S.ctrl->tasks_done) >= ntask) {
                    s_alldone = 1;
                } else {
                    while (rn_ld_acquire(S.done + b) < M * RN_STEP_PARTS) __nanosleep(32);
                }
            }
            __syncthreads();
            alldone = s_alldone != 0;
        }
    };

    if (n < S.nchunks) {  // context of the first chunk's first image
        const int b0 = (n * CR) / A;
        wait_image(b0);
        rn_step_load_ctx(S, b0, ctx[cur]);
        __syncthreads();
    }
#pragma unroll 1
    while (n < S.nchunks) {
        const int lo = n * CR, hi = min(S.total_rows, lo + CR);
        if (tid == 64) s_next = atomicAdd(&S.ctrl->ticket_b, 1);  // the ticket after the next one is already known: see below
        double cs_t = 0.0, rs_t = 0.0;  // this thread's sums over the chunk's segments, per slot handled below
        int slot = 0;
#pragma unroll 1
        for (int r = lo; r < hi; ++slot) {
            const int b = r / A;
            const int s0 = r - b * A;
            const int s1 = min(A, hi - b * A);
            if (ctx[cur].image != b) {  // only the later images of a chunk that crosses an image boundary (rare): blocking reload
                __syncthreads();
                wait_image(b);
                rn_step_load_ctx(S, b, ctx[cur]);
                __syncthreads();
            }
            const int n_pos = ctx[cur].npos;
            const float n_norm = fmaxf((float)n_pos, 1.0f);      // clamp(min=1), Vision.py:1530
            const float gl = __fdiv_rn(P.wc_over_bs, n_norm);    // upstream of every focal term
            const float ge = n_pos > 0 ? __fdiv_rn(P.wr_over_bs, (float)(4 * n_pos)) : 0.0f;  // mean() backward
            const uint8_t *m_img = S.m8 + (size_t)b * A;
            const float *x_img = P.clas + (size_t)b * A * C;
            float *dx_img = GRAD ? P.dclas + (size_t)b * A * C : nullptr;
            const float fc = rn_step_chunk<V, CVT, G2, GRAD, LOGITS>(E, b, x_img, dx_img, m_img, ctx[cur].cat, CV, nvec, s0 * CV, s1 * CV, gl);
            __syncthreads();  // every thread has consumed its rows' bytes: the owners may now clean them
            const float fr = rn_step_rows<GRAD>(S, g, b, s0, s1, ge, ctx[cur].box);
            // ---- warp sums (fixed order) of this segment ----
            const double cw = rn_warp_sum_d((double)fc), rw = rn_warp_sum_d((double)fr);
            if (hi - b * A > A) {  // another segment follows: finish this one now (rare)
                if (lane == 0) {
                    s_red[parity][0][warp] = cw;
                    s_red[parity][1][warp] = rw;
                }
                __syncthreads();
                if (tid == 0) {
                    double cs = 0.0, rs = 0.0;
#pragma unroll
                    for (int w = 0; w < RN_STEP_WARPS; ++w) {
                        cs += s_red[parity][0][w];
                        rs += s_red[parity][1][w];
                    }
                    S.partials[(size_t)n * S.slots + slot] = make_double2(cs, rs);
                }
                __syncthreads();
            } else {
                cs_t = cw;
                rs_t = rw;
            }
            r = b * A + s1;
        }
        // ---- end of the chunk: publish the warp sums, and load the context of the NEXT chunk's image into the other buffer;
        // one barrier covers both.  s_next was written at the start of this chunk (a whole chunk ago) ----
        if (lane == 0) {
            s_red[parity][0][warp] = cs_t;
            s_red[parity][1][warp] = rs_t;
        }
        const int nn = s_next;
        int next_cur = cur;
        if (nn < S.nchunks) {
            const int bn = (nn * CR) / A;
            if (ctx[cur].image != bn) {
                next_cur = cur ^ 1;
                wait_image(bn);
                rn_step_load_ctx(S, bn, ctx[next_cur]);
            }
        }
        __syncthreads();
        if (tid == 0) {  // the last segment's partial (fixed order over the warps); the others already stream the next chunk
            double cs = 0.0, rs = 0.0;
#pragma unroll
            for (int w = 0; w < RN_STEP_WARPS; ++w) {
                cs += s_red[parity][0][w];
                rs += s_red[parity][1][w];
            }
            S.partials[(size_t)n * S.slots + (slot - 1)] = make_double2(cs, rs);
        }
        parity ^= 1;
        cur = next_cur;
        n = nn;
    }
    stamp(2);

    // ---- phase C: the last CTA reduces ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();  // release this CTA's partials
        s_last = (atomicAdd(&S.ctrl->finished, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();  // acquire the others'
    rn_step_final(S, P.B * sizeof(float2) <= sizeof(s_raw) ? reinterpret_cast<float2 *>(s_raw) : nullptr);
    stamp(3);
}

#endif  // RN_EXPERIMENTAL

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline size_t rn_up256(size_t x) { return (x + 255) / 256 * 256; }

#ifdef RN_EXPERIMENTAL
// Chunk grid for `grid` persistent CTAs: k sub-tiles per chunk, k in 2..6 chosen so that the number of chunks is just under
// a whole number of rounds of the grid (ties: the larger chunk); rows per chunk aligned so that every chunk of clas starts on
// a 128-byte boundary.  Pure function of (B*A, C, grid).
struct RnStepPlan {
    int chunk_rows, nchunks, slots;
};
static RnStepPlan rn_step_plan(long long total_rows, int A, int C, int grid) {
    const int V = (C % 4 == 0) ? 4 : 1, CV = C / V;
    int align = 32;  // rows * C * 4 bytes must be a multiple of 128
    for (int a = 1; a <= 32; a <<= 1)
        if (((long long)a * C * 4) % 128 == 0) { align = a; break; }
    RnStepPlan cand[7];
    double eff[7], best_eff = -1.0;
    for (int k = 2; k <= 6; ++k) {
        long long rows = ((long long)k * RN_LOSS_TILE) / CV / align * align;
        if (rows < align) rows = align;
        const long long nch = (total_rows + rows - 1) / rows;
        eff[k] = ((double)nch / grid) / (double)((nch + grid - 1) / grid);  // busy fraction of the last round included
        cand[k].chunk_rows = (int)rows;
        cand[k].nchunks = (int)nch;
        if (eff[k] > best_eff) best_eff = eff[k];
    }
    RnStepPlan best = cand[2];
    for (int k = 2; k <= 6; ++k)
        if (eff[k] >= best_eff - 0.01) best = cand[k];  // within 1 %: the larger chunk (less per-chunk overhead)
    best.slots = (best.chunk_rows + A - 1) / A + 1;
    return best;
}

#endif

// Two caller-owned buffers.  `state` (ctrl | done | npos_acc | byte map) must be all-zero when a call starts and is all-zero
// again when it ends -- whatever the shapes, so one grow-only zero-initialised buffer serves calls of any shape.  `workspace`
// is plain scratch.
struct RnStepState {
    size_t ctrl, done, npos_acc, clean_cnt, m8, total;
};
static RnStepState rn_step_state_layout(int B, int A) {
    RnStepState w;
    size_t o = 0;
    w.ctrl = o;      o += rn_up256(sizeof(RnStepCtrl));
    w.done = o;      o += rn_up256(sizeof(int) * (size_t)B);
    w.npos_acc = o;  o += rn_up256(sizeof(int) * (size_t)B);
    w.clean_cnt = o; o += rn_up256(sizeof(int) * (size_t)B);
    w.m8 = o;        o += rn_up256((size_t)B * (size_t)A);
    w.total = o;
    return w;
}
struct RnStepWs {
    size_t partials, per_image, matches32, npos32, loss_ws, total;  // matches32 doubles as the clean list of the byte-map chain
};
static size_t rn_step_max_partials(int B, int A, int C) {
    // the smallest chunk (2 sub-tiles) gives the most chunks; slots is largest for the largest chunk (6 sub-tiles)
    const int V = (C % 4 == 0) ? 4 : 1, CV = C / V;
    long long rows_min = (2LL * RN_LOSS_TILE) / CV / 32 * 32;
    if (rows_min < 1) rows_min = 1;
    const long long rows_max = (6LL * RN_LOSS_TILE) / CV + 32;
    const long long nch = ((long long)B * A + rows_min - 1) / rows_min + 1;
    const long long slots = (rows_max + A - 1) / A + 1;
    return (size_t)(nch * slots);
}
static RnStepWs rn_step_layout(int B, int A, int C) {
    RnStepWs w;
    size_t o = 0;
    w.partials = o;  o += rn_up256(sizeof(double2) * rn_step_max_partials(B, A, C));
    w.per_image = o; o += rn_up256(sizeof(float) * 2 * (size_t)B);
    // the separate-kernel path (the default; see the measurements at the top of this file)
    w.matches32 = o; o += rn_up256(sizeof(int32_t) * (size_t)B * (size_t)A);
    w.npos32 = o;    o += rn_up256(sizeof(int32_t) * (size_t)B);
    w.loss_ws = o;   o += rn_up256(rn_loss_workspace_bytes(B, A, C));
    w.total = o;
    return w;
}

extern "C" size_t rn_loss_step_workspace_bytes(int B, int A, int C) {
    if (B <= 0 || A <= 0 || C <= 0) return 256;
    return rn_step_layout(B, A, C).total;
}

extern "C" size_t rn_loss_step_state_bytes(int B, int A) {
    if (B <= 0 || A <= 0) return 256;
    return rn_step_state_layout(B, A).total;
}

extern "C" int rn_loss_step_state_init(void *state, size_t state_bytes, void *stream) {
    if (!state || (((uintptr_t)state) & 255)) return rn_set_error(RN_ERR_WORKSPACE, "rn_loss_step_state_init: null or misaligned state buffer");
    cudaError_t e = cudaMemsetAsync(state, 0, state_bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "rn_loss_step_state_init: %s", cudaGetErrorString(e));
    return RN_OK;
}

#ifdef RN_EXPERIMENTAL
#ifdef RN_STEP_TIMING
static unsigned long long *g_step_dbg = nullptr;
extern "C" void rn_step_debug_buffer(void *p) { g_step_dbg = reinterpret_cast<unsigned long long *>(p); }
#endif

// resident CTAs of a kernel on the current device (queried once per kernel and device)
template <typename Kern>
static int rn_step_max_grid(Kern kernel, int slot) {
    static int cache[64][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || slot < 0 || slot >= 64) return 148 * RN_STEP_CTAS;
    if (cache[dev][slot] == 0) {
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RN_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 1;
        int n = per_sm * sms;
        cache[dev][slot] = n > RN_STEP_MAX_GRID ? RN_STEP_MAX_GRID : n;
    }
    return cache[dev][slot];
}

template <int V, int CVT, bool G2, bool GRAD, bool LOGITS>
static int rn_step_launch(RnStepParams &S, const RnGeom &g, cudaStream_t s) {
    auto kernel = rn_step_kernel<V, CVT, G2, GRAD, LOGITS>;
    const int slot = ((V == 4 ? 0 : 1) * 3 + (CVT == 20 ? 0 : (CVT == 5 ? 1 : 2))) * 8 + (G2 ? 4 : 0) + (GRAD ? 2 : 0) + (LOGITS ? 1 : 0);
    const int max_grid = rn_step_max_grid(kernel, slot);
    const RnStepPlan plan = rn_step_plan(S.total_rows, S.L.A, S.L.C, max_grid);
    S.chunk_rows = plan.chunk_rows;
    S.nchunks = plan.nchunks;
    S.slots = plan.slots;
    const int grid = plan.nchunks < max_grid ? plan.nchunks : max_grid;
    kernel<<<grid, RN_THREADS, 0, s>>>(S, g);
    return rn_check_launch("rn_loss_step");
}

template <int V, int CVT, bool LOGITS>
static int rn_step_dispatch(bool g2, bool grad, RnStepParams &S, const RnGeom &g, cudaStream_t s) {
    if (g2 && grad) return rn_step_launch<V, CVT, true, true, LOGITS>(S, g, s);
    // the rarer variants (forward only, gamma != 2) share the generic row width
    if constexpr (CVT != 0) {
        return rn_step_dispatch<V, 0, LOGITS>(g2, grad, S, g, s);
    } else {
        if (g2) return rn_step_launch<V, 0, true, false, LOGITS>(S, g, s);
        if (grad) return rn_step_launch<V, 0, false, true, LOGITS>(S, g, s);
        return rn_step_launch<V, 0, false, false, LOGITS>(S, g, s);
    }
}

#endif  // RN_EXPERIMENTAL

extern "C" int rn_loss_step(const float *clas, const float *reg, const float *gt_boxes, const int64_t *gt_cats, int B, int A,
                            int C, int M, int H, int W, const double *base, int K, const float *anchors, float pos_thr,
                            float neg_thr, double alpha, double gamma, double beta, int B_global, int from_logits,
                            float *dclas, float *dreg, float *probs_out, float *out3, int32_t *npos_out,
                            int32_t *matches_out, void *state, size_t state_bytes, void *workspace, size_t workspace_bytes,
                            void *stream) {
    if (B < 0 || A <= 0 || C <= 0 || M < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: B=%d A=%d C=%d M=%d", B, A, C, M);
    if (B == 0) {  // an empty image shard (more ranks than images): this rank's share of the three scalars is zero
        if (!out3) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: null pointer");
        cudaError_t e = cudaMemsetAsync(out3, 0, 3 * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? RN_OK : rn_set_error(RN_ERR_CUDA, "rn_loss_step memset: %s", cudaGetErrorString(e));
    }
    if (!clas || !reg || !out3 || (M > 0 && (!gt_boxes || !gt_cats))) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: null pointer");
    if ((dclas == nullptr) != (dreg == nullptr))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: dclas and dreg must both be given or both be NULL");
    if (B_global < B) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: B_global=%d < B=%d", B_global, B);
    const RnStepWs L = rn_step_layout(B, A, C);
    const RnStepState Z = rn_step_state_layout(B, A);
    if (!workspace || workspace_bytes < L.total || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_loss_step: workspace needs %zu bytes, 256-byte aligned", L.total);
    if (!state || state_bytes < Z.total || (((uintptr_t)state) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_loss_step: state needs %zu bytes, 256-byte aligned and zero-initialised", Z.total);
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
#ifdef RN_EXPERIMENTAL
    unsigned char *zs = reinterpret_cast<unsigned char *>(state);
    cudaStream_t s = (cudaStream_t)stream;
#endif

#ifdef RN_EXPERIMENTAL
    const bool fused = !anchors && M >= 1 && M <= RN_STEP_MAXM && neg_thr >= 0.2f && pos_thr >= neg_thr &&
                       (long long)B * A <= 0x7fffffffLL - 65536 && rn_opt(RN_OPT_STEP_FUSED) != 0;
    const bool bytemap = !fused && !anchors && M >= 1 && M < RN_STEP_MAXM && neg_thr >= 0.2f && pos_thr >= neg_thr &&
                         rn_opt(RN_OPT_STEP_BYTEMAP) != 0;
    if (bytemap) {
        // Opt-in: THREE kernels and no [B,A] int32 array.  One CTA per ground-truth box writes a byte per non-background
        // anchor into the persistent zeroed map (+ its index into a clean list), the streaming loss kernel reads the bytes,
        // and the final reduction zeroes them again.  Against the four-kernel chain of rn_assign + rn_loss this removes the
        // background fill (4 us, 4*A*B bytes written and read back) -- but the fill already overlaps the assignment's
        // prologue through PDL, the byte loads make the streaming kernel 2 % slower and the cleaning lengthens the
        // single-CTA final kernel: measured COCO B=16 0.359 vs 0.353 ms, Pascal B=32 75 vs 65 us (profiles/r02_summary.md).
        RnGeom g;
        int rc = rn_build_geom(&g, H, W, base, K, nullptr, A);
        if (rc) return rc;
        RnLossBytes by;
        by.m8 = zs + Z.m8;
        by.clean_list = reinterpret_cast<const int32_t *>(ws + L.matches32);
        by.clean_cnt = reinterpret_cast<int32_t *>(zs + Z.clean_cnt);
        by.npos_acc = reinterpret_cast<int32_t *>(zs + Z.npos_acc);
        by.npos_out = npos_out;
        rc = rn_assign_bytes(gt_boxes, gt_cats, B, M, g, pos_thr, neg_thr, by.m8, by.npos_acc,
                             reinterpret_cast<int32_t *>(ws + L.matches32), by.clean_cnt, s);
        if (rc) return rc;
        return rn_loss_impl(from_logits != 0, from_logits ? probs_out : nullptr, clas, reg, gt_boxes, gt_cats, matches_out,
                            by.npos_acc, &by, B, A, C, M, H, W, base, K, nullptr, alpha, gamma, beta, B_global, dclas, dreg, out3,
                            ws + L.loss_ws, rn_loss_workspace_bytes(B, A, C), stream);
    }
#else
    const bool fused = false;
#endif
    if (!fused) {  // the separate kernels: rn_assign (dense or sparse) + rn_loss
        int32_t *m32 = matches_out ? matches_out : reinterpret_cast<int32_t *>(ws + L.matches32);
        int32_t *n32 = npos_out ? npos_out : reinterpret_cast<int32_t *>(ws + L.npos32);
        int rc = rn_assign(gt_boxes, gt_cats, B, M, H, W, base, K, anchors, A, pos_thr, neg_thr, m32, n32, nullptr, stream);
        if (rc) return rc;
        const size_t lw = rn_loss_workspace_bytes(B, A, C);
        if (from_logits)
            return rn_loss_logits(clas, reg, gt_boxes, gt_cats, m32, n32, B, A, C, M, H, W, base, K, anchors, alpha, gamma, beta,
                                  B_global, dclas, dreg, probs_out, out3, ws + L.loss_ws, lw, stream);
        return rn_loss(clas, reg, gt_boxes, gt_cats, m32, n32, B, A, C, M, H, W, base, K, anchors, alpha, gamma, beta, B_global,
                       dclas, dreg, out3, ws + L.loss_ws, lw, stream);
    }

#ifdef RN_EXPERIMENTAL
    const int V = (C % 4 == 0) ? 4 : 1;
    if ((long long)A * (C / V) > 0x7fffffffLL - RN_LOSS_TILE) return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: A*C too large");
    if (V == 4 && ((((uintptr_t)clas) | ((uintptr_t)dclas) | ((uintptr_t)probs_out)) & 15))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: clas/dclas/probs_out must be 16-byte aligned");
    if ((((uintptr_t)reg) | ((uintptr_t)dreg) | ((uintptr_t)gt_boxes)) & 15)
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: reg/dreg/gt_boxes must be 16-byte aligned");
    RnGeom g;
    int rc = rn_build_geom(&g, H, W, base, K, nullptr, A);
    if (rc) return rc;

    RnStepParams S;
    memset(&S, 0, sizeof(S));
    RnLossParams &P = S.L;
    P.clas = clas; P.reg = reg;
    P.gt_boxes = reinterpret_cast<const float4 *>(gt_boxes); P.gt_cats = gt_cats;
    P.dclas = dclas; P.dreg = dreg; P.probs = from_logits ? probs_out : nullptr;
    P.B = B; P.A = A; P.C = C; P.CV = C / V; P.M = M;
    P.a_pos = (float)alpha; P.a_neg = (float)(1.0 - alpha);  // Vision.py:1526
    P.gamma = (float)gamma;
    P.lo = (float)1e-4; P.hi = (float)(1.0 - 1e-4);          // Vision.py:1524
    const float bs = (float)B_global;
    S.w_reg = (float)(1.0 - beta); S.w_clas = (float)beta; S.bs = bs;  // Vision.py:1644
    P.wc_over_bs = S.w_clas / bs;
    P.wr_over_bs = S.w_reg / bs;
    S.ctrl = reinterpret_cast<RnStepCtrl *>(zs + Z.ctrl);
    S.done = reinterpret_cast<int *>(zs + Z.done);
    S.npos_acc = reinterpret_cast<int *>(zs + Z.npos_acc);
    S.m8 = zs + Z.m8;
    S.partials = reinterpret_cast<double2 *>(ws + L.partials);
    S.per_image = reinterpret_cast<float *>(ws + L.per_image);
    S.out3 = out3; S.npos_out = npos_out; S.matches_out = matches_out;
    S.total_rows = B * A;
    S.pf_bytes = RN_STEP_PF_BYTES;
    S.pos_thr = pos_thr; S.neg_thr = neg_thr;
#ifdef RN_STEP_TIMING
    S.dbg = g_step_dbg;
#endif

    const bool g2 = (gamma == 2.0), grad = dclas != nullptr;
    if (from_logits) {
        if (V == 4 && C == 80) return rn_step_dispatch<4, 20, true>(g2, grad, S, g, s);
        if (V == 4 && C == 20) return rn_step_dispatch<4, 5, true>(g2, grad, S, g, s);
        if (V == 4) return rn_step_dispatch<4, 0, true>(g2, grad, S, g, s);
        return rn_step_dispatch<1, 0, true>(g2, grad, S, g, s);
    }
    if (V == 4 && C == 80) return rn_step_dispatch<4, 20, false>(g2, grad, S, g, s);
    if (V == 4 && C == 20) return rn_step_dispatch<4, 5, false>(g2, grad, S, g, s);
    if (V == 4) return rn_step_dispatch<4, 0, false>(g2, grad, S, g, s);
    return rn_step_dispatch<1, 0, false>(g2, grad, S, g, s);
#else
    return rn_set_error(RN_ERR_INVALID_ARG, "rn_loss_step: unreachable");
#endif
}
