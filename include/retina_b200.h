/*
 * retina_b200.h -- C ABI of libretina_sm100.so: the B200-native (sm_100a) implementation of the
 * RetinaNet loss / post-processing hot path of NickTravers/NeuralNetworkLibrary.
 *
 * The reference has no FFI layer: its boundary is Python duck typing at four call sites (SURVEY.md
 * section 8b).  The entry points below are what a binding for that boundary calls; each one names
 * the reference code it replaces (file:line under the reference checkout).  INTEGRATION.md shows the
 * ctypes stubs; neuralnetworklibrary_b200/_lib.py is the binding used by the drop-in wrappers.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless marked "host";
 *   - the caller owns every buffer (the wrappers let torch's caching allocator provide them);
 *   - no hidden allocation, no device synchronisation, no default-stream use: every launch goes to
 *     the `stream` argument (a cudaStream_t passed as void*), so calls are CUDA-graph capturable and
 *     re-entrant per stream as long as each stream has its own workspace;
 *   - return value: RN_OK or an RN_ERR_* code; rn_last_error() gives the message (thread local);
 *   - there is no CPU fallback.  Without a CUDA device every compute entry returns RN_ERR_CUDA.
 *
 * Anchors.  Anchor a lives at level l (P3..P7), grid cell (iy, ix) and slot k of the K = nr*ns base
 * boxes, a = off_l + (iy*gw_l + ix)*K + k (reference retinanet.py:467-469, :491-495).  Kernels take
 * EITHER a device table `anchors` [A,4] fp32 (any anchor set, e.g. the reference's own tensor) OR
 * anchors == NULL plus the image size (H, W) and the host table `base` [5][K][4] float64 =
 * size_l * get_anchor_set() (retinanet.py:439-451, :480, :492) and generate the anchor on the fly as
 * float32(base + shift) -- the reference's float64 arithmetic followed by TEN()'s rounding
 * (General/Core.py:61-62), bit for bit.
 *
 * matches encoding: >= 0 index of the matched ground-truth box (among the image's non-padding
 * rows, in order), RN_MATCH_NEG background, RN_MATCH_IGNORE max IoU in [neg_thr, pos_thr].
 */
#ifndef RETINA_B200_H
#define RETINA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_OK 0
#define RN_ERR_INVALID_ARG 1 /* bad shape / null pointer / unsupported parameter value */
#define RN_ERR_WORKSPACE 2   /* workspace too small or misaligned */
#define RN_ERR_CUDA 3        /* CUDA runtime error (message carries cudaGetErrorString) */

#define RN_MATCH_NEG (-1)
#define RN_MATCH_IGNORE (-2)

#define RN_NUM_LEVELS 5 /* P3..P7, retinanet.py:478 */
#define RN_MAX_K 16     /* max base boxes per cell (the reference uses 9) */
#define RN_MAX_TOP_K 4096

/* Message of the last error on the calling thread ("" if none). */
const char *rn_last_error(void);
/* ABI version (bumped on any signature change). */
int rn_abi_version(void);

/* Process-wide tuning / test switches (all default 0; nothing is ever read from the environment).  Names:
 *   "assign_dense" (!= 0: rn_assign always takes the dense kernel), "assign_no_balance", "assign_wbase" (dense kernel work
 *   balancing), "loss_iters" (> 0: sub-tiles per CTA of rn_loss), "levels_nchunks" (> 0: class chunks per row tile of
 *   rn_loss_levels), "step_fused" (!= 0: rn_loss_step runs as one persistent kernel where its conditions hold),
 *   "step_bytemap" (!= 0: rn_loss_step takes its three-kernel byte-map chain instead of rn_assign + rn_loss),
 *   "loss_prefetch" (L2 prefetch of rn_loss before its dependency wait: 1 first sub-tile / no look-ahead in rn_loss_levels,
 *   2 all of the CTA's, 3 + next wave), "assign_parts" (> 0: CTAs per ground-truth box of the sparse assignment kernel).
 * "step_fused" / "step_bytemap" can only be switched on in a library built with -DRN_EXPERIMENTAL (both measured slower than
 * the default and are left out of the default build); rn_get_option("experimental") tells (1 / 0, read-only).
 * rn_set_option returns RN_ERR_INVALID_ARG for an unknown name; rn_get_option returns -1 for one. */
int rn_set_option(const char *name, int value);
int rn_get_option(const char *name);

/* A = K * sum_l ceil(H/2^l)*ceil(W/2^l), l = 3..7 (retinanet.py:488).  Host-only arithmetic. */
int rn_num_anchors(int H, int W, int K);

/* AnchorGenerator.__call__ (retinanet.py:485-495): writes the [A,4] fp32 anchor table. */
int rn_anchors(int H, int W, const double *base /*host [5][K][4]*/, int K, float *anchors_out /*[A,4]*/,
               void *stream);

/* jaccard + match_anchors_objects for a whole batch (Vision.py:234-256, :1474-1511, and the padding
 * strip of Vision.py:1637-1638: a ground-truth row is padding iff gt_cats < 0).
 *   gt_boxes [B,M,4] fp32, gt_cats [B,M] int64 -> matches [B,A] int32, npos [B] int32 (#positives).
 * max_iou ([B,A] fp32) may be NULL.
 * Two implementations with identical results: with generated anchors, max_iou == NULL, M <= 128 and
 * 0.2 <= neg_thr <= pos_thr a background fill plus one CTA per ground-truth box over that box's candidate anchors
 * (sparse); otherwise a dense walk over all anchors.  rn_set_option("assign_dense", 1) forces the dense one (used by
 * the tests that compare the two). */
int rn_assign(const float *gt_boxes, const int64_t *gt_cats, int B, int M, int H, int W,
              const double *base /*host*/, int K, const float *anchors /*[A,4] or NULL*/, int A,
              float pos_thr, float neg_thr, int32_t *matches, int32_t *npos, float *max_iou, void *stream);

/* ComputeMaxOverlaps (Vision.py:1666-1694): per ground-truth row the max IoU with any anchor.
 *   out [B,M] fp32; padding rows receive -1. */
int rn_max_overlaps(const float *gt_boxes, const int64_t *gt_cats, int B, int M, int H, int W,
                    const double *base /*host*/, int K, const float *anchors /*or NULL*/, int A,
                    float *out, void *stream);

/* The bounding-box half of AspectRatioCollater (Vision.py:770-785 scale + jitter, :798-809 -1 padding) from one
 * ragged upload: boxes [N,4] float64 and cats [N] int64 are the images' boxes
 * concatenated, offsets [B+1] int32 the image boundaries, scales [B] float64 the per-image resize factors (all
 * DEVICE).  out_boxes [B,M,4] fp32 = float32((box*scale_b)*rand_scale + jitter), out_cats [B,M] int64, rows
 * beyond an image's count are -1.  M = max(1, longest image). */
int rn_stage_targets(const double *boxes, const int64_t *cats, const int32_t *offsets, const double *scales,
                     double rand_scale, int row_jit, int col_jit, int B, int M, float *out_boxes,
                     int64_t *out_cats, void *stream);

/* The pixel half of AspectRatioCollater after its cv2.resize (Vision.py:775-777 jitter placement, :786 HWC -> CHW,
 * :790-796 zero padding to the batch's common size): pixels holds the images' float32 HWC data back to back, offsets [B]
 * int64 the first element of each image, dims [B][2] int32 its (rows, cols) -- all DEVICE, one ragged upload.
 * out [B, C, Hp, Wp] fp32: out[b, c, y, x] = img_b[y - row_jit, x - col_jit, c] inside the image, 0 elsewhere. */
int rn_stage_images(const float *pixels, const int64_t *offsets, const int32_t *dims, int B, int C, int Hp, int Wp,
                    int row_jit, int col_jit, float *out, void *stream);

/* Extension of rn_stage_images for loaders that keep 8-bit images (4x fewer upload bytes): pixels are uint8 HWC, and with
 * mean / std (host [C], both or neither) out = (float(x) / 255 - mean[c]) / std[c] in fp32, each operation rounded on its
 * own -- the normalisation the reference's transforms apply on the host (Vision.py section 3) moved behind the upload;
 * without them out = float(x). */
int rn_stage_images_u8(const unsigned char *pixels, const int64_t *offsets, const int32_t *dims, int B, int C, int Hp, int Wp,
                       int row_jit, int col_jit, const float *mean /*host or NULL*/, const float *std /*host or NULL*/,
                       float *out, void *stream);

/* Workspace for rn_loss (bytes; 256-byte aligned base required). */
size_t rn_loss_workspace_bytes(int B, int A, int C);

/* ssd1 + focal_loss_retina + smoothL1_loss_retina + SSD_loss.__call__ forward AND backward in one
 * streaming pass (Vision.py:1513-1644), given the assignment from rn_assign.
 *   clas [B,A,C] fp32 probabilities (post-sigmoid, as the reference's heads return them), reg [B,A,4].
 *   out3 [3] fp32 = {loss, reg_loss, clas_loss}: the value SSD_loss returns and the two attributes it
 *     stores (Vision.py:1643-1644).  B_global >= B is the batch size the sums are divided by (an image
 *     shard passes the global batch; the three outputs are then this shard's additive share).
 *   dclas [B,A,C], dreg [B,A,4]: d loss / d clas and d loss / d reg for upstream gradient 1, fully
 *     written (zeros where the reference's gradient is zero).  Pass both NULL for a forward-only call
 *     (the reference's `evaluate` under no_grad). */
int rn_loss(const float *clas, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
            const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
            const double *base /*host*/, int K, const float *anchors /*or NULL*/, double alpha,
            double gamma, double beta, int B_global, float *dclas, float *dreg, float *out3,
            void *workspace, size_t workspace_bytes, void *stream);

/* The whole training step in ONE launch: rn_assign + rn_loss (or rn_loss_logits when from_logits != 0) + the final
 * reduction, i.e. SSD_loss.__call__ with its assignment (Vision.py:1474-1511, :1568-1644) and the autograd replay
 * (General/Learner.py:514).  Same inputs, outputs and numerics as the two separate calls; pos_thr / neg_thr are the
 * thresholds of match_anchors_objects (0.5 / 0.4).  npos_out [B] (positives per image) and matches_out [B,A] (the
 * rn_assign encoding; costs 4*A*B bytes of stores, meant for inspection and tests) may be NULL.  B == 0 (an empty image
 * shard) only zeroes out3.
 * By default the call launches the kernels of rn_assign + rn_loss (background fill, one CTA per ground-truth box, the
 * streaming loss kernel, the final reduction; chained with programmatic dependent launch) -- the fastest variant measured
 * (profiles/r02_summary.md).  Two alternatives exist behind options, for generated anchors (anchors == NULL), M < 128 and
 * 0.2 <= neg_thr <= pos_thr: rn_set_option("step_bytemap", 1): three kernels around a persistent all-zero BYTE map kept in
 * `state` (no fill, no [B,A] int32 array; the final reduction zeroes the written bytes again);
 * rn_set_option("step_fused", 1): one persistent kernel (rn_step.cu).  All three give identical gradients.
 * Two caller-owned buffers, both 256-byte aligned, one pair per stream: `workspace` (rn_loss_step_workspace_bytes) is
 * plain scratch; `state` (rn_loss_step_state_bytes) must be ZERO-INITIALISED once before its first use
 * (rn_loss_step_state_init, or any memset) and every call leaves it all-zero again, whatever the shapes -- so one
 * grow-only zeroed buffer serves calls of any shape. */
size_t rn_loss_step_workspace_bytes(int B, int A, int C);
size_t rn_loss_step_state_bytes(int B, int A);
int rn_loss_step_state_init(void *state, size_t state_bytes, void *stream);
int rn_loss_step(const float *clas, const float *reg, const float *gt_boxes, const int64_t *gt_cats, int B, int A,
                 int C, int M, int H, int W, const double *base /*host*/, int K, const float *anchors /*or NULL*/,
                 float pos_thr, float neg_thr, double alpha, double gamma, double beta, int B_global, int from_logits,
                 float *dclas, float *dreg, float *probs_out, float *out3, int32_t *npos_out, int32_t *matches_out,
                 void *state, size_t state_bytes, void *workspace, size_t workspace_bytes, void *stream);

/* Same as rn_loss, but the class activations are LOGITS: the head's nn.Sigmoid (retinanet.py:258, :286) is
 * fused into the kernel (y = 1/(1+exp(-z)) within ~5 ulp of the correctly rounded value) and
 * dlogits = d loss / d logits (chained through sigmoid's backward, grad*(1-y)*y).  SURVEY.md section 8f row 1:
 * not a drop-in (ObjectDetectionNet.forward must return logits); it removes one full read+write pass over
 * [B,A,C] from the model's forward and one from its backward.  probs_out ([B,A,C], may be NULL) receives
 * sigmoid(logits) exactly as the kernel used it (for inference-time reuse and for parity checks). */
int rn_loss_logits(const float *logits, const float *reg, const float *gt_boxes, const int64_t *gt_cats,
                   const int32_t *matches, const int32_t *npos, int B, int A, int C, int M, int H, int W,
                   const double *base /*host*/, int K, const float *anchors /*or NULL*/, double alpha,
                   double gamma, double beta, int B_global, float *dlogits, float *dreg, float *probs_out,
                   float *out3, void *workspace, size_t workspace_bytes, void *stream);

/* Workspace for rn_loss_levels (bytes; 256-byte aligned base required). */
size_t rn_loss_levels_workspace_bytes(int B, int H, int W, int K, int C);

/* rn_loss / rn_loss_logits on the heads' NCHW level tensors as the convolutions produce them, i.e. WITHOUT the
 * permute(0,2,3,1).contiguous().view() of retinanet.py:215-217, :289-295 and the torch.cat of Vision.py:1467-1468
 * (removes two more read+write passes over [B,A,C] from the model's forward and two from
 * its backward).  clas_levels / reg_levels / dclas_levels / dreg_levels / probs_levels are HOST arrays of
 * RN_NUM_LEVELS device pointers (P3..P7): clas_l [B, K*C, gh_l, gw_l] fp32 with channel = k*C + c, reg_l
 * [B, K*4, gh_l, gw_l] with channel = k*4 + j, gh_l = ceil(H/2^l), gw_l = ceil(W/2^l); gradients and probs_out have
 * the layout of their inputs.  from_logits != 0: the class tensors hold logits (sigmoid fused, see rn_loss_logits).
 * matches / npos come from rn_assign with anchors == NULL (generated anchors; anchor a = off_l + (iy*gw_l+ix)*K + k).
 * Same value, gradients and out3 as rn_loss on the permuted + concatenated tensors (summation order differs). */
int rn_loss_levels(const float *const *clas_levels /*host[5]*/, const float *const *reg_levels /*host[5]*/,
                   int from_logits, const float *gt_boxes, const int64_t *gt_cats, const int32_t *matches,
                   const int32_t *npos, int B, int C, int M, int H, int W, const double *base /*host*/, int K,
                   double alpha, double gamma, double beta, int B_global, float *const *dclas_levels /*host[5] or NULL*/,
                   float *const *dreg_levels /*host[5] or NULL*/, float *const *probs_levels /*host[5] or NULL*/,
                   float *out3, void *workspace, size_t workspace_bytes, void *stream);

/* The one exchange of the image-sharded multi-GPU path: out3 (this rank's additive share of
 * {loss, reg_loss, clas_loss}) is replaced by the sum over all ranks, computed in rank order (bit-identical on every rank)
 * by ONE tiny kernel on `stream` -- capturable in the step's CUDA graph -- that stores the three scalars into every peer's
 * buffer through peer-mapped memory (NVLink), publishes them with a release store of a sequence number and waits for the
 * peers' with acquire loads.  peer_bufs: HOST array of `world` device pointers, peer_bufs[r] = rank r's buffer of
 * rn_peer_exchange_bytes(world) bytes as mapped into THIS process (e.g. torch symmetric memory / CUDA IPC), zeroed once
 * before the first use; seq: one zero-initialised uint32 in local device memory.  world <= 16.  Every rank must call it
 * the same number of times (the kernel waits for its peers). */
size_t rn_peer_exchange_bytes(int world);
int rn_peer_exchange(float *out3, void *const *peer_bufs /*host [world]*/, int rank, int world, uint32_t *seq, void *stream);

/* The same with separate input and output: total3 receives the sums, in3 (this rank's share) is left untouched -- repeating
 * the call for the same step is then harmless.  Used by the pipelined form of the captured step (SSD_loss.capture(...,
 * pipelined_exchange=True)): the exchange of step k runs on a parallel branch at the START of step k+1's CUDA graph, beside the
 * assignment and streaming kernels, so nothing on the step's critical path waits for NVLink or for a peer. */
int rn_peer_exchange_to(const float *in3, float *total3, void *const *peer_bufs /*host [world]*/, int rank, int world,
                        uint32_t *seq, void *stream);

/* Backward with a non-unit upstream gradient: scales dclas[n_clas], dreg[n_reg] in place by the
 * DEVICE scalar *grad_out; the kernel exits immediately when *grad_out == 1 (what loss.backward()
 * passes, General/Learner.py:514), so the common case costs one empty launch and no host sync. */
int rn_scale_grads(float *dclas, size_t n_clas, float *dreg, size_t n_reg, const float *grad_out,
                   void *stream);

/* Workspace for rn_postproc (bytes). */
size_t rn_postproc_workspace_bytes(int B, int A, int top_k);

/* BBoxPredictor.__call__ + nms with rel_thresh/inc/dup = None (retinanet.py:732-812, :523-711):
 * class max + threshold + decode + clip + empty-box filter, top_k by score, class-aware greedy NMS
 * (bitmask form), first max_keep survivors.
 *   mean, std: host [4].  max_keep <= top_k <= RN_MAX_TOP_K.
 *   outputs [B,max_keep,...] score-descending; counts [B]; anchor_idx [B,max_keep] (the keep indices
 *   the reference only has implicitly) and n_candidates [B] (#boxes handed to nms) may be NULL.
 *   Ties in score are ordered by ascending anchor index (the reference's sort is unstable). */
int rn_postproc(const float *clas, const float *reg, int B, int A, int C, int H, int W,
                const double *base /*host*/, int K, const float *anchors /*or NULL*/,
                const float *mean /*host*/, const float *std /*host*/, float thresh, float max_overlap,
                int top_k, int max_keep, float *boxes, int64_t *classes, float *scores,
                int32_t *anchor_idx, int32_t *counts, int32_t *n_candidates, void *workspace,
                size_t workspace_bytes, void *stream);

/* rn_postproc on the heads' NCHW level tensors (see rn_loss_levels for the layout): clas_levels / reg_levels are HOST
 * arrays of RN_NUM_LEVELS device pointers; from_logits != 0: the class tensors hold logits, the score is
 * sigmoid(max logit) computed like torch's CUDA sigmoid (expf + IEEE divide) and class ties are resolved on the
 * probabilities, so the result equals torch.sigmoid + permute/view/cat + rn_postproc bit for bit while the three passes
 * over [B,A,C] those ops cost disappear.  Generated anchors only; needs A <= 2^24 and C <= 256.  Workspace:
 * rn_postproc_workspace_bytes(B, rn_num_anchors(H, W, K), top_k). */
int rn_postproc_levels(const float *const *clas_levels /*host[5]*/, const float *const *reg_levels /*host[5]*/,
                       int from_logits, int B, int C, int H, int W, const double *base /*host*/, int K,
                       const float *mean /*host*/, const float *std /*host*/, float thresh, float max_overlap,
                       int top_k, int max_keep, float *boxes, int64_t *classes, float *scores, int32_t *anchor_idx,
                       int32_t *counts, int32_t *n_candidates, void *workspace, size_t workspace_bytes, void *stream);

/* Workspace for rn_nms (bytes). */
size_t rn_nms_workspace_bytes(int n, int top_k);

/* nms() on caller-provided boxes (retinanet.py:523-711 with rel_thresh/inc/dup = None; the TTA call
 * site Vision.py:2118).  boxes [n,4] fp32, classes [n] int64, scores [n] fp32.
 *   keep_idx [max_keep] int32: indices into the inputs, score-descending; count [1] int32. */
int rn_nms(const float *boxes, const int64_t *classes, const float *scores, int n, float max_overlap,
           int top_k, int max_keep, int32_t *keep_idx, int32_t *count, void *workspace,
           size_t workspace_bytes, void *stream);

/* nms() for L images in ONE launch (the merge step of ImageLearner.TTA_bbox, Vision.py:2104-2119; also what the drop-in
 * nms() uses with L = 1): boxes [n_total,4] fp32, classes [n_total] int64, scores [n_total] fp32 are the images' candidates
 * concatenated, offsets [L+1] int32 (DEVICE) the image boundaries.  Outputs per image, score-descending: out_boxes
 * [L,max_keep,4], out_classes [L,max_keep] int64, out_scores [L,max_keep], out_idx [L,max_keep] int32 (index inside the
 * image's own candidates; may be NULL), counts [L] -- one buffer, one device->host copy for the caller.
 * Optional fused pre-pass (S > 0): the un-transform of Vision.py:2091-2097 per SEGMENT of consecutive boxes (seg_off [S+1]
 * int32, seg_par [S][5] float64 = col_jit, row_jit, 1/(rand_scale*scale), flip, cols; DEVICE): x -= col_jit, y -= row_jit,
 * times the factor, then a horizontal flip about cols if flip != 0 -- float64 arithmetic, rounded to float32 once. */
size_t rn_nms_batch_workspace_bytes(int n_total, int L, int top_k);
int rn_nms_batch(const float *boxes, const int64_t *classes, const float *scores, const int32_t *offsets, int L, int n_total,
                 const int32_t *seg_off, const double *seg_par, int S, float max_overlap, int top_k, int max_keep,
                 float *out_boxes, int64_t *out_classes, float *out_scores, int32_t *out_idx, int32_t *counts,
                 void *workspace, size_t workspace_bytes, void *stream);

/* The per-image matching of mAP1 (Vision.py:1716-1727) for a whole validation set.
 * Predictions and ground-truth boxes of all images are concatenated: pred_boxes [NP,4] fp32, pred_cls [NP] int32,
 * pred_off [N+1] int32 (image boundaries); targ_boxes [NT,4] fp32, targ_cls [NT] int32, targ_img [NT] int32 (image of
 * each box); thresholds [T] fp32 (all DEVICE).  is_correct [T][NP] uint8 is cleared and then set to 1 where a
 * ground-truth box's best same-category prediction (first maximal IoU) exceeds the threshold. */
int rn_map_match(const float *pred_boxes, const int32_t *pred_cls, const int32_t *pred_off, const float *targ_boxes,
                 const int32_t *targ_cls, const int32_t *targ_img, int NT, int NP, const float *thresholds, int T,
                 unsigned char *is_correct, void *stream);

/* The precision/recall integration of mAP1 (Vision.py:1729-1747) for every (threshold, category) pair in one launch
 * (SURVEY.md section 8f row 4): sort the category's predictions by (score, is_correct) descending, running count of correct
 * ones times 1/n, running maximum from the right, pairwise sum over the correct positions, divided by the category's number
 * of ground-truth boxes -- float64, bit for bit what NumPy computes (nan for a category without ground truth).
 *   pred_scores [NP] fp32; perm [NP] int32 = prediction indices grouped by category, cls_off [C+1] int32 the group boundaries;
 *   is_correct [T][NP] uint8 from rn_map_match; ntrue [C] int32; table [T][C] float64 (all DEVICE). */
size_t rn_map_ap_workspace_bytes(int NP, int T);
int rn_map_ap(const float *pred_scores, const int32_t *perm, const int32_t *cls_off, const unsigned char *is_correct,
              const int32_t *ntrue, int NP, int C, int T, double *table, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif
