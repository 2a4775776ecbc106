/*
 * retina_oracle.h -- CPU restatement (plain C) of the reference's RetinaNet loss / post-processing
 * path.  TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call it, and only as the checker or the timed CPU baseline.  The
 * product (neuralnetworklibrary_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference ships no tests or golden vectors for this path, so the
 * oracle is pinned against the live reference imported in the build container
 * (tests/golden/make_golden.py -> the .npz fixtures under tests/golden; tests/test_oracle_golden.py) and, when
 * /root/reference is present, directly (tests/test_oracle_vs_reference.py).
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef RETINA_ORACLE_H
#define RETINA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_LEVELS 5 /* pyramid levels P3..P7, Applications/VisionModels/retinanet.py:478 */
#define ORC_MATCH_NEG (-1)    /* max IoU <  neg_thresh : background, Applications/Vision.py:1507 */
#define ORC_MATCH_IGNORE (-2) /* max IoU in [neg_thresh, pos_thresh] : not used in the loss      */

/* A = 9 * sum_l ceil(H/2^l) * ceil(W/2^l), retinanet.py:488 (9 -> nr*ns in general). */
int orc_num_anchors(int H, int W, int anchors_per_cell);

/* get_anchor_set scaled by the level size 2^(l+2): retinanet.py:439-451, :480, :492.
 * base is [5][nr*ns][4] float64. */
void orc_base_anchors(const double *ratios, int nr, const double *scales, int ns, double *base);

/* AnchorGenerator.__call__: retinanet.py:485-495 (+ TEN's f64->f32 rounding, General/Core.py:61-62).
 * out is [A,4] float32. Returns A. */
int orc_anchors(int H, int W, const double *ratios, int nr, const double *scales, int ns, float *out);

/* jaccard + match_anchors_objects for ONE image: Applications/Vision.py:234-256, :1474-1511.
 * gt rows with gt_cats[j] < 0 are padding and are dropped first (Vision.py:1637-1638); indices in
 * `matches` refer to the compacted list.  matches[a] = gt index | ORC_MATCH_NEG | ORC_MATCH_IGNORE.
 * max_iou may be NULL. Returns the number of positive anchors. */
int orc_assign(const float *anchors, int A, const float *gt_boxes, const int64_t *gt_cats, int M,
               float pos_thr, float neg_thr, int32_t *matches, float *max_iou);

/* SSD_loss.__call__ over a batch, forward and backward: Vision.py:1513-1644.
 * clas [B,A,C] (probabilities), reg [B,A,4], gt_boxes [B,M,4], gt_cats [B,M].
 * out3 = {loss, reg_loss, clas_loss} as the reference returns/stores them when B_global == B; with
 * B_global > B (an image shard of a larger batch) the three values are this shard's contribution,
 * i.e. sums over the local images divided by B_global.
 * dclas/dreg (may be NULL) receive d loss / d clas, d loss / d reg for upstream gradient 1.
 * matches_out [B,A] and npos_out [B] may be NULL. */
void orc_loss(const float *anchors, const float *clas, const float *reg, const float *gt_boxes,
              const int64_t *gt_cats, int B, int A, int C, int M, double alpha, double gamma,
              double beta, int B_global, float pos_thr, float neg_thr, float *out3, float *dclas,
              float *dreg, int32_t *matches_out, int32_t *npos_out);

/* torch.sigmoid in fp32 (the classification head's output activation, retinanet.py:258, :286). */
void orc_sigmoid(const float *z, size_t n, float *y);

/* The same on LOGITS: nn.Sigmoid of the classification head (retinanet.py:258, :286) + SSD_loss, gradient
 * w.r.t. the logits. */
void orc_loss_logits(const float *anchors, const float *logits, const float *reg, const float *gt_boxes,
                     const int64_t *gt_cats, int B, int A, int C, int M, double alpha, double gamma, double beta,
                     int B_global, float pos_thr, float neg_thr, float *out3, float *dlogits, float *dreg,
                     int32_t *matches_out, int32_t *npos_out);

/* nms with rel_thresh/inc/dup = None: retinanet.py:523-711 (sort :573-576, greedy :590-602, cap
 * :702-704).  Ties in score are ordered by ascending input index (the reference's torch.sort is
 * unstable).  keep_idx receives indices into the input arrays, score-descending.
 * Returns the number kept (<= max_boxes). */
int orc_nms(const float *boxes, const int64_t *classes, const float *scores, int n, float max_overlap,
            int top_k, int max_boxes, int32_t *keep_idx);

/* BBoxPredictor.__call__ for ONE image (decode/clip/threshold) followed by nms:
 * retinanet.py:732-812.  clas [A,C], reg [A,4], anchors [A,4]; mean/std are 4 floats each.
 * Outputs hold up to max_boxes rows.  n_candidates (may be NULL) receives the number of boxes that
 * survive the threshold and the empty-box filter (what nms receives).  Returns the number kept. */
int orc_postproc(const float *clas, const float *reg, const float *anchors, int A, int C, int img_h,
                 int img_w, const float *mean, const float *std, float thresh, float max_overlap,
                 int top_k, int max_boxes, float *out_boxes, int64_t *out_classes, float *out_scores,
                 int32_t *out_anchor_idx, int32_t *n_candidates);

/* The per-image loop of BBoxPredictor.__call__ (retinanet.py:756) over a batch; outputs are
 * [B,max_boxes,...], counts [B]; out_anchor_idx and n_candidates ([B]) may be NULL. */
void orc_postproc_batch(const float *clas, const float *reg, const float *anchors, int B, int A, int C,
                        int img_h, int img_w, const float *mean, const float *std, float thresh,
                        float max_overlap, int top_k, int max_boxes, float *out_boxes,
                        int64_t *out_classes, float *out_scores, int32_t *out_anchor_idx,
                        int32_t *counts, int32_t *n_candidates);

/* Host threads the batch entry points (orc_loss, orc_postproc_batch) spread images over:
 * ORC_NUM_THREADS if set, else all online cores. */
int orc_num_threads(void);

/* Decode + clip of a single anchor, exposed for unit tests: retinanet.py:750-753, :772-793. */
void orc_decode_one(const float *anchor, const float *reg, const float *mean, const float *std,
                    int img_h, int img_w, float *box);

#ifdef __cplusplus
}
#endif
#endif
