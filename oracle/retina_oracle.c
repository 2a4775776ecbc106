/*
 * retina_oracle.c -- CPU restatement of the reference's RetinaNet loss / post-processing path.
 *
 * TEST INFRASTRUCTURE ONLY (see retina_oracle.h).  Written from the behaviour of the reference's
 * Python (citations are file:line under /root/reference); no reference source is copied.  Every fp32
 * operation is rounded individually, in the order the reference's tensor expressions apply them:
 * build with -ffp-contract=off (oracle/Makefile) so gcc never fuses a multiply-add.
 */
#include "retina_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------------
 * A minimal pthread parallel-for.  Images are independent on this path (Vision.py:1636-1641,
 * retinanet.py:756); threads only spread them over host cores for the CPU-baseline timing and
 * every per-image result is combined in image order afterwards, so results do not depend on the
 * thread count.  ORC_NUM_THREADS overrides the default (all online cores).
 * ---------------------------------------------------------------------------------------------- */
typedef void (*orc_item_fn)(int i, void *ctx);
typedef struct {
    orc_item_fn fn;
    void *ctx;
    int n;
    int next;
    pthread_mutex_t mu;
} orc_pf_t;

static void *orc_pf_worker(void *arg) {
    orc_pf_t *pf = (orc_pf_t *)arg;
    for (;;) {
        pthread_mutex_lock(&pf->mu);
        int i = pf->next++;
        pthread_mutex_unlock(&pf->mu);
        if (i >= pf->n) break;
        pf->fn(i, pf->ctx);
    }
    return NULL;
}

int orc_num_threads(void) {
    const char *e = getenv("ORC_NUM_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

static void orc_parallel_for(int n, orc_item_fn fn, void *ctx) {
    int nt = orc_num_threads();
    if (nt > n) nt = n;
    if (nt <= 1) {
        for (int i = 0; i < n; ++i) fn(i, ctx);
        return;
    }
    orc_pf_t pf = {fn, ctx, n, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nt - 1; ++t) {
        if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
    }
    orc_pf_worker(&pf);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------------
 * Anchors: Applications/VisionModels/retinanet.py:439-495
 * ---------------------------------------------------------------------------------------------- */

int orc_num_anchors(int H, int W, int anchors_per_cell) {
    int total = 0;
    for (int l = 3; l < 3 + ORC_NUM_LEVELS; ++l) {
        int s = 1 << l;
        total += ((H + s - 1) / s) * ((W + s - 1) / s); /* retinanet.py:488 */
    }
    return total * anchors_per_cell;
}

void orc_base_anchors(const double *ratios, int nr, const double *scales, int ns, double *base) {
    /* Scales = tile(scales, nr), Ratios = repeat(ratios, ns): anchor k = ir*ns + is (ratio-major),
     * retinanet.py:445-446.  h = s / sqrt(r), w = s * sqrt(r), row = [-w/2, -h/2, w/2, h/2]
     * (:447-451); scaled by the level size 2^(l+2) (:480, :492).  All float64. */
    int K = nr * ns;
    for (int l = 0; l < ORC_NUM_LEVELS; ++l) {
        double size = (double)(1 << (l + 3 + 2));
        for (int ir = 0; ir < nr; ++ir) {
            for (int is = 0; is < ns; ++is) {
                int k = ir * ns + is;
                double sq = sqrt(ratios[ir]);
                double h = scales[is] / sq;
                double w = scales[is] * sq;
                double *row = base + ((size_t)l * K + k) * 4;
                row[0] = size * (-w / 2.0);
                row[1] = size * (-h / 2.0);
                row[2] = size * (w / 2.0);
                row[3] = size * (h / 2.0);
            }
        }
    }
}

int orc_anchors(int H, int W, const double *ratios, int nr, const double *scales, int ns, float *out) {
    int K = nr * ns;
    double *base = (double *)malloc(sizeof(double) * ORC_NUM_LEVELS * K * 4);
    orc_base_anchors(ratios, nr, scales, ns, base);
    size_t a = 0;
    for (int l = 0; l < ORC_NUM_LEVELS; ++l) {
        int stride = 1 << (l + 3);
        int gh = (H + stride - 1) / stride, gw = (W + stride - 1) / stride;
        for (int iy = 0; iy < gh; ++iy) {
            double sy = ((double)iy + 0.5) * (double)stride; /* retinanet.py:459 */
            for (int ix = 0; ix < gw; ++ix) {
                double sx = ((double)ix + 0.5) * (double)stride; /* retinanet.py:458 */
                for (int k = 0; k < K; ++k) {
                    const double *b = base + ((size_t)l * K + k) * 4;
                    /* f64 add (retinanet.py:468), then TEN() rounds to f32 (Core.py:61-62) */
                    out[a * 4 + 0] = (float)(b[0] + sx);
                    out[a * 4 + 1] = (float)(b[1] + sy);
                    out[a * 4 + 2] = (float)(b[2] + sx);
                    out[a * 4 + 3] = (float)(b[3] + sy);
                    ++a;
                }
            }
        }
    }
    free(base);
    return (int)a;
}

/* ------------------------------------------------------------------------------------------------
 * IoU and assignment: Applications/Vision.py:234-256, :1474-1511
 * ---------------------------------------------------------------------------------------------- */

static inline float box_area(const float *b) {
    /* (x2 - x1) * (y2 - y1), Vision.py:248-249 / retinanet.py:517-518 */
    float w = b[2] - b[0];
    float h = b[3] - b[1];
    return w * h;
}

static inline float iou_f32(const float *b1, float area1, const float *b2, float area2) {
    /* Vision.py:251-256 and retinanet.py:506-521: each op rounded to fp32 on its own. */
    float iw = fminf(b1[2], b2[2]) - fmaxf(b1[0], b2[0]);
    if (!(iw > 0.0f)) iw = (iw != iw) ? iw : 0.0f; /* clamp(min=0) */
    float ih = fminf(b1[3], b2[3]) - fmaxf(b1[1], b2[1]);
    if (!(ih > 0.0f)) ih = (ih != ih) ? ih : 0.0f;
    float inter = iw * ih;
    float uni = (area1 + area2) - inter;
    return inter / uni;
}

/* GT rows are already compacted here. */
static int assign_compact(const float *anchors, int A, const float *gt, int m, float pos_thr,
                          float neg_thr, int32_t *matches, float *max_iou) {
    int npos = 0;
    if (m == 0) { /* Vision.py:1498-1501: no objects -> every anchor is a negative */
        for (int a = 0; a < A; ++a) {
            matches[a] = ORC_MATCH_NEG;
            if (max_iou) max_iou[a] = 0.0f;
        }
        return 0;
    }
    float *garea = (float *)malloc(sizeof(float) * (size_t)m);
    for (int j = 0; j < m; ++j) garea[j] = box_area(gt + 4 * j);
    for (int a = 0; a < A; ++a) {
        const float *ab = anchors + 4 * (size_t)a;
        float aarea = box_area(ab);
        float best = iou_f32(gt, garea[0], ab, aarea);
        int bi = 0;
        for (int j = 1; j < m; ++j) { /* torch.max(dim=0): first maximal index, Vision.py:1505 */
            float v = iou_f32(gt + 4 * j, garea[j], ab, aarea);
            if (v > best) {
                best = v;
                bi = j;
            }
        }
        if (best > pos_thr) { /* Vision.py:1506, :1508-1509 */
            matches[a] = bi;
            ++npos;
        } else if (best < neg_thr) { /* Vision.py:1507 */
            matches[a] = ORC_MATCH_NEG;
        } else {
            matches[a] = ORC_MATCH_IGNORE;
        }
        if (max_iou) max_iou[a] = best;
    }
    free(garea);
    return npos;
}

static int compact_gt(const float *gt_boxes, const int64_t *gt_cats, int M, float *boxes,
                      int64_t *cats) {
    /* Vision.py:1637-1638 strips the -1 padding; a row is padding iff its category is negative. */
    int m = 0;
    for (int j = 0; j < M; ++j) {
        if (gt_cats[j] >= 0) {
            memcpy(boxes + 4 * m, gt_boxes + 4 * j, 4 * sizeof(float));
            cats[m] = gt_cats[j];
            ++m;
        }
    }
    return m;
}

int orc_assign(const float *anchors, int A, const float *gt_boxes, const int64_t *gt_cats, int M,
               float pos_thr, float neg_thr, int32_t *matches, float *max_iou) {
    float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)(M > 0 ? M : 1));
    int64_t *cats = (int64_t *)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
    int m = compact_gt(gt_boxes, gt_cats, M, boxes, cats);
    int npos = assign_compact(anchors, A, boxes, m, pos_thr, neg_thr, matches, max_iou);
    free(boxes);
    free(cats);
    return npos;
}

/* ------------------------------------------------------------------------------------------------
 * Loss forward + backward: Applications/Vision.py:1513-1644
 * ---------------------------------------------------------------------------------------------- */

static inline float pow_gamma(float x, float gamma) {
    /* torch special-cases exponent 2 as x*x (verified bitwise) */
    if (gamma == 2.0f) return x * x;
    if (gamma == 1.0f) return x;
    if (gamma == 0.0f) return 1.0f;
    return powf(x, gamma);
}

static inline float pow_gamma_m1(float x, float gamma) { /* x^(gamma-1), used by pow's backward */
    if (gamma == 2.0f) return x;
    if (gamma == 1.0f) return 1.0f;
    return powf(x, gamma - 1.0f);
}

typedef struct {
    const float *anchors, *clas, *reg, *gt_boxes;
    const int64_t *gt_cats;
    int B, A, C, M, B_global;
    double alpha, gamma, beta;
    float pos_thr, neg_thr;
    float *dclas, *dreg;
    int32_t *matches_out, *npos_out;
    float *reg_per_img, *clas_per_img;
} orc_loss_ctx_t;

/* One iteration of the per-image loop of SSD_loss.__call__ (Vision.py:1636-1641) = ssd1 (:1568-1605). */
static void loss_one_image(int i, void *vctx) {
    const orc_loss_ctx_t *k = (const orc_loss_ctx_t *)vctx;
    const float *anchors = k->anchors, *clas = k->clas, *reg = k->reg, *gt_boxes = k->gt_boxes;
    const int64_t *gt_cats = k->gt_cats;
    const int A = k->A, C = k->C, M = k->M;
    float *dclas = k->dclas, *dreg = k->dreg;
    int32_t *matches_out = k->matches_out, *npos_out = k->npos_out;
    float *reg_per_img = k->reg_per_img, *clas_per_img = k->clas_per_img;
    const float pos_thr = k->pos_thr, neg_thr = k->neg_thr;
    const float lo = (float)1e-4, hi = (float)(1.0 - 1e-4); /* Vision.py:1524 */
    const float a_pos = (float)k->alpha, a_neg = (float)(1.0 - k->alpha); /* Vision.py:1526 */
    const float gam = (float)k->gamma;
    const float w_reg = (float)(1.0 - k->beta), w_clas = (float)k->beta; /* Vision.py:1644 */
    const float bs = (float)k->B_global;
    const float inv_std[4] = {(float)0.1, (float)0.1, (float)0.2, (float)0.2}; /* Vision.py:1562 */
    const float sl1_knee = (float)(1.0 / 9.0), sl1_off = (float)(0.5 / 9.0);  /* Vision.py:1565 */
    const int Mc = M > 0 ? M : 1;

    float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)Mc);
    int64_t *cats = (int64_t *)malloc(sizeof(int64_t) * (size_t)Mc);
    int32_t *matches = (int32_t *)malloc(sizeof(int32_t) * (size_t)A);
    const float *clas_i = clas + (size_t)i * A * C;
    const float *reg_i = reg + (size_t)i * A * 4;
    float *dclas_i = dclas ? dclas + (size_t)i * A * C : NULL;
    float *dreg_i = dreg ? dreg + (size_t)i * A * 4 : NULL;

    int m = compact_gt(gt_boxes + (size_t)i * M * 4, gt_cats + (size_t)i * M, M, boxes, cats);
    int npos = assign_compact(anchors, A, boxes, m, pos_thr, neg_thr, matches, NULL);
    if (matches_out) memcpy(matches_out + (size_t)i * A, matches, sizeof(int32_t) * (size_t)A);
    if (npos_out) npos_out[i] = npos;

    /* ---- focal loss over pos+neg anchors, Vision.py:1513-1530 ---- */
    float n_norm = (float)npos; /* target.sum() */
    if (n_norm < 1.0f) n_norm = 1.0f;
    /* upstream gradient of this image's clas loss: beta, then / bs (Vision.py:1644), then / N */
    float g_l = (w_clas / bs) / n_norm;
    double csum = 0.0;
    for (int a = 0; a < A; ++a) {
        int mt = matches[a];
        const float *x = clas_i + (size_t)a * C;
        float *dx = dclas_i ? dclas_i + (size_t)a * C : NULL;
        if (mt == ORC_MATCH_IGNORE) {
            if (dx) memset(dx, 0, sizeof(float) * (size_t)C);
            continue;
        }
        int64_t cat = (mt >= 0) ? cats[mt] : -1; /* Vision.py:1588-1593 */
        for (int c = 0; c < C; ++c) {
            float t = (c == cat) ? 1.0f : 0.0f;
            float u = 1.0f - t;
            float xv = x[c];
            float p = xv < lo ? lo : (xv > hi ? hi : xv);
            float q = 1.0f - p;
            float pt = p * t + q * u;                 /* Vision.py:1525 */
            float wa = a_pos * t + a_neg * u;         /* Vision.py:1526 */
            float r = 1.0f - pt;
            float w = wa * pow_gamma(r, gam);         /* Vision.py:1527 */
            float lp = logf(p), lq = logf(q);
            float inner = t * lp + u * lq;
            float l = (-w) * inner;                   /* Vision.py:1528 */
            csum += (double)l;
            if (dx) {
                /* reverse-mode through the expression graph above, fp32 like autograd */
                float g_inner = g_l * (-w);
                float g_w = -(g_l * inner);
                float g_r = (g_w * wa) * (gam * pow_gamma_m1(r, gam));
                float g_pt = -g_r;
                float gp1 = g_pt * t;           /* via p*t            */
                float gp2 = -(g_pt * u);        /* via (1-p)*(1-t)    */
                float gp3 = (g_inner * t) / p;  /* via log(p)         */
                float gp4 = -((g_inner * u) / q); /* via log(1-p)     */
                float g_p = (gp1 + gp2) + (gp3 + gp4);
                dx[c] = (xv >= lo && xv <= hi) ? g_p : 0.0f; /* clamp backward, inclusive */
            }
        }
    }
    float clas_loss_i = (float)csum / n_norm; /* Vision.py:1529-1530 */

    /* ---- smooth L1 over positives, Vision.py:1532-1566 ---- */
    float reg_loss_i = 0.0f; /* Vision.py:1604 */
    if (dreg_i) memset(dreg_i, 0, sizeof(float) * 4 * (size_t)A);
    if (npos > 0) {
        double rsum = 0.0;
        float numel = (float)(4 * npos);
        float g_e = (w_reg / bs) / numel; /* mean() backward */
        for (int a = 0; a < A; ++a) {
            int mt = matches[a];
            if (mt < 0) continue;
            const float *an = anchors + 4 * (size_t)a;
            const float *tg = boxes + 4 * mt;
            const float *pr = reg_i + 4 * (size_t)a;
            float aw = an[2] - an[0], ah = an[3] - an[1];
            float acx = an[0] + 0.5f * aw, acy = an[1] + 0.5f * ah;
            float tw = tg[2] - tg[0], th = tg[3] - tg[1];
            float tcx = tg[0] + 0.5f * tw, tcy = tg[1] + 0.5f * th;
            if (tw < 1.0f) tw = 1.0f; /* Vision.py:1553-1554 */
            if (th < 1.0f) th = 1.0f;
            float ts[4];
            ts[0] = (tcx - acx) / aw;
            ts[1] = (tcy - acy) / ah;
            ts[2] = logf(tw / aw);
            ts[3] = logf(th / ah);
            for (int k = 0; k < 4; ++k) {
                float tv = ts[k] / inv_std[k]; /* Vision.py:1562 (a division) */
                float d = tv - pr[k];
                float diff = fabsf(d);
                float l;
                float gd;
                if (diff < sl1_knee) { /* Vision.py:1565 */
                    l = 4.5f * (diff * diff);
                    gd = (g_e * 4.5f) * (2.0f * diff);
                } else {
                    l = diff - sl1_off;
                    gd = g_e;
                }
                rsum += (double)l;
                if (dreg_i) {
                    float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
                    dreg_i[4 * (size_t)a + k] = -(gd * sg); /* d|t-p|/dp = -sign(t-p) */
                }
            }
        }
        reg_loss_i = (float)rsum / numel; /* losses.mean(), Vision.py:1566 */
    }
    reg_per_img[i] = reg_loss_i;
    clas_per_img[i] = clas_loss_i;
    free(boxes);
    free(cats);
    free(matches);
}

void orc_loss(const float *anchors, const float *clas, const float *reg, const float *gt_boxes,
              const int64_t *gt_cats, int B, int A, int C, int M, double alpha, double gamma,
              double beta, int B_global, float pos_thr, float neg_thr, float *out3, float *dclas,
              float *dreg, int32_t *matches_out, int32_t *npos_out) {
    const float w_reg = (float)(1.0 - beta), w_clas = (float)beta; /* Vision.py:1644 */
    const float bs = (float)B_global;
    float *reg_per_img = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    float *clas_per_img = (float *)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
    orc_loss_ctx_t ctx = {anchors, clas, reg, gt_boxes, gt_cats, B, A, C, M, B_global, alpha, gamma, beta,
                          pos_thr, neg_thr, dclas, dreg, matches_out, npos_out, reg_per_img, clas_per_img};
    orc_parallel_for(B, loss_one_image, &ctx);
    float reg_total = 0.0f, clas_total = 0.0f; /* Vision.py:1634, :1640-1641: fp32 accumulation */
    for (int i = 0; i < B; ++i) {
        reg_total += reg_per_img[i];
        clas_total += clas_per_img[i];
    }
    float reg_loss = reg_total / bs, clas_loss = clas_total / bs; /* Vision.py:1643 */
    out3[0] = w_reg * reg_loss + w_clas * clas_loss;               /* Vision.py:1644 */
    out3[1] = reg_loss;
    out3[2] = clas_loss;
    free(reg_per_img);
    free(clas_per_img);
}

/* torch.sigmoid in fp32: 1 / (1 + exp(-z)) (the head's output activation, retinanet.py:258, :286). */
void orc_sigmoid(const float *z, size_t n, float *y) {
    for (size_t i = 0; i < n; ++i) y[i] = 1.0f / (1.0f + expf(-z[i]));
}

/* SSD_loss on LOGITS: the head's nn.Sigmoid (retinanet.py:258, :286) followed by SSD_loss, and the gradient
 * chained through sigmoid's backward (grad * (1 - y) * y).  SURVEY.md section 8f row 1. */
void orc_loss_logits(const float *anchors, const float *logits, const float *reg, const float *gt_boxes,
                     const int64_t *gt_cats, int B, int A, int C, int M, double alpha, double gamma, double beta,
                     int B_global, float pos_thr, float neg_thr, float *out3, float *dlogits, float *dreg,
                     int32_t *matches_out, int32_t *npos_out) {
    size_t n = (size_t)B * A * C;
    float *y = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    orc_sigmoid(logits, n, y);
    orc_loss(anchors, y, reg, gt_boxes, gt_cats, B, A, C, M, alpha, gamma, beta, B_global, pos_thr, neg_thr, out3,
             dlogits, dreg, matches_out, npos_out);
    if (dlogits)
        for (size_t i = 0; i < n; ++i) dlogits[i] = (dlogits[i] * (1.0f - y[i])) * y[i]; /* sigmoid_backward */
    free(y);
}

/* ------------------------------------------------------------------------------------------------
 * Post-processing: Applications/VisionModels/retinanet.py:523-812
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    float score;
    int32_t idx;
} orc_cand_t;

static int cand_cmp(const void *pa, const void *pb) {
    const orc_cand_t *a = (const orc_cand_t *)pa, *b = (const orc_cand_t *)pb;
    if (a->score > b->score) return -1; /* descending score, retinanet.py:573 */
    if (a->score < b->score) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx); /* ties: ascending index (documented choice) */
}

int orc_nms(const float *boxes, const int64_t *classes, const float *scores, int n, float max_overlap,
            int top_k, int max_boxes, int32_t *keep_idx) {
    if (n <= 0) return 0; /* retinanet.py:570 */
    orc_cand_t *cand = (orc_cand_t *)malloc(sizeof(orc_cand_t) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        cand[i].score = scores[i];
        cand[i].idx = i;
    }
    qsort(cand, (size_t)n, sizeof(orc_cand_t), cand_cmp);
    int k = n < top_k ? n : top_k; /* retinanet.py:574-576 */
    if (k < 0) k = 0;
    unsigned char *removed = (unsigned char *)calloc((size_t)(k > 0 ? k : 1), 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
    for (int i = 0; i < k; ++i) area[i] = box_area(boxes + 4 * (size_t)cand[i].idx);
    int nkeep = 0, nout = 0;
    for (int i = 0; i < k; ++i) { /* greedy loop, retinanet.py:590-602 */
        if (removed[i]) continue;
        if (nkeep < max_boxes) keep_idx[nout++] = cand[i].idx; /* cap, retinanet.py:702-704 */
        ++nkeep;
        if (nkeep >= max_boxes) break; /* later keeps cannot reach the output */
        const float *bi = boxes + 4 * (size_t)cand[i].idx;
        int64_t ci = classes[cand[i].idx];
        for (int j = i + 1; j < k; ++j) {
            if (removed[j] || classes[cand[j].idx] != ci) continue;
            float v = iou_f32(bi, area[i], boxes + 4 * (size_t)cand[j].idx, area[j]);
            if (v > max_overlap) removed[j] = 1; /* retinanet.py:592-594 */
        }
    }
    free(cand);
    free(removed);
    free(area);
    return nout;
}

void orc_decode_one(const float *an, const float *rg, const float *mean, const float *std, int img_h,
                    int img_w, float *box) {
    float w = an[2] - an[0], h = an[3] - an[1];            /* retinanet.py:750-751 */
    float cx = an[0] + 0.5f * w, cy = an[1] + 0.5f * h;    /* retinanet.py:752-753 */
    float dx = rg[0] * std[0] + mean[0];                   /* retinanet.py:772-775 */
    float dy = rg[1] * std[1] + mean[1];
    float dw = rg[2] * std[2] + mean[2];
    float dh = rg[3] * std[3] + mean[3];
    float pcx = cx + w * dx, pcy = cy + h * dy;            /* retinanet.py:777-778 */
    float pw = w * expf(dw), ph = h * expf(dh);            /* retinanet.py:779-780 */
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph;      /* retinanet.py:782-785 */
    float x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
    if (x1 < 0.0f) x1 = 0.0f;                              /* retinanet.py:790-793 */
    if (y1 < 0.0f) y1 = 0.0f;
    if (x2 > (float)img_w) x2 = (float)img_w;
    if (y2 > (float)img_h) y2 = (float)img_h;
    box[0] = x1;
    box[1] = y1;
    box[2] = x2;
    box[3] = y2;
}

int orc_postproc(const float *clas, const float *reg, const float *anchors, int A, int C, int img_h,
                 int img_w, const float *mean, const float *std, float thresh, float max_overlap,
                 int top_k, int max_boxes, float *out_boxes, int64_t *out_classes, float *out_scores,
                 int32_t *out_anchor_idx, int32_t *n_candidates) {
    float *cb = (float *)malloc(sizeof(float) * 4 * (size_t)A);
    int64_t *cc = (int64_t *)malloc(sizeof(int64_t) * (size_t)A);
    float *cs = (float *)malloc(sizeof(float) * (size_t)A);
    int32_t *ca = (int32_t *)malloc(sizeof(int32_t) * (size_t)A);
    int n = 0;
    for (int a = 0; a < A; ++a) {
        const float *row = clas + (size_t)a * C;
        float best = row[0];
        int bc = 0;
        for (int c = 1; c < C; ++c) { /* clas[i].max(dim=1): first maximal class, retinanet.py:759 */
            if (row[c] > best) {
                best = row[c];
                bc = c;
            }
        }
        if (!(best > thresh)) continue; /* strict, retinanet.py:760 */
        float box[4];
        orc_decode_one(anchors + 4 * (size_t)a, reg + 4 * (size_t)a, mean, std, img_h, img_w, box);
        if (!((box[2] - box[0]) > 0.0f && (box[3] - box[1]) > 0.0f)) continue; /* :796-798 */
        memcpy(cb + 4 * (size_t)n, box, sizeof(box));
        cc[n] = bc;
        cs[n] = best;
        ca[n] = a;
        ++n;
    }
    if (n_candidates) *n_candidates = n;
    int32_t *keep = (int32_t *)malloc(sizeof(int32_t) * (size_t)(max_boxes > 0 ? max_boxes : 1));
    int nk = orc_nms(cb, cc, cs, n, max_overlap, top_k, max_boxes, keep);
    for (int i = 0; i < nk; ++i) {
        int j = keep[i];
        memcpy(out_boxes + 4 * (size_t)i, cb + 4 * (size_t)j, 4 * sizeof(float));
        out_classes[i] = cc[j];
        out_scores[i] = cs[j];
        if (out_anchor_idx) out_anchor_idx[i] = ca[j];
    }
    free(cb);
    free(cc);
    free(cs);
    free(ca);
    free(keep);
    return nk;
}

typedef struct {
    const float *clas, *reg, *anchors, *mean, *std;
    int A, C, img_h, img_w, top_k, max_boxes;
    float thresh, max_overlap;
    float *out_boxes, *out_scores;
    int64_t *out_classes;
    int32_t *out_anchor_idx, *counts, *n_candidates;
} orc_post_ctx_t;

static void post_one_image(int i, void *vctx) {
    const orc_post_ctx_t *k = (const orc_post_ctx_t *)vctx;
    size_t o = (size_t)i * (size_t)k->max_boxes;
    k->counts[i] = orc_postproc(k->clas + (size_t)i * k->A * k->C, k->reg + (size_t)i * k->A * 4, k->anchors,
                                k->A, k->C, k->img_h, k->img_w, k->mean, k->std, k->thresh, k->max_overlap,
                                k->top_k, k->max_boxes, k->out_boxes + 4 * o, k->out_classes + o,
                                k->out_scores + o, k->out_anchor_idx ? k->out_anchor_idx + o : NULL,
                                k->n_candidates ? k->n_candidates + i : NULL);
}

void orc_postproc_batch(const float *clas, const float *reg, const float *anchors, int B, int A, int C,
                        int img_h, int img_w, const float *mean, const float *std, float thresh,
                        float max_overlap, int top_k, int max_boxes, float *out_boxes,
                        int64_t *out_classes, float *out_scores, int32_t *out_anchor_idx,
                        int32_t *counts, int32_t *n_candidates) {
    /* the per-image loop of BBoxPredictor.__call__, retinanet.py:756 */
    orc_post_ctx_t ctx = {clas, reg, anchors, mean, std, A, C, img_h, img_w, top_k, max_boxes, thresh,
                          max_overlap, out_boxes, out_scores, out_classes, out_anchor_idx, counts,
                          n_candidates};
    orc_parallel_for(B, post_one_image, &ctx);
}
