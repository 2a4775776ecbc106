"""ctypes/numpy front end of the C oracle (oracle/retina_oracle.c).  TEST INFRASTRUCTURE ONLY.

Everything here takes and returns numpy arrays on the host.  See retina_oracle.h for the
reference file:line each entry point restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libretina_oracle.so")

MATCH_NEG = -1
MATCH_IGNORE = -2

DEFAULT_RATIOS = [0.5, 1, 2]
DEFAULT_SCALES = [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)]  # retinanet.py:477

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    """Compiles libretina_oracle.so with gcc (oracle/Makefile) if it is missing or stale."""
    src = os.path.join(_HERE, "retina_oracle.c")
    hdr = os.path.join(_HERE, "retina_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libretina_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_num_anchors.restype = C.c_int
        L.orc_num_anchors.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_base_anchors.restype = None
        L.orc_base_anchors.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p]
        L.orc_anchors.restype = C.c_int
        L.orc_anchors.argtypes = [C.c_int, C.c_int, _f64p, C.c_int, _f64p, C.c_int, _f32p]
        L.orc_assign.restype = C.c_int
        L.orc_assign.argtypes = [_f32p, C.c_int, _f32p, _i64p, C.c_int, C.c_float, C.c_float, _i32p, _f32p]
        L.orc_loss.restype = None
        L.orc_loss.argtypes = [_f32p, _f32p, _f32p, _f32p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_double, C.c_double, C.c_double, C.c_int, C.c_float, C.c_float,
                               _f32p, _f32p, _f32p, _i32p, _i32p]
        L.orc_loss_logits.restype = None
        L.orc_loss_logits.argtypes = L.orc_loss.argtypes
        L.orc_sigmoid.restype = None
        L.orc_sigmoid.argtypes = [_f32p, C.c_size_t, _f32p]
        L.orc_nms.restype = C.c_int
        L.orc_nms.argtypes = [_f32p, _i64p, _f32p, C.c_int, C.c_float, C.c_int, C.c_int, _i32p]
        L.orc_postproc_batch.restype = None
        L.orc_postproc_batch.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         _f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int,
                                         _f32p, _i64p, _f32p, _i32p, _i32p, _i32p]
        L.orc_decode_one.restype = None
        L.orc_decode_one.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p]
        _lib = L
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def num_threads():
    return int(lib().orc_num_threads())


def num_anchors(H, W, per_cell=9):
    return int(lib().orc_num_anchors(int(H), int(W), int(per_cell)))


def base_anchors(ratios=None, scales=None):
    r = np.ascontiguousarray(DEFAULT_RATIOS if ratios is None else ratios, dtype=np.float64)
    s = np.ascontiguousarray(DEFAULT_SCALES if scales is None else scales, dtype=np.float64)
    out = np.empty((5, len(r) * len(s), 4), dtype=np.float64)
    lib().orc_base_anchors(_ptr(r, _f64p), len(r), _ptr(s, _f64p), len(s), _ptr(out, _f64p))
    return out


def anchors(H, W, ratios=None, scales=None):
    r = np.ascontiguousarray(DEFAULT_RATIOS if ratios is None else ratios, dtype=np.float64)
    s = np.ascontiguousarray(DEFAULT_SCALES if scales is None else scales, dtype=np.float64)
    A = num_anchors(H, W, len(r) * len(s))
    out = np.empty((A, 4), dtype=np.float32)
    n = lib().orc_anchors(int(H), int(W), _ptr(r, _f64p), len(r), _ptr(s, _f64p), len(s), _ptr(out, _f32p))
    assert n == A
    return out


def assign(anchors_, gt_boxes, gt_cats, pos_thr=0.5, neg_thr=0.4):
    """One image. Returns (matches[A] int32, npos, max_iou[A] float32)."""
    an = _f32(anchors_)
    gb = _f32(gt_boxes).reshape(-1, 4)
    gc = np.ascontiguousarray(gt_cats, dtype=np.int64).reshape(-1)
    A, M = an.shape[0], gc.shape[0]
    matches = np.empty(A, dtype=np.int32)
    miou = np.empty(A, dtype=np.float32)
    npos = lib().orc_assign(_ptr(an, _f32p), A, _ptr(gb, _f32p), _ptr(gc, _i64p), M,
                            np.float32(pos_thr), np.float32(neg_thr), _ptr(matches, _i32p), _ptr(miou, _f32p))
    return matches, int(npos), miou


def loss(anchors_, clas, reg, gt_boxes, gt_cats, alpha=0.25, gamma=2.0, beta=0.5, B_global=None,
         pos_thr=0.5, neg_thr=0.4, want_grads=True, want_matches=False, from_logits=False):
    """Batch loss. Returns dict(out3, dclas, dreg, matches, npos).  from_logits: `clas` holds logits, the
    sigmoid of the classification head is applied first and dclas is the gradient w.r.t. the logits."""
    an, cl, rg = _f32(anchors_), _f32(clas), _f32(reg)
    B, A, Cc = cl.shape
    gb = _f32(gt_boxes).reshape(B, -1, 4)
    gc = np.ascontiguousarray(gt_cats, dtype=np.int64).reshape(B, -1)
    M = gc.shape[1]
    out3 = np.zeros(3, dtype=np.float32)
    dclas = np.empty_like(cl) if want_grads else None
    dreg = np.empty_like(rg) if want_grads else None
    matches = np.empty((B, A), dtype=np.int32) if want_matches else None
    npos = np.empty(B, dtype=np.int32)
    fn = lib().orc_loss_logits if from_logits else lib().orc_loss
    fn(_ptr(an, _f32p), _ptr(cl, _f32p), _ptr(rg, _f32p), _ptr(gb, _f32p), _ptr(gc, _i64p),
       B, A, Cc, M, float(alpha), float(gamma), float(beta), int(B_global or B),
       np.float32(pos_thr), np.float32(neg_thr), _ptr(out3, _f32p), _ptr(dclas, _f32p),
       _ptr(dreg, _f32p), _ptr(matches, _i32p), _ptr(npos, _i32p))
    return dict(out3=out3, dclas=dclas, dreg=dreg, matches=matches, npos=npos)


def sigmoid(z):
    """fp32 sigmoid exactly as orc_loss_logits applies it."""
    z = _f32(z)
    y = np.empty_like(z)
    lib().orc_sigmoid(_ptr(z, _f32p), z.size, _ptr(y, _f32p))
    return y


def nms(boxes, classes, scores, max_overlap=0.5, top_k=1000, max_boxes=20):
    """Returns keep indices (into the inputs), score-descending."""
    b = _f32(boxes).reshape(-1, 4)
    c = np.ascontiguousarray(classes, dtype=np.int64).reshape(-1)
    s = _f32(scores).reshape(-1)
    n = s.shape[0]
    keep = np.empty(max(int(max_boxes), 1), dtype=np.int32)
    k = lib().orc_nms(_ptr(b, _f32p), _ptr(c, _i64p), _ptr(s, _f32p), n, np.float32(max_overlap),
                      int(top_k), int(max_boxes), _ptr(keep, _i32p))
    return keep[:k].copy()


def postproc(clas, reg, anchors_, img_h, img_w, mean=(0., 0., 0., 0.), std=(0.1, 0.1, 0.2, 0.2),
             thresh=0.05, max_overlap=0.5, top_k=1000, max_boxes=20):
    """Batch post-processing. Returns dict(boxes[B,K,4], classes[B,K], scores[B,K], anchor_idx[B,K],
    counts[B], n_candidates[B]) with K = max_boxes."""
    cl, rg, an = _f32(clas), _f32(reg), _f32(anchors_)
    B, A, Cc = cl.shape
    mean_, std_ = _f32(mean), _f32(std)
    K = max(int(max_boxes), 1)
    boxes = np.zeros((B, K, 4), dtype=np.float32)
    classes = np.zeros((B, K), dtype=np.int64)
    scores = np.zeros((B, K), dtype=np.float32)
    aidx = np.full((B, K), -1, dtype=np.int32)
    counts = np.zeros(B, dtype=np.int32)
    ncand = np.zeros(B, dtype=np.int32)
    lib().orc_postproc_batch(_ptr(cl, _f32p), _ptr(rg, _f32p), _ptr(an, _f32p), B, A, Cc, int(img_h),
                             int(img_w), _ptr(mean_, _f32p), _ptr(std_, _f32p), np.float32(thresh),
                             np.float32(max_overlap), int(top_k), int(max_boxes), _ptr(boxes, _f32p),
                             _ptr(classes, _i64p), _ptr(scores, _f32p), _ptr(aidx, _i32p),
                             _ptr(counts, _i32p), _ptr(ncand, _i32p))
    return dict(boxes=boxes, classes=classes, scores=scores, anchor_idx=aidx, counts=counts,
                n_candidates=ncand)


def decode_one(anchor, reg, mean, std, img_h, img_w):
    a, r, m, s = _f32(anchor), _f32(reg), _f32(mean), _f32(std)
    out = np.empty(4, dtype=np.float32)
    lib().orc_decode_one(_ptr(a, _f32p), _ptr(r, _f32p), _ptr(m, _f32p), _ptr(s, _f32p), int(img_h),
                         int(img_w), _ptr(out, _f32p))
    return out


def stage_targets(bboxes, cats, scales, rand_scale=1.0, row_jit=0, col_jit=0):
    """numpy restatement of the target half of AspectRatioCollater (reference Vision.py:770-785 scale + jitter,
    :798-809 padding with -1).  Returns (bboxes_padded [bs,M,4] float32, cats_padded [bs,M] int64)."""
    bs = len(bboxes)
    out = []
    for i in range(bs):
        b = np.asarray(bboxes[i])
        if len(b) > 0:
            b = b * scales[i] * rand_scale                                        # Vision.py:773
            b = np.array([b[:, 0] + col_jit, b[:, 1] + row_jit, b[:, 2] + col_jit, b[:, 3] + row_jit]).T  # :783-784
        out.append(b)
    M = max([len(b) for b in out]) if bs else 0
    if M > 0:                                                                     # Vision.py:799-806
        bp = np.ones((bs, M, 4)).astype(np.float32) * (-1)
        cp = np.ones((bs, M)).astype(np.int64) * (-1)
        for i, (b, c) in enumerate(zip(out, cats)):
            if len(b) > 0:
                bp[i, :len(b), :] = b
                cp[i, :len(c)] = c
    else:                                                                         # Vision.py:807-809
        bp = np.ones((bs, 1, 4)).astype(np.float32) * (-1)
        cp = np.ones((bs, 1)).astype(np.int64) * (-1)
    return bp, cp


def jaccard_f32(a, b):
    """Vision.jaccard (reference Vision.py:234-256) in numpy float32: [n,4] x [m,4] -> [n,m]."""
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
    iw = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0, None)
    ih = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0, None)
    inter = (iw * ih).astype(np.float32)
    return inter / ((area_a + area_b) - inter)


def map_table(predictions, targets, C, thresholds):
    """numpy restatement of mAP / mAP1 (reference Vision.py:1696-1800): the [len(thresholds), C] table of
    per-(threshold, category) average precisions whose mean the reference returns."""
    N = len(predictions)
    table = np.zeros((len(thresholds), C))
    for c in range(C):
        targs = [[b for b, cc in targets[i] if cc == c] for i in range(N)]                       # Vision.py:1781-1789
        sel = [[j for j in range(len(predictions[i][0])) if predictions[i][1][j] == c] for i in range(N)]
        for t, thresh in enumerate(thresholds):
            is_correct, scores = [], []
            for i in range(N):                                                                   # Vision.py:1720-1727
                ic = [0] * len(sel[i])
                if len(sel[i]) > 0 and len(targs[i]) > 0:
                    jac = jaccard_f32(np.array(targs[i]), np.array([predictions[i][0][j] for j in sel[i]]))
                    for j in range(jac.shape[0]):
                        idx = int(np.argmax(jac[j]))          # first maximal index, as torch.max(dim=1)
                        if jac[j, idx] > np.float32(thresh):  # fp32 tensor vs python float compares in fp32
                            ic[idx] = 1
                is_correct += ic
                scores += [predictions[i][2][j] for j in sel[i]]
            pairs = sorted(zip(scores, is_correct), reverse=True)                                # Vision.py:1730-1731
            ic_sorted = np.array([ic for _, ic in pairs])
            L = len(ic_sorted)
            ntrue = sum(len(x) for x in targs)
            tp = np.cumsum(ic_sorted)
            prec = tp * np.array([1 / n for n in range(1, L + 1)])
            prec_max = np.flip(np.maximum.accumulate(np.flip(prec)))
            with np.errstate(divide="ignore", invalid="ignore"):
                table[t, c] = np.sum(prec_max[ic_sorted.nonzero()[0]]) / np.float64(ntrue)       # Vision.py:1741-1747
    return table


def heads_to_flat(levels, n):
    """The tail of the reference's heads + the model's concatenation in numpy: each level [B, K*n, gh, gw] ->
    permute(0,2,3,1).contiguous().view(B, -1, n) (retinanet.py:215-217, :289-295), levels concatenated on axis 1
    (Vision.py:1467-1468).  Returns [B, A, n]."""
    return np.concatenate([np.ascontiguousarray(np.transpose(x, (0, 2, 3, 1))).reshape(x.shape[0], -1, n) for x in levels],
                          axis=1)


def flat_to_heads(flat, level_shapes):
    """Inverse of heads_to_flat (what autograd does to a gradient w.r.t. the flat tensor): [B, A, n] -> list of
    [B, K*n, gh, gw]; level_shapes = [(K*n, gh, gw)]."""
    B, _, n = flat.shape
    out, a0 = [], 0
    for (ch, gh, gw) in level_shapes:
        rows = gh * gw * (ch // n)
        out.append(np.ascontiguousarray(np.transpose(flat[:, a0:a0 + rows].reshape(B, gh, gw, ch), (0, 3, 1, 2))))
        a0 += rows
    assert a0 == flat.shape[1]
    return out


def stage_images(images, row_jit=0, col_jit=0):
    """numpy restatement of the pixel half of AspectRatioCollater after its cv2.resize (reference Vision.py:775-777
    jitter placement, :786 transpose, :790-796 padding).  Returns imgs_padded [bs, C, H, W] float32."""
    out = []
    for img in images:
        rows, cols, channels = img.shape
        new_img = np.zeros((rows + row_jit, cols + col_jit, channels)).astype(np.float32)
        new_img[row_jit:, col_jit:, :] = img.astype(np.float32)
        out.append(new_img.transpose(2, 0, 1))
    max_h = int(32 * np.ceil(max(x.shape[1] for x in out) / 32))
    max_w = int(32 * np.ceil(max(x.shape[2] for x in out) / 32))
    padded = np.zeros((len(out), out[0].shape[0], max_h, max_w)).astype(np.float32)
    for i, x in enumerate(out):
        padded[i, :, :x.shape[1], :x.shape[2]] = x
    return padded
