"""Drop-in replacements for the anchor / post-processing objects of the reference's
Applications/VisionModels/retinanet.py, backed by libretina_sm100.so:

    AnchorGenerator   retinanet.py:473-495   -> rn_anchors (+ geometry tag for on-the-fly anchors)
    BBoxPredictor     retinanet.py:713-812   -> rn_postproc
    nms               retinanet.py:523-711   -> rn_nms  (+ the optional rel_thresh / inc / dup stages
                                                on the <= top_k survivors, on the host as in the reference)

Signatures, return types and side effects follow SURVEY.md section 8b so that
`model.AnchorGenerator = AnchorGenerator()`, `model.BBoxPredictor = BBoxPredictor()` and
`vmods.retinanet.nms = nms` work without touching Vision.py / Learner.py.  They are plain objects (no
nn.Module, no state_dict keys), like the reference's.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

PYRAMID_LEVELS = [3, 4, 5, 6, 7]


def get_anchor_set(ratios=[0.5, 1, 2], scales=[2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)]):
    """Base boxes centred on the origin, one per (ratio, scale) pair, ratio-major, as float64 rows
    [xmin, ymin, xmax, ymax]: area scale^2, width/height = ratio (reference retinanet.py:439-451)."""
    r = np.repeat(np.asarray(ratios, dtype=np.float64), len(scales))
    s = np.tile(np.asarray(scales, dtype=np.float64), len(ratios))
    root = np.sqrt(r)
    half_w, half_h = (s * root) / 2, (s / root) / 2
    return np.stack([-half_w, -half_h, half_w, half_h], axis=1)


class AnchorGeometry(object):
    """What a kernel needs to regenerate the anchors of one image shape on the fly."""
    __slots__ = ("H", "W", "K", "A", "base", "data_ptr", "version")

    def __init__(self, H, W, base, tensor):
        self.H, self.W = int(H), int(W)
        self.base = base                      # float64 [5, K, 4], C-contiguous, host
        self.K = int(base.shape[1])
        self.A = int(tensor.shape[0])
        self.data_ptr = tensor.data_ptr()
        self.version = tensor._version


def anchor_geometry(anchors):
    """The AnchorGeometry tag of a tensor produced by our AnchorGenerator, or None when the tensor
    is anything else (another generator, a copy, or modified in place since)."""
    g = getattr(anchors, "_rn_geom", None)
    if g is None or g.data_ptr != anchors.data_ptr() or g.version != anchors._version or g.A != anchors.shape[0]:
        return None
    return g


class AnchorGenerator(object):
    """Generates the [A,4] anchor table for the shape of an image batch (reference retinanet.py:473-495).

    The table is produced on the device by rn_anchors (float64 add, rounded to float32: bit-identical to
    the reference's NumPy float64 + TEN()) and cached per (H, W, device) instead of being rebuilt on the
    host and copied on every forward.  The returned tensor carries a geometry tag that lets the loss and
    post-processing kernels regenerate anchors on the fly instead of reading the table."""

    def __init__(self, ratios=[0.5, 1, 2], scales=[2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)]):
        self.pyramid_levels = list(PYRAMID_LEVELS)
        self.strides = [2 ** x for x in self.pyramid_levels]
        self.sizes = [2 ** (x + 2) for x in self.pyramid_levels]
        self.ratios = np.array(ratios)
        self.scales = np.array(scales)
        self.anchor_set = get_anchor_set(ratios, scales)
        if self.anchor_set.shape[0] > _lib.MAX_K:
            raise ValueError("at most %d anchors per cell are supported" % _lib.MAX_K)
        # size_l * anchor_set, float64 (reference retinanet.py:492)
        self.base = np.ascontiguousarray(np.stack([s * self.anchor_set for s in self.sizes]), dtype=np.float64)
        self._cache = {}

    def __call__(self, img_batch):
        H, W = int(img_batch.shape[2]), int(img_batch.shape[3])
        device = img_batch.device if img_batch.is_cuda else torch.device("cuda", torch.cuda.current_device())
        key = (H, W, device.index)
        hit = self._cache.get(key)
        if hit is not None and anchor_geometry(hit) is not None:
            return hit
        lib = _lib.load()
        K = self.base.shape[1]
        A = lib.rn_num_anchors(H, W, K)
        if A <= 0:
            raise ValueError("bad image shape %dx%d" % (H, W))
        out = torch.empty((A, 4), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(lib.rn_anchors(H, W, _lib.base_ptr(self.base), K, _lib.ptr(out), _lib.stream_ptr(device)))
        out._rn_geom = AnchorGeometry(H, W, self.base, out)
        self._cache[key] = out
        return out


def anchor_args(anchors, H=None, W=None):
    """(H, W, base_ptr, K, table_ptr, A) for a kernel call: on-the-fly generation when the tensor
    carries a valid geometry tag, else the tensor itself as a table."""
    g = anchor_geometry(anchors)
    A = int(anchors.shape[0])
    if g is not None and (H is None or (g.H == H and g.W == W)):
        return g.H, g.W, _lib.base_ptr(g.base), g.K, None, A
    _lib.require_cuda(anchors, "anchors", torch.float32)
    if not anchors.is_contiguous():
        raise ValueError("anchors must be contiguous")
    return int(H or 0), int(W or 0), None, 1, _lib.ptr(anchors), A


# --------------------------------------------------------------------------------------------------
# nms
# --------------------------------------------------------------------------------------------------
_ws = _lib.Workspace()


def _inter_f32(b1, b2):
    """Pairwise intersection areas in float32 (reference retinanet.py:500-509)."""
    iw = (np.minimum(b1[:, None, 2], b2[None, :, 2]) - np.maximum(b1[:, None, 0], b2[None, :, 0])).clip(0, None)
    ih = (np.minimum(b1[:, None, 3], b2[None, :, 3]) - np.maximum(b1[:, None, 1], b2[None, :, 1])).clip(0, None)
    return iw * ih


def _iou_f32(b1, b2):
    """Pairwise IoU in float32 (reference retinanet.py:511-521)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    inter = _inter_f32(b1, b2)
    return inter / (a1[:, None] + a2[None, :] - inter)


def _host_stages(boxes, classes, scores, rel_thresh, inc, dup):
    """The optional pruning stages of the reference's nms (retinanet.py:612-695) on the score-sorted
    NMS survivors.  boxes [n,4] f32, classes [n] i64, scores [n] f32 (numpy).  Returns a keep index
    array.  Scalar arithmetic is float32 (python-float * np.float32 under NumPy >= 2 promotion)."""
    keep = np.arange(len(scores))
    f32 = np.float32

    if rel_thresh:  # retinanet.py:613-634
        t1, t2 = f32(rel_thresh[0]), f32(rel_thresh[1])
        s = scores[keep]
        below = np.nonzero(s < t1 * s[0])[0]
        if below.size:
            keep = keep[:below[0]]
        s, c = scores[keep], classes[keep]
        same = c[:, None] == c[None, :]
        weak = s[None, :] < (t2 * s)[:, None]          # [i, j]: s_j < t2 * s_i
        later = np.triu(np.ones((len(keep), len(keep)), dtype=bool), k=1)
        drop = (same & weak & later).any(axis=0)
        keep = keep[~drop]

    if inc:  # retinanet.py:641-671
        inc_thresh, inc_classes = inc
        inc_classes = set(int(v) for v in inc_classes)
        b, c, s = boxes[keep], classes[keep], scores[keep]
        L = len(keep)
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        contained = (_inter_f32(b, b) / area) * (c[:, None] == c[None, :]).astype(int)   # [i,j]: share of j inside i
        size_ratio = area[None, :] / area[:, None]                                        # [i,j]: area_j / area_i
        incl = ((contained > inc_thresh).astype(int) - np.identity(L, int)) * (size_ratio > 0.25).astype(int)
        single = [i for i in np.nonzero(incl.sum(axis=1) == 1)[0] if int(c[i]) not in inc_classes]
        partners = set(int(np.argmax(incl[i])) for i in single)
        drop = set()
        for i in single:
            if int(i) in partners:
                continue
            j = int(np.argmax(incl[i]))
            if s[i] < f32(0.75) * s[j]:
                drop.add(int(i))
            elif s[j] < f32(0.75) * s[i]:
                drop.add(j)
        if drop:
            keep = np.array([k for n_, k in enumerate(keep) if n_ not in drop], dtype=keep.dtype)

    if dup:  # retinanet.py:678-695
        dup_thresh, dup_pairs = dup
        pairs = set((int(p), int(q)) for p, q in dup_pairs)
        changed = True
        while changed:
            changed = False
            b, c, s = boxes[keep], classes[keep], scores[keep]
            jac = _iou_f32(b, b)
            L = len(keep)
            for i in range(L - 1):
                hit = [j for j in range(i + 1, L)
                       if jac[i, j] > dup_thresh and (int(c[i]), int(c[j])) in pairs and s[j] < f32(0.75) * s[i]]
                if hit:
                    keep = np.delete(keep, hit[0])
                    changed = True
                    break
    return keep


def _nms_layout(L, K):
    # boxes | classes | scores | idx | counts  (byte offsets into one buffer; boxes 16-byte, classes 8-byte aligned)
    o_box, o_cls, o_sc, o_idx = 0, 16 * L * K, 24 * L * K, 28 * L * K
    o_cnt = 32 * L * K
    return o_box, o_cls, o_sc, o_idx, o_cnt, o_cnt + 4 * L


def nms_batch_device(boxes, classes, scores, offsets, max_overlap=0.5, top_k=1000, max_keep=20, seg_off=None, seg_par=None):
    """rn_nms_batch on device tensors: boxes [n,4] f32, classes [n] i64, scores [n] f32 are the candidates of L images
    concatenated, offsets [L+1] i32 the image boundaries (optionally the TTA un-transform table seg_off [S+1] i32 / seg_par
    [S,5] f64, see include/retina_b200.h).  One launch sequence, no host synchronisation; returns (uint8 device buffer, K)
    laid out by _nms_layout -- decode it after ONE .cpu()."""
    lib = _lib.load()
    dev = boxes.device
    L, n = int(offsets.shape[0]) - 1, int(scores.shape[0])
    S = 0 if seg_off is None else int(seg_off.shape[0]) - 1
    o_box, o_cls, o_sc, o_idx, o_cnt, total = _nms_layout(L, max_keep)
    with torch.cuda.device(dev):
        buf = torch.empty(total, dtype=torch.uint8, device=dev)
        ws = _ws.get(lib.rn_nms_batch_workspace_bytes(n, L, top_k), dev)
        p = buf.data_ptr()
        _lib.check(lib.rn_nms_batch(_lib.ptr(boxes), _lib.ptr(classes), _lib.ptr(scores), _lib.ptr(offsets), L, n,
                                    _lib.ptr(seg_off), _lib.ptr(seg_par), S, float(max_overlap), int(top_k), int(max_keep),
                                    C.c_void_p(p + o_box), C.c_void_p(p + o_cls), C.c_void_p(p + o_sc), C.c_void_p(p + o_idx),
                                    C.c_void_p(p + o_cnt), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
    return buf, max_keep


def decode_nms_buffer(host, L, K):
    """Views into the host copy of nms_batch_device's buffer: boxes [L,K,4] f32, classes [L,K] i64, scores [L,K] f32,
    idx [L,K] i32, counts [L] i32."""
    o_box, o_cls, o_sc, o_idx, o_cnt, total = _nms_layout(L, K)
    return (host[o_box:o_cls].view(np.float32).reshape(L, K, 4), host[o_cls:o_sc].view(np.int64).reshape(L, K),
            host[o_sc:o_idx].view(np.float32).reshape(L, K), host[o_idx:o_cnt].view(np.int32).reshape(L, K),
            host[o_cnt:total].view(np.int32))


def nms(pred_boxes, pred_classes, conf_scores, max_overlap=0.5, rel_thresh=None,
        top_k=1000, max_boxes=20, dup=None, inc=None, print_it=False):
    """Non-maximum suppression for one image (reference retinanet.py:523-711), same arguments and
    return value: three lists (np.ndarray[4] float32, np.int64, np.float32), score-descending.

    Sort, top_k and the class-aware greedy suppression run on the GPU (rn_nms_batch with one image: radix select + bitonic
    sort + bitmask sweep, survivors gathered on the device) and come back in ONE device->host copy; score ties are
    ordered by input position (the reference's sort is unstable)."""
    if len(pred_boxes) == 0:
        return [], [], []
    boxes = torch.as_tensor(pred_boxes)
    dev = boxes.device if boxes.is_cuda else torch.device("cuda", torch.cuda.current_device())
    boxes = boxes.detach().to(device=dev, dtype=torch.float32).contiguous()
    classes = torch.as_tensor(pred_classes).detach().to(device=dev, dtype=torch.int64).contiguous()
    scores = torch.as_tensor(conf_scores).detach().to(device=dev, dtype=torch.float32).contiguous()
    n = int(scores.shape[0])
    if boxes.shape != (n, 4) or classes.shape != (n,):
        raise ValueError("nms expects boxes [n,4], classes [n], scores [n]")
    top_k = int(top_k)
    if top_k < 1:
        return [], [], []
    if top_k > _lib.MAX_TOP_K:
        raise ValueError("top_k > %d is not supported" % _lib.MAX_TOP_K)
    extra = bool(rel_thresh) or bool(inc) or bool(dup)
    max_keep = top_k if extra else max(1, min(int(max_boxes), top_k))
    offsets = torch.tensor([0, n], dtype=torch.int32, device=dev)
    buf, K = nms_batch_device(boxes, classes, scores, offsets, max_overlap, top_k, max_keep)
    kb, kc, ks, _, cnt = decode_nms_buffer(buf.cpu().numpy(), 1, K)   # the one device->host copy (and synchronisation)
    k = int(cnt[0])
    kb, kc, ks = kb[0, :k], kc[0, :k], ks[0, :k]
    if print_it:
        print('after non-max-supress')
        print(len(kb), len(kc), len(ks))
    if extra:
        sel = _host_stages(kb, kc, ks, rel_thresh, inc, dup)
        kb, kc, ks = kb[sel], kc[sel], ks[sel]
    m = max(int(max_boxes), 0)
    return list(kb[:m]), list(kc[:m]), list(ks[:m])


# --------------------------------------------------------------------------------------------------
# BBoxPredictor
# --------------------------------------------------------------------------------------------------
class BBoxPredictor(object):
    """Turns RetinaNet activations into pruned box predictions (reference retinanet.py:713-812): per
    image class max + score threshold + decode (cx += w*dx, w *= exp(dw), ...) + clip to the image +
    drop empty boxes, then nms().  One library call (rn_postproc) handles the whole batch; a single
    device->host copy brings back at most max_boxes rows per image."""

    def __init__(self, mean=[0., 0., 0., 0.], std=[0.1, 0.1, 0.2, 0.2]):
        super().__init__()
        self._mean = np.ascontiguousarray(mean, dtype=np.float32)   # TEN(list) rounds to float32
        self._std = np.ascontiguousarray(std, dtype=np.float32)
        self.mean, self.std = torch.from_numpy(self._mean), torch.from_numpy(self._std)

    def __call__(self, img_batch, reg, clas, anchors, thresh=0.05, max_overlap=0.5,
                 rel_thresh=None, top_k=1000, max_boxes=20, dup=None, inc=None):
        bs, _, height, width = img_batch.shape
        out = self.predict_arrays(int(height), int(width), reg, clas, anchors, thresh, max_overlap, top_k,
                                  max_boxes, full=bool(rel_thresh) or bool(inc) or bool(dup))
        boxes, classes, scores, counts = out["boxes"], out["classes"], out["scores"], out["counts"]
        PredBoxes, PredClasses, ConfScores = [], [], []
        m = max(int(max_boxes), 0)
        for i in range(int(bs)):
            n = int(counts[i])
            b, c, s = boxes[i, :n], classes[i, :n], scores[i, :n]
            if n and (rel_thresh or inc or dup):
                sel = _host_stages(b, c, s, rel_thresh, inc, dup)
                b, c, s = b[sel], c[sel], s[sel]
            PredBoxes.append(list(b[:m]))
            PredClasses.append(list(c[:m]))
            ConfScores.append(list(s[:m]))
        return PredBoxes, PredClasses, ConfScores

    from_logits = False   # set True when the model's class head returns logits (see vision.SSD_loss(from_logits=True))

    def flatten_levels(self, reg_levels, clas_levels):
        """The reference's layout ops (retinanet.py:215-217, :258, :286-295; Vision.py:1467-1468) with torch's kernels:
        per-level NCHW conv outputs -> ([B,A,4], [B,A,C] probabilities).  Not used by the predictor itself (which reads
        the level tensors directly, rn_postproc_levels); kept for callers that want the flat tensors."""
        K = int(reg_levels[0].shape[1]) // 4
        n = int(clas_levels[0].shape[1]) // K
        reg = torch.cat([x.permute(0, 2, 3, 1).contiguous().view(x.shape[0], -1, 4) for x in reg_levels], dim=1)
        clas = torch.cat([x.permute(0, 2, 3, 1).contiguous().view(x.shape[0], -1, n) for x in clas_levels], dim=1)
        return reg, (torch.sigmoid(clas) if self.from_logits else clas)

    def predict_arrays(self, height, width, reg, clas, anchors, thresh=0.05, max_overlap=0.5, top_k=1000,
                       max_boxes=20, full=False):
        """The array form of __call__: dict of host numpy arrays boxes [B,K,4] f32, classes [B,K] i64,
        scores [B,K] f32, anchor_idx [B,K] i32 (the NMS keep indices), counts [B], n_candidates [B]."""
        dev_out = self.predict_device(height, width, reg, clas, anchors, thresh, max_overlap, top_k, max_boxes, full)
        B = int((clas[0] if isinstance(clas, (list, tuple)) else clas).shape[0])
        if dev_out is None:
            return dict(boxes=np.zeros((B, 0, 4), np.float32), classes=np.zeros((B, 0), np.int64),
                        scores=np.zeros((B, 0), np.float32), anchor_idx=np.zeros((B, 0), np.int32),
                        counts=np.zeros(B, np.int32), n_candidates=np.zeros(B, np.int32))
        buf, K = dev_out
        host = buf.cpu().numpy()   # the one device->host copy (and synchronisation) of the call
        o_box, o_cls, o_sc, o_idx, o_cnt, o_cand, total = self._layout(B, K)
        return dict(boxes=host[o_box:o_cls].view(np.float32).reshape(B, K, 4),
                    classes=host[o_cls:o_sc].view(np.int64).reshape(B, K),
                    scores=host[o_sc:o_idx].view(np.float32).reshape(B, K),
                    anchor_idx=host[o_idx:o_cnt].view(np.int32).reshape(B, K),
                    counts=host[o_cnt:o_cand].view(np.int32), n_candidates=host[o_cand:total].view(np.int32))

    @staticmethod
    def _layout(B, K):
        # boxes | classes | scores | anchor_idx | counts | n_candidates  (byte offsets into one buffer)
        return 0, 16 * B * K, 24 * B * K, 28 * B * K, 32 * B * K, 32 * B * K + 4 * B, 32 * B * K + 8 * B

    def predict_device(self, height, width, reg, clas, anchors, thresh=0.05, max_overlap=0.5, top_k=1000,
                       max_boxes=20, full=False):
        """Launches rn_postproc on the current stream and returns (uint8 device buffer, K) without any
        host synchronisation (None when nothing can be returned); predict_arrays() decodes the buffer."""
        if isinstance(clas, (list, tuple)):
            return self._predict_device_levels(height, width, reg, clas, anchors, thresh, max_overlap, top_k, max_boxes, full)
        _lib.require_cuda(clas, "clas", torch.float32)
        _lib.require_cuda(reg, "reg", torch.float32)
        clas, reg = clas.detach().contiguous(), reg.detach().contiguous()
        B, A, Cn = (int(v) for v in clas.shape)
        if reg.shape != (B, A, 4) or anchors.shape != (A, 4):
            raise ValueError("expected reg [B,A,4], clas [B,A,C], anchors [A,4]")
        top_k = int(top_k)
        dev = clas.device
        if top_k < 1 or (not full and int(max_boxes) < 1) or B == 0:
            return None
        if top_k > _lib.MAX_TOP_K:
            raise ValueError("top_k > %d is not supported" % _lib.MAX_TOP_K)
        K = top_k if full else min(int(max_boxes), top_k)
        lib = _lib.load()
        H, W, base, Kc, table, _ = anchor_args(anchors, height, width)
        o_box, o_cls, o_sc, o_idx, o_cnt, o_cand, total = self._layout(B, K)
        with torch.cuda.device(dev):
            buf = torch.empty(total, dtype=torch.uint8, device=dev)   # one buffer => one device->host copy
            ws = _ws.get(lib.rn_postproc_workspace_bytes(B, A, top_k), dev)
            p = buf.data_ptr()
            _lib.check(lib.rn_postproc(
                _lib.ptr(clas), _lib.ptr(reg), B, A, Cn, height, width,
                base, Kc, table, self._mean.ctypes.data_as(_lib._hf32p), self._std.ctypes.data_as(_lib._hf32p),
                float(thresh), float(max_overlap), top_k, K,
                C.c_void_p(p + o_box), C.c_void_p(p + o_cls), C.c_void_p(p + o_sc), C.c_void_p(p + o_idx),
                C.c_void_p(p + o_cnt), C.c_void_p(p + o_cand), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return buf, K

    def _predict_device_levels(self, height, width, reg_levels, clas_levels, anchors, thresh, max_overlap, top_k, max_boxes,
                               full):
        """predict_device for a model that hands over the heads' NCHW conv outputs per pyramid level (reg_l [B, K*4, gh,
        gw], clas_l [B, K*C, gh, gw]; logits when self.from_logits): rn_postproc_levels reads them as they are -- no
        sigmoid / permute / view / cat pass (retinanet.py:215-217, :258, :286-295; Vision.py:1467-1468)."""
        H, W, base, Kc, table, A = anchor_args(anchors, height, width)
        if table is not None:
            raise ValueError("level tensors need anchors from this package's AnchorGenerator (geometry tag)")
        if len(reg_levels) != _lib.NUM_LEVELS or len(clas_levels) != _lib.NUM_LEVELS:
            raise ValueError("expected %d level tensors (P3..P7)" % _lib.NUM_LEVELS)
        if int(clas_levels[0].shape[1]) % Kc:
            raise ValueError("class head channels must be a multiple of the %d anchors per cell" % Kc)
        B, Cn = int(clas_levels[0].shape[0]), int(clas_levels[0].shape[1]) // Kc
        clas_levels = [t.detach() for t in clas_levels]
        reg_levels = [t.detach() for t in reg_levels]
        for l, (c, r) in enumerate(zip(clas_levels, reg_levels)):
            gh, gw = -(-H // (8 << l)), -(-W // (8 << l))
            _lib.require_cuda(c, "clas level", torch.float32)
            _lib.require_cuda(r, "reg level", torch.float32)
            if tuple(c.shape) != (B, Kc * Cn, gh, gw) or tuple(r.shape) != (B, Kc * 4, gh, gw) or not c.is_contiguous() \
                    or not r.is_contiguous():
                raise ValueError("level %d: expected contiguous NCHW clas %s and reg %s" % (l, (B, Kc * Cn, gh, gw), (B, Kc * 4, gh, gw)))
        top_k = int(top_k)
        dev = clas_levels[0].device
        if top_k < 1 or (not full and int(max_boxes) < 1) or B == 0:
            return None
        if top_k > _lib.MAX_TOP_K:
            raise ValueError("top_k > %d is not supported" % _lib.MAX_TOP_K)
        K = top_k if full else min(int(max_boxes), top_k)
        lib = _lib.load()
        o_box, o_cls, o_sc, o_idx, o_cnt, o_cand, total = self._layout(B, K)
        with torch.cuda.device(dev):
            buf = torch.empty(total, dtype=torch.uint8, device=dev)
            ws = _ws.get(lib.rn_postproc_workspace_bytes(B, A, top_k), dev)
            p = buf.data_ptr()
            _lib.check(lib.rn_postproc_levels(
                _lib.ptr_array(clas_levels), _lib.ptr_array(reg_levels), int(bool(self.from_logits)), B, Cn, H, W, base, Kc,
                self._mean.ctypes.data_as(_lib._hf32p), self._std.ctypes.data_as(_lib._hf32p),
                float(thresh), float(max_overlap), top_k, K,
                C.c_void_p(p + o_box), C.c_void_p(p + o_cls), C.c_void_p(p + o_sc), C.c_void_p(p + o_idx),
                C.c_void_p(p + o_cnt), C.c_void_p(p + o_cand), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return buf, K
