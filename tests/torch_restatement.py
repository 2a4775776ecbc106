"""TEST INFRASTRUCTURE: the reference's hot path restated as the same sequence of torch tensor operations, so that
it can run ON THE GPU BOX (where /root/reference does not exist) with torch's own CUDA kernels as a second checker
next to the C oracle.  SURVEY.md section 8c names "the reference running on CUDA" as the primary parity oracle: its
ATen kernels call the device libm (expf / logf) our kernels call, so decoded boxes and everything derived from them
(keep sets) can be compared bit for bit, which the host-libm oracle cannot offer.

Pinned in the build container against the UNMODIFIED reference on CPU (tests/test_torch_restatement.py: assignments,
decoded candidate lists and keep lists bit-exact, loss and autograd gradients bit-exact as well since the operation
sequence per image is the same).  Nothing in the product imports this module.

Each function cites the reference lines it follows (paths relative to the reference checkout)."""
import numpy as np
import torch


def jaccard(b1, b2):
    """Applications/Vision.py:234-256: pairwise IoU [n,4] x [m,4] -> [n,m], every op a separate fp32 kernel."""
    b1, b2 = b1.float(), b2.float()
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    p, q = b1.unsqueeze(1), b2.unsqueeze(0)
    iw = (torch.min(p[:, :, 2], q[:, :, 2]) - torch.max(p[:, :, 0], q[:, :, 0])).clamp(min=0)
    ih = (torch.min(p[:, :, 3], q[:, :, 3]) - torch.max(p[:, :, 1], q[:, :, 1])).clamp(min=0)
    inter = iw * ih
    union = a1.unsqueeze(dim=1) + a2.unsqueeze(dim=0) - inter
    return inter / union


def match(objects, anchors, pos_thresh=0.5, neg_thresh=0.4):
    """Applications/Vision.py:1474-1511 in this library's encoding: matches [A] int64 with the object index for
    positives, -1 background (max IoU < neg_thresh), -2 ignored; both compares strict."""
    A = len(anchors)
    if len(objects) == 0:
        return torch.full((A,), -1, dtype=torch.int64, device=anchors.device)
    best, arg = torch.max(jaccard(objects, anchors), dim=0)   # first maximal index
    out = torch.full((A,), -2, dtype=torch.int64, device=anchors.device)
    out[best < neg_thresh] = -1
    pos = best > pos_thresh
    out[pos] = arg[pos]
    return out


def focal(pred, target, alpha, gamma):
    """Applications/Vision.py:1513-1530."""
    p = pred.clamp(1e-4, 1.0 - 1e-4)
    pt = p * target + (1 - p) * (1 - target)
    wa = alpha * target + (1 - alpha) * (1 - target)
    w = wa * (1 - pt).pow(gamma)
    losses = -w * (target * torch.log(p) + (1 - target) * torch.log(1 - p))
    return losses.sum() / target.sum().clamp(min=1)


def smooth_l1(anchs, pred_shift, target):
    """Applications/Vision.py:1532-1566."""
    aw, ah = anchs[:, 2] - anchs[:, 0], anchs[:, 3] - anchs[:, 1]
    acx, acy = anchs[:, 0] + 0.5 * aw, anchs[:, 1] + 0.5 * ah
    tw, th = target[:, 2] - target[:, 0], target[:, 3] - target[:, 1]
    tcx, tcy = target[:, 0] + 0.5 * tw, target[:, 1] + 0.5 * th
    tw, th = tw.clamp(min=1), th.clamp(min=1)
    true_shift = torch.stack(((tcx - acx) / aw, (tcy - acy) / ah, torch.log(tw / aw), torch.log(th / ah))).t()
    true_shift = true_shift / torch.tensor([[0.1, 0.1, 0.2, 0.2]], dtype=torch.float32, device=anchs.device)
    diff = torch.abs(true_shift - pred_shift)
    losses = 0.5 * 9 * pow(diff, 2) * (diff < 1 / 9).float() + (diff - 0.5 / 9) * (diff >= 1 / 9).float()
    return losses.mean()


def ssd_loss(anchors, reg, clas, BBoxes, Cats, beta=0.5, alpha=0.25, gamma=2.0):
    """Applications/Vision.py:1568-1605 (ssd1) + :1620-1644 (SSD_loss.__call__).  Returns (loss, reg_loss, clas_loss,
    matches [bs,A]); differentiable w.r.t. reg and clas.  Padding rows are those with Cats < 0 (the reference filters
    negative ELEMENTS, identical for non-negative coordinates)."""
    bs, A, C = clas.shape
    dev = clas.device
    reg_total = torch.zeros((), dtype=torch.float32, device=dev)
    clas_total = torch.zeros((), dtype=torch.float32, device=dev)
    all_matches = []
    for i in range(bs):
        valid = Cats[i] >= 0
        boxes, cats = BBoxes[i][valid].view(-1, 4), Cats[i][valid]
        m = match(boxes, anchors)
        all_matches.append(m)
        pos_idx = (m >= 0).nonzero().view(-1)
        neg_idx = (m == -1).nonzero().view(-1)
        well_defined = torch.cat([pos_idx, neg_idx])
        target = torch.zeros(A, C, device=dev)
        if len(pos_idx) > 0:
            target[pos_idx, cats[m[pos_idx]]] = 1
        clas_total = clas_total + focal(clas[i][well_defined], target[well_defined], alpha, gamma)
        if len(pos_idx) > 0:
            reg_total = reg_total + smooth_l1(anchors[pos_idx], reg[i][pos_idx], boxes[m[pos_idx]])
    reg_loss, clas_loss = reg_total / bs, clas_total / bs
    return (1 - beta) * reg_loss + beta * clas_loss, reg_loss, clas_loss, torch.stack(all_matches)


def _jaccard_np(b1, b2):
    """Applications/VisionModels/retinanet.py:500-521 (NumPy float32)."""
    iw = (np.minimum(b1[:, None, 2], b2[None, :, 2]) - np.maximum(b1[:, None, 0], b2[None, :, 0])).clip(0, None)
    ih = (np.minimum(b1[:, None, 3], b2[None, :, 3]) - np.maximum(b1[:, None, 1], b2[None, :, 1])).clip(0, None)
    inter = iw * ih
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    return inter / (a1[:, None] + a2[None, :] - inter)


def greedy_nms(boxes, classes, scores, max_overlap=0.5, top_k=1000, max_boxes=20):
    """Applications/VisionModels/retinanet.py:570-602 + :702-704 with rel_thresh/inc/dup = None: device sort, top_k,
    host greedy loop.  Returns indices into the inputs (score-descending)."""
    if len(boxes) == 0:
        return np.zeros(0, np.int64)
    _, order = scores.sort(descending=True)
    order = order[:top_k]
    b = boxes[order].detach().cpu().numpy()
    c = classes[order].detach().cpu().numpy()
    alive = np.arange(len(b))
    keep = []
    while len(alive) > 0:
        jac = _jaccard_np(b[alive[:1]], b[alive])[0]
        drop = (jac > max_overlap) & (c[alive] == c[alive[0]])
        keep.append(alive[0])
        alive = alive[~drop]
    keep = np.array(keep[:max_boxes], dtype=np.int64)
    return order.cpu().numpy()[keep]


def predict(H, W, reg, clas, anchors, thresh=0.05, max_overlap=0.5, top_k=1000, max_boxes=20,
            mean=(0., 0., 0., 0.), std=(0.1, 0.1, 0.2, 0.2)):
    """Applications/VisionModels/retinanet.py:732-812 (BBoxPredictor.__call__).  Per image: dict(anchor_idx, boxes,
    classes, scores) of the kept detections plus n_candidates (boxes handed to nms)."""
    dev = clas.device
    mean = torch.tensor(mean, dtype=torch.float32, device=dev)
    std = torch.tensor(std, dtype=torch.float32, device=dev)
    aw, ah = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
    acx, acy = anchors[:, 0] + 0.5 * aw, anchors[:, 1] + 0.5 * ah
    out = []
    for i in range(len(clas)):
        scores, classes = clas[i].max(dim=1)
        idx = (scores > thresh).nonzero().view(-1)
        r = reg[i][idx]
        w, h, cx, cy = aw[idx], ah[idx], acx[idx], acy[idx]
        dx, dy = r[:, 0] * std[0] + mean[0], r[:, 1] * std[1] + mean[1]
        dw, dh = r[:, 2] * std[2] + mean[2], r[:, 3] * std[3] + mean[3]
        pcx, pcy = cx + w * dx, cy + h * dy
        pw, ph = w * torch.exp(dw), h * torch.exp(dh)
        boxes = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=1)
        boxes[:, 0] = torch.clamp(boxes[:, 0], min=0)
        boxes[:, 1] = torch.clamp(boxes[:, 1], min=0)
        boxes[:, 2] = torch.clamp(boxes[:, 2], max=W)
        boxes[:, 3] = torch.clamp(boxes[:, 3], max=H)
        good = (((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)).nonzero().view(-1)
        boxes, idx = boxes[good], idx[good]
        keep = greedy_nms(boxes, classes[idx], scores[idx], max_overlap, top_k, max_boxes)
        out.append(dict(n_candidates=int(len(idx)), anchor_idx=idx.cpu().numpy()[keep],
                        boxes=boxes.detach().cpu().numpy()[keep], classes=classes[idx].cpu().numpy()[keep],
                        scores=scores[idx].detach().cpu().numpy()[keep]))
    return out
