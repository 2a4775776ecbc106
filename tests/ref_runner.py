"""Runs the UNMODIFIED reference (imported via oracle/ref_shim.py from /root/reference, or from the staged copy
oracle/_ref on the GPU box) on numpy / torch inputs.  `device="cpu"`: the build container (golden fixtures,
tests/test_oracle_vs_reference.py); `device="cuda:0"`: the reference running ON CUDA, its intended mode and the
primary parity oracle of SURVEY.md section 8c (tests/test_reference_cuda.py)."""
import numpy as np
import torch

from oracle import ref_shim


def available():
    return ref_shim.available()


def _t(x, device):
    return torch.as_tensor(x).to(device)


def anchors(H, W, device="cpu"):
    rn, _ = ref_shim.load()
    return rn.AnchorGenerator()(torch.zeros(1, 3, H, W, device=device)).cpu().numpy()


def assign(anchors_, boxes, cats, device="cpu"):
    """match_anchors_objects for one image (padding already stripped by the caller)."""
    _, vis = ref_shim.load()
    pos, neg, matches = vis.match_anchors_objects(_t(boxes, device).view(-1, 4), _t(anchors_, device))
    return pos.cpu().numpy(), neg.cpu().numpy(), matches.cpu().numpy()


def loss(anchors_, clas, reg, gt_boxes, gt_cats, beta=0.5, alpha=0.25, gamma=2.0, from_logits=False, device="cpu"):
    """SSD_loss forward + autograd backward. Returns dict(out3, dclas, dreg).  from_logits: `clas` holds
    logits and the classification head's nn.Sigmoid (retinanet.py:258,286) is applied before the loss."""
    _, vis = ref_shim.load()
    an = _t(anchors_, device)
    cl = _t(clas, device).clone().requires_grad_(True)
    rg = _t(reg, device).clone().requires_grad_(True)
    f = vis.SSD_loss(beta=beta, alpha=alpha, gamma=gamma)
    out = f([an, rg, torch.nn.Sigmoid()(cl) if from_logits else cl], [_t(gt_boxes, device), _t(gt_cats, device)])
    out.backward()
    out3 = np.array([out.item(), float(f.reg_loss), float(f.clas_loss)], dtype=np.float32)
    dreg = rg.grad.cpu().numpy() if rg.grad is not None else np.zeros_like(np.asarray(reg))
    return dict(out3=out3, dclas=cl.grad.cpu().numpy(), dreg=dreg)


def postproc(clas, reg, anchors_, H, W, thresh=0.05, max_overlap=0.5, top_k=1000, max_boxes=20,
             rel_thresh=None, dup=None, inc=None, device="cpu"):
    """BBoxPredictor.__call__ -> (boxes, classes, scores) lists per image."""
    rn, _ = ref_shim.load()
    bp = rn.BBoxPredictor()
    B = len(clas)
    with torch.no_grad():
        out = bp(torch.zeros(B, 3, H, W, device=device), _t(reg, device), _t(clas, device), _t(anchors_, device),
                 thresh, max_overlap, rel_thresh, top_k, max_boxes, dup, inc)
    return out


def nms(boxes, classes, scores, device="cpu", **kw):
    rn, _ = ref_shim.load()
    return rn.nms(_t(boxes, device), _t(classes, device), _t(scores, device), **kw)
