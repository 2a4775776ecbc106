"""Runs the UNMODIFIED reference (imported from /root/reference via oracle/ref_shim.py) on numpy /
torch CPU inputs.  Build-container only: used by tests/golden/make_golden.py and
tests/test_oracle_vs_reference.py (skipped where the reference checkout is absent)."""
import numpy as np
import torch

from oracle import ref_shim


def available():
    return ref_shim.available()


def anchors(H, W):
    rn, _ = ref_shim.load()
    return rn.AnchorGenerator()(torch.zeros(1, 3, H, W)).numpy()


def assign(anchors_, boxes, cats):
    """match_anchors_objects for one image (padding already stripped by the caller)."""
    _, vis = ref_shim.load()
    pos, neg, matches = vis.match_anchors_objects(torch.as_tensor(boxes).view(-1, 4), torch.as_tensor(anchors_))
    return pos.numpy(), neg.numpy(), matches.numpy()


def loss(anchors_, clas, reg, gt_boxes, gt_cats, beta=0.5, alpha=0.25, gamma=2.0, from_logits=False):
    """SSD_loss forward + autograd backward. Returns dict(out3, dclas, dreg).  from_logits: `clas` holds
    logits and the classification head's nn.Sigmoid (retinanet.py:258,286) is applied before the loss."""
    _, vis = ref_shim.load()
    an = torch.as_tensor(anchors_)
    cl = torch.as_tensor(clas).clone().requires_grad_(True)
    rg = torch.as_tensor(reg).clone().requires_grad_(True)
    f = vis.SSD_loss(beta=beta, alpha=alpha, gamma=gamma)
    out = f([an, rg, torch.nn.Sigmoid()(cl) if from_logits else cl], [torch.as_tensor(gt_boxes), torch.as_tensor(gt_cats)])
    out.backward()
    out3 = np.array([out.item(), float(f.reg_loss), float(f.clas_loss)], dtype=np.float32)
    dreg = rg.grad.numpy() if rg.grad is not None else np.zeros_like(np.asarray(reg))
    return dict(out3=out3, dclas=cl.grad.numpy(), dreg=dreg)


def postproc(clas, reg, anchors_, H, W, thresh=0.05, max_overlap=0.5, top_k=1000, max_boxes=20,
             rel_thresh=None, dup=None, inc=None):
    """BBoxPredictor.__call__ -> (boxes, classes, scores) lists per image."""
    rn, _ = ref_shim.load()
    bp = rn.BBoxPredictor()
    B = len(clas)
    with torch.no_grad():
        out = bp(torch.zeros(B, 3, H, W), torch.as_tensor(reg), torch.as_tensor(clas), torch.as_tensor(anchors_),
                 thresh, max_overlap, rel_thresh, top_k, max_boxes, dup, inc)
    return out


def nms(boxes, classes, scores, **kw):
    rn, _ = ref_shim.load()
    return rn.nms(torch.as_tensor(boxes), torch.as_tensor(classes), torch.as_tensor(scores), **kw)
