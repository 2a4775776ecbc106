"""Detection metrics of the reference's Applications/Vision.py section 6.3, backed by libretina_sm100.so:

    mAP   Vision.py:1749-1800 (+ mAP1 :1696-1747)  -> rn_map_match for the IoU matching of every image, category
                                                      and threshold in one launch, rn_map_ap for the precision/recall
                                                      integration of every (threshold, category) pair in a second one;
                                                      one [T, C] float64 table comes back to the host.
SURVEY.md section 8f row 4.  The reference spends its time here in a Python triple loop
(categories x thresholds x images) with a tensor round trip per image; the bytes involved are tiny.
"""
import numpy as np
import torch

from . import _lib

COCO_thresholds = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]   # reference Vision.py:48
Pascal_thresholds = [0.5]                                                    # reference Vision.py:47


def match_flags(predictions, targets, thresholds, device=None, to_host=True):
    """is_correct [T, NP] uint8 for the concatenated predictions of all images (a host array, or the device tensor with
    to_host=False), plus the flat host arrays (pred_cls [NP], pred_scores [NP], targ_cls [NT]) of the inputs."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    pb, pc, ps, poff = [], [], [], [0]
    tb, tc, ti = [], [], []
    for i, (pred, targ) in enumerate(zip(predictions, targets)):
        boxes, classes, scores = pred
        for j in range(len(boxes)):
            pb.append(np.asarray(boxes[j], dtype=np.float32))
            pc.append(int(classes[j]))
            ps.append(np.float32(scores[j]))
        poff.append(len(pb))
        for b, c in targ:
            tb.append(np.asarray(b, dtype=np.float32))   # TEN(np.array(targs[i])) rounds to float32
            tc.append(int(c))
            ti.append(i)
    NP, NT, T = len(pb), len(tb), len(thresholds)
    pc_a, ps_a, tc_a = np.array(pc, np.int32), np.array(ps, np.float32), np.array(tc, np.int32)
    if NP == 0:
        empty = np.zeros((T, 0), np.uint8)
        return (empty if to_host else torch.zeros((T, 1), dtype=torch.uint8, device=device)), pc_a, ps_a, tc_a
    with torch.cuda.device(device):
        d_pb = torch.from_numpy(np.stack(pb).astype(np.float32)).to(device)
        d_pc = torch.from_numpy(pc_a).to(device)
        d_po = torch.from_numpy(np.array(poff, np.int32)).to(device)
        d_tb = torch.from_numpy(np.stack(tb).astype(np.float32) if NT else np.zeros((1, 4), np.float32)).to(device)
        d_tc = torch.from_numpy(tc_a if NT else np.zeros(1, np.int32)).to(device)
        d_ti = torch.from_numpy(np.array(ti, np.int32) if NT else np.zeros(1, np.int32)).to(device)
        d_th = torch.tensor([float(t) for t in thresholds], dtype=torch.float32, device=device)  # compared in fp32
        flags = torch.empty((T, NP), dtype=torch.uint8, device=device)
        _lib.check(lib.rn_map_match(_lib.ptr(d_pb), _lib.ptr(d_pc), _lib.ptr(d_po), _lib.ptr(d_tb), _lib.ptr(d_tc),
                                    _lib.ptr(d_ti), NT, NP, _lib.ptr(d_th), T, _lib.ptr(flags), _lib.stream_ptr(device)))
        return (flags.cpu().numpy() if to_host else flags), pc_a, ps_a, tc_a


_ws = _lib.Workspace()


def mAP_table(predictions, targets, C, thresholds=COCO_thresholds, device=None):
    """The [len(thresholds), C] float64 table of per-(threshold, category) average precisions the reference averages
    (Vision.py:1791-1794), computed on the device: rn_map_match (IoU matching) + rn_map_ap (sort by score, running counts,
    right-to-left running maximum, pairwise sum -- Vision.py:1729-1747) and ONE device->host copy of the table.  A category
    without ground truth gives nan, as mAP1 does."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    d_flags, pred_cls, pred_scores, targ_cls = match_flags(predictions, targets, thresholds, device, to_host=False)
    NP, T = len(pred_cls), len(thresholds)
    if np.any((pred_cls < 0) | (pred_cls >= C)) or np.any((targ_cls < 0) | (targ_cls >= C)):
        raise IndexError("category outside [0, %d)" % C)   # the reference indexes its per-category lists (Vision.py:1780-1786)
    perm = np.argsort(pred_cls, kind="stable").astype(np.int32)             # predictions grouped by category (layout only)
    cls_off = np.zeros(C + 1, np.int32)
    cls_off[1:] = np.cumsum(np.bincount(pred_cls, minlength=C))
    ntrue = np.bincount(targ_cls, minlength=C).astype(np.int32)             # sum(len(targs[i])) per category, Vision.py:1733
    with torch.cuda.device(device):
        d_perm = torch.from_numpy(perm if NP else np.zeros(1, np.int32)).to(device)
        d_off = torch.from_numpy(cls_off).to(device)
        d_nt = torch.from_numpy(ntrue).to(device)
        d_sc = torch.from_numpy(pred_scores if NP else np.zeros(1, np.float32)).to(device)
        table = torch.empty((T, C), dtype=torch.float64, device=device)
        ws = _ws.get(lib.rn_map_ap_workspace_bytes(NP, T), device)
        _lib.check(lib.rn_map_ap(_lib.ptr(d_sc), _lib.ptr(d_perm), _lib.ptr(d_off), _lib.ptr(d_flags), _lib.ptr(d_nt), NP, C, T,
                                 _lib.ptr(table), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device)))
        return table.cpu().numpy()


def mAP(predictions, targets, categories, thresholds=COCO_thresholds, verbose=True):
    """Mean average precision over categories and IoU thresholds, same arguments, value and printed report as the
    reference's mAP (Vision.py:1749-1800): predictions[i] = [pred_boxes, pred_classes, conf_scores] (the format
    BBoxPredictor / Learner.predict return), targets[i] = [(box, cat), ...]."""
    table = mAP_table(predictions, targets, len(categories), thresholds)
    overall = np.mean(table)
    if verbose:
        print(_report(table, categories, thresholds, overall))
    return overall


def _report(table, categories, thresholds, overall):
    """The text the reference prints while it fills its table (Vision.py:1793-1798): one block per (category, threshold)
    in category-major order, then the overall mean."""
    blocks = ["cat = {} : {}  thresh = {}\ncat-thresh mAP =  {}\n".format(c, categories[c], th, table[j, c])
              for c in range(len(categories)) for j, th in enumerate(thresholds)]
    return "\n".join(blocks + ["Overall mAP =  {}".format(overall)])
