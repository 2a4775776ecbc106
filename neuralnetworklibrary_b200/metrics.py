"""Detection metrics of the reference's Applications/Vision.py section 6.3, backed by libretina_sm100.so:

    mAP   Vision.py:1749-1800 (+ mAP1 :1696-1747)  -> rn_map_match for the IoU matching of every image, category
                                                      and threshold in one launch; the precision/recall integration
                                                      (sort + cumulative sums over <= #predictions values) in NumPy
                                                      float64 with the reference's exact expressions.
SURVEY.md section 8f row 4.  The reference spends its time here in a Python triple loop
(categories x thresholds x images) with a tensor round trip per image; the bytes involved are tiny.
"""
import numpy as np
import torch

from . import _lib

COCO_thresholds = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]   # reference Vision.py:48
Pascal_thresholds = [0.5]                                                    # reference Vision.py:47


def match_flags(predictions, targets, thresholds, device=None):
    """is_correct [T, NP] uint8 (host) for the concatenated predictions of all images, plus the flat arrays
    (pred_cls [NP], pred_scores [NP], targ_cls [NT]) the integration needs."""
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
        return np.zeros((T, 0), np.uint8), pc_a, ps_a, tc_a
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
        return flags.cpu().numpy(), pc_a, ps_a, tc_a


def average_precision(scores, is_correct, ntrue):
    """The integration half of mAP1 (reference Vision.py:1729-1747) for one (category, threshold)."""
    order = np.lexsort((is_correct, scores))[::-1]          # sorted(zip(Scores, IsCorrect), reverse=True)
    IsCorrect = np.asarray(is_correct)[order].astype(np.int64)
    L = len(IsCorrect)
    running_total_true_pos = np.cumsum(IsCorrect)
    precision_vals = running_total_true_pos * np.array([1 / n for n in range(1, L + 1)])
    precision_maxes = np.flip(np.maximum.accumulate(np.flip(precision_vals)))
    precision_smoothed = precision_maxes[IsCorrect.nonzero()[0]]
    return np.sum(precision_smoothed) / ntrue


def mAP_table(predictions, targets, C, thresholds=COCO_thresholds):
    """The [len(thresholds), C] table of mAP1 values (Vision.py:1791-1794) the reference averages."""
    flags, pred_cls, pred_scores, targ_cls = match_flags(predictions, targets, thresholds)
    scores_table = np.zeros((len(thresholds), C))
    with np.errstate(divide="ignore", invalid="ignore"):
        for c in range(C):
            sel = np.nonzero(pred_cls == c)[0]
            ntrue = np.float64(np.count_nonzero(targ_cls == c))   # a category without ground truth gives nan, as mAP1 does
            for j in range(len(thresholds)):
                scores_table[j, c] = average_precision(pred_scores[sel], flags[j, sel], ntrue)
    return scores_table


def mAP(predictions, targets, categories, thresholds=COCO_thresholds, verbose=True):
    """Mean average precision over categories and IoU thresholds, same arguments, value and printed report as the
    reference's mAP (Vision.py:1749-1800): predictions[i] = [pred_boxes, pred_classes, conf_scores] (the format
    BBoxPredictor / Learner.predict return), targets[i] = [(box, cat), ...]."""
    C = len(categories)
    mAP_scores = mAP_table(predictions, targets, C, thresholds)
    if verbose:
        for c in range(C):
            for j, thresh in enumerate(thresholds):
                print('cat =', c, ':', categories[c], ' thresh =', thresh)
                print('cat-thresh mAP = ', mAP_scores[j, c])
                print('')
        print('Overall mAP = ', np.mean(mAP_scores))
    return np.mean(mAP_scores)
