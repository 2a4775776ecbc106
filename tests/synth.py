"""Seeded synthetic inputs for the RetinaNet loss / post-processing path.

Shared by tests/, bench.py and tests/golden/make_golden.py so that every consumer (the CUDA path,
the CPU oracle, the reference) sees identical tensors.  Nothing here touches the GPU unless a
`device` is passed; generation itself always happens on the CPU generator so values do not depend
on the device.
"""
import math

import torch


def num_anchors(H, W, per_cell=9):
    """A = per_cell * sum_l ceil(H/2^l) * ceil(W/2^l), l = 3..7 (reference retinanet.py:488)."""
    return per_cell * sum(((H + 2 ** l - 1) // 2 ** l) * ((W + 2 ** l - 1) // 2 ** l) for l in range(3, 8))


def make_targets(B, M, H, W, C, seed, force_empty_and_full=True, min_side=16.0, max_frac=0.4):
    """Ground truth in the collater's format (reference Vision.py:798-812): BBoxes [B,M,4] fp32
    min-max pixels, Cats [B,M] int64, both padded with -1.  All coordinates are >= 0."""
    g = torch.Generator().manual_seed(seed)
    boxes = torch.full((B, M, 4), -1.0, dtype=torch.float32)
    cats = torch.full((B, M), -1, dtype=torch.int64)
    counts = torch.randint(0, M + 1, (B,), generator=g)
    if force_empty_and_full and B >= 2:
        counts[0] = M
        counts[B - 1] = 0
    hi = max_frac * min(H, W)
    for i in range(B):
        m = int(counts[i])
        if m == 0:
            continue
        w = torch.rand(m, generator=g) * (hi - min_side) + min_side
        h = torch.rand(m, generator=g) * (hi - min_side) + min_side
        x1 = torch.rand(m, generator=g) * (W - w)
        y1 = torch.rand(m, generator=g) * (H - h)
        boxes[i, :m] = torch.stack([x1, y1, x1 + w, y1 + h], dim=1)
        cats[i, :m] = torch.randint(0, C, (m,), generator=g)
    return boxes, cats


def make_train_activations(B, A, C, seed, mu=-4.6, edge_cases=256):
    """`clas` = sigmoid(N(mu,1)) (post-sigmoid probabilities; prior 0.01 like the head init,
    reference Vision.py:1434) with a few entries forced onto / outside the clamp bounds of
    focal_loss_retina (Vision.py:1524); `reg` ~ N(0, 0.5^2)."""
    g = torch.Generator().manual_seed(seed)
    clas = torch.sigmoid(torch.randn(B, A, C, generator=g) + mu)
    reg = torch.randn(B, A, 4, generator=g) * 0.5
    if edge_cases:
        flat = clas.view(-1)
        idx = torch.randint(0, flat.numel(), (edge_cases,), generator=g)
        lo, hi = torch.tensor(1e-4, dtype=torch.float32), torch.tensor(1.0 - 1e-4, dtype=torch.float32)
        vals = torch.stack([lo, hi, lo * 0.5, (hi + 1.0) * 0.5, torch.tensor(0.0), torch.tensor(1.0),
                            torch.nextafter(lo, torch.tensor(0.0)), torch.nextafter(hi, torch.tensor(1.0))])
        flat[idx] = vals[torch.arange(edge_cases) % vals.numel()]
    return clas, reg


def make_infer_activations(B, A, C, seed, anchors=None, mu=-6.0, clusters=20, per_cluster=30):
    """Inference-shaped activations: sparse scores above the 0.05 threshold plus planted clusters
    of overlapping high-score anchors of one class so that NMS really suppresses.  Scores are made
    tie-free per image (the reference's sort is unstable)."""
    g = torch.Generator().manual_seed(seed)
    clas = torch.sigmoid(torch.randn(B, A, C, generator=g) + mu)
    reg = torch.randn(B, A, 4, generator=g) * 0.5
    if anchors is not None and clusters > 0:
        an = torch.as_tensor(anchors, dtype=torch.float32)
        cx, cy = (an[:, 0] + an[:, 2]) * 0.5, (an[:, 1] + an[:, 3]) * 0.5
        for i in range(B):
            seeds = torch.randint(0, A, (clusters,), generator=g)
            for s in seeds.tolist():
                size = (an[s, 2] - an[s, 0]).item()
                d = (cx - cx[s]).abs() + (cy - cy[s]).abs() + (an[:, 2] - an[:, 0] - size).abs()
                near = torch.topk(-d, min(per_cluster, A)).indices
                c = int(torch.randint(0, C, (1,), generator=g))
                clas[i, near, c] = torch.rand(near.numel(), generator=g) * 0.69 + 0.3
                reg[i, near] *= 0.2
    # break score ties per image (max over classes) by nudging duplicates
    for i in range(B):
        top = clas[i].max(dim=1).values
        srt, order = torch.sort(top)
        dup = (srt[1:] == srt[:-1]).nonzero().view(-1)
        tries = 0
        while dup.numel() > 0 and tries < 8:
            rows = order[dup + 1]
            cols = clas[i, rows].argmax(dim=1)
            clas[i, rows, cols] = torch.nextafter(clas[i, rows, cols], torch.tensor(2.0))
            top = clas[i].max(dim=1).values
            srt, order = torch.sort(top)
            dup = (srt[1:] == srt[:-1]).nonzero().view(-1)
            tries += 1
    return clas, reg


# --------------------------------------------------------------------------------------------------
# Comparison helpers shared by tests/, smoke() and bench.py (tolerances: BASELINE.json north_star)
# --------------------------------------------------------------------------------------------------
RTOL = 1e-5


def assert_rel(actual, expected, rtol=RTOL, what=""):
    """Element-wise |a-e| <= rtol*|e| with NO absolute slack, and identical zero patterns."""
    import numpy as np
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    assert np.array_equal(actual == 0, expected == 0), what + ": zero patterns differ"
    nz = expected != 0
    if nz.any():
        err = np.abs(actual[nz].astype(np.float64) - expected[nz]) / np.abs(expected[nz])
        assert err.max() <= rtol, "%s: max rel err %.3g" % (what, err.max())


def assert_dreg_close(actual, expected, rtol=RTOL, what="dreg"):
    """Smooth-L1 gradients.  Below the knee the gradient is 9*(t-p)*g_e, i.e. proportional to a
    DIFFERENCE of the encoded target t (which contains logf) and the prediction; one ulp of logf
    (CUDA's vs the host libm's) therefore moves a near-zero element by up to ~2e-6*g_e no matter how
    small the element is.  So: identical zero pattern, and |a-e| <= rtol*|e| + rtol*g_e, where
    g_e = max|e| of the image is the gradient of any element above the knee."""
    import numpy as np
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape
    assert np.array_equal(actual == 0, expected == 0), what + ": zero patterns differ"
    for i in range(expected.shape[0]):
        scale = np.abs(expected[i]).max()
        err = np.abs(actual[i].astype(np.float64) - expected[i])
        bound = rtol * np.abs(expected[i]) + rtol * scale
        assert (err <= bound).all(), "%s image %d: max excess %.3g (scale %.3g)" % (what, i, (err - bound).max(), scale)


def assert_boxes_close(actual, expected, rtol=RTOL, what="boxes"):
    """Decoded boxes [n,4]: |a-e| <= rtol * (largest |coordinate| of that box).  A coordinate is a
    difference centre -/+ size/2 with size = w*expf(.), so its error scales with the box, not with the
    coordinate itself (a corner near the image origin has no meaningful relative error)."""
    import numpy as np
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape
    if expected.size == 0:
        return
    scale = np.abs(expected).max(axis=-1, keepdims=True)
    err = np.abs(actual.astype(np.float64) - expected)
    assert (err <= rtol * np.maximum(scale, 1e-30)).all(), "%s: max err/scale %.3g" % (what, (err / np.maximum(scale, 1e-30)).max())
