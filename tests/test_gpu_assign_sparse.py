"""The sparse assignment path of rn_assign (one CTA per ground-truth box enumerating that box's candidate anchors) against
the dense kernel of the same library (rn_set_option("assign_dense", 1)) and against the CPU oracle: matches and positive counts must be
identical for every box geometry -- tiny, huge, partly or completely outside the image, degenerate, duplicated, touching
the thresholds' neighbourhood -- because a box whose window missed a single anchor would silently turn it into background."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _assign(anchors, gb, gc, dense, **kw):
    from neuralnetworklibrary_b200 import _lib
    from neuralnetworklibrary_b200.vision import assign_batch
    with _lib.option("assign_dense", 1 if dense else 0):
        m, n, _ = assign_batch(anchors, gb, gc, **kw)
        torch.cuda.synchronize()
    return m.cpu().numpy(), n.cpu().numpy()


def _boxes(rng, B, M, H, W, kind):
    gb = np.full((B, M, 4), -1, np.float32)
    gc = np.full((B, M), -1, np.int64)
    for b in range(B):
        n = int(rng.integers(0, M + 1)) if b else M
        rows = rng.permutation(M)[:n]          # padding rows in the middle
        for r in rows:
            if kind == "tiny":
                w, h = rng.uniform(1, 24, 2)
            elif kind == "huge":
                w, h = rng.uniform(0.5, 1.3) * W, rng.uniform(0.5, 1.3) * H
            elif kind == "thin":
                w, h = (rng.uniform(2, 10), rng.uniform(0.3 * H, 0.9 * H)) if rng.random() < 0.5 else (rng.uniform(0.3 * W, 0.9 * W), rng.uniform(2, 10))
            elif kind == "anchor_like":        # boxes that coincide with / sit close to anchors: IoU near 1, 0.5 and 0.4
                s = 32 * 2 ** rng.integers(0, 5) * 2 ** (rng.integers(0, 3) / 3)
                ratio = [0.5, 1, 2][rng.integers(0, 3)]
                w, h = s * np.sqrt(ratio) * rng.uniform(0.7, 1.4), s / np.sqrt(ratio) * rng.uniform(0.7, 1.4)
            else:
                w, h = rng.uniform(8, 0.6 * W), rng.uniform(8, 0.6 * H)
            x1, y1 = rng.uniform(-0.2 * W, W), rng.uniform(-0.2 * H, H)
            gb[b, r] = [x1, y1, x1 + w, y1 + h]
            gc[b, r] = rng.integers(0, 20)
        if n >= 3:
            gb[b, rows[1]] = gb[b, rows[0]]                    # duplicated box: first index wins
            gb[b, rows[2], 2:] = gb[b, rows[2], :2]            # zero-area box
    return torch.from_numpy(gb), torch.from_numpy(gc)


@pytest.mark.parametrize("kind", ["mixed", "tiny", "huge", "thin", "anchor_like"])
@pytest.mark.parametrize("H,W", [(512, 512), (800, 1344), (100, 167), (33, 47)])
def test_sparse_equals_dense_and_oracle(kind, H, W):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    rng = np.random.default_rng(hash((kind, H, W)) % (2 ** 32))
    B, M = 4, 12
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    gb, gc = _boxes(rng, B, M, H, W, kind)
    ms, ns = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=False)
    md, nd = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=True)
    assert np.array_equal(ms, md) and np.array_equal(ns, nd)
    an = orc.anchors(H, W)
    for b in range(B):
        v = gc[b] >= 0
        mo, _, _ = orc.assign(an, gb[b][v].numpy(), gc[b][v].numpy())
        assert np.array_equal(ms[b], mo)
        assert ns[b] == int((mo >= 0).sum())


@pytest.mark.parametrize("pos,neg", [(0.5, 0.4), (0.7, 0.3), (0.45, 0.45), (0.9, 0.2)])
def test_sparse_other_thresholds_and_anchor_sets(pos, neg):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    rng = np.random.default_rng(int(pos * 100) * 100 + int(neg * 100))
    H, W, B, M = 320, 416, 3, 9
    for ratios, scales in (([0.5, 1, 2], [1, 2 ** (1 / 3), 2 ** (2 / 3)]), ([1.0], [1.0]), ([0.3, 3.0], [0.8, 1.7])):
        anchors = AnchorGenerator(ratios, scales)(torch.zeros(1, 3, H, W, device=dev()))
        gb, gc = _boxes(rng, B, M, H, W, "anchor_like")
        ms, ns = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=False, pos_thresh=pos, neg_thresh=neg)
        md, nd = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=True, pos_thresh=pos, neg_thresh=neg)
        assert np.array_equal(ms, md) and np.array_equal(ns, nd)
        assert (ms != -1).sum() > 0   # the case is not vacuous


def test_sparse_boxes_far_outside_and_non_finite():
    """Boxes far outside the image (1e12 px), with infinite or NaN coordinates: no window, no overflow, same result as the
    dense kernel (they overlap no anchor, or compare false everywhere)."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    H, W = 256, 320
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    gb = torch.tensor([[[40., 50., 140., 170.], [1e12, 1e12, 1e12 + 100, 1e12 + 80], [-1e12, 10., -1e12 + 50, 90.],
                        [10., 10., float("inf"), 60.], [float("nan"), 5., 50., 60.], [100., 90., 180., 200.]]])
    gc = torch.tensor([[1, 2, 3, 4, 5, 6]])
    ms, ns = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=False)
    md, nd = _assign(anchors, gb.to(dev()), gc.to(dev()), dense=True)
    assert np.array_equal(ms, md) and np.array_equal(ns, nd) and ns[0] > 0
