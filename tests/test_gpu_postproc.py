"""GPU parity: post-processing (rn_postproc / rn_nms through BBoxPredictor and nms) against the CPU
oracle and the reference-generated golden fixtures.  Keep indices, classes, scores and counts bit-exact;
box coordinates rtol 1e-5 (the decode uses expf)."""
import os

import numpy as np
import pytest
import torch

from tests import synth as syn
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def dev():
    return torch.device("cuda:0")


def make_anchors(H, W, table_mode=False):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    a = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    return a.clone() if table_mode else a


def compare(out, po):
    assert np.array_equal(out["counts"], po["counts"])
    assert np.array_equal(out["n_candidates"], po["n_candidates"])
    for i, n in enumerate(po["counts"]):
        assert np.array_equal(out["anchor_idx"][i, :n], po["anchor_idx"][i, :n]), "keep indices differ (image %d)" % i
        assert np.array_equal(out["classes"][i, :n], po["classes"][i, :n])
        assert np.array_equal(out["scores"][i, :n], po["scores"][i, :n])
        syn.assert_boxes_close(out["boxes"][i, :n], po["boxes"][i, :n])


POST_CASES = [  # seed, H, W, C, B, mu, kwargs
    (71, 128, 160, 20, 3, -5.0, {}),
    (72, 128, 160, 80, 2, -6.0, {}),
    (73, 256, 256, 80, 2, -3.0, {}),                                   # > 4096 candidates: radix select
    (74, 256, 256, 20, 2, -2.0, dict(top_k=300, max_boxes=300)),       # everything over threshold
    (75, 96, 96, 7, 3, -4.0, dict(thresh=0.1, max_overlap=0.3)),       # C % 4 != 0
    (76, 128, 128, 12, 2, -4.0, dict(top_k=50, max_boxes=100)),        # max_boxes > top_k
    (77, 64, 64, 80, 4, -7.5, dict(max_boxes=1)),
    (78, 128, 160, 16, 2, -4.0, dict(top_k=4096, max_boxes=4096, thresh=0.01)),
]


@pytest.mark.parametrize("table_mode", [False, True])
@pytest.mark.parametrize("seed,H,W,C,B,mu,kw", POST_CASES)
def test_postproc_vs_oracle(seed, H, W, C, B, mu, kw, table_mode):
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    anchors = make_anchors(H, W, table_mode)
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=mu, clusters=8)
    out = BBoxPredictor().predict_arrays(H, W, reg.to(dev()), clas.to(dev()), anchors, **kw)
    kw2 = dict(kw)
    kw2["max_boxes"] = min(kw.get("max_boxes", 20), kw.get("top_k", 1000))
    po = orc.postproc(clas.numpy(), reg.numpy(), an, H, W, **kw2)
    compare(out, po)


def test_postproc_golden_lists(golden_dir):
    """The list-of-lists return convention of BBoxPredictor.__call__ against the reference's output."""
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    g = np.load(os.path.join(golden_dir, "postproc_small.npz"))
    H, W, B = int(g["H"]), int(g["W"]), int(g["B"])
    anchors = make_anchors(H, W)
    img = torch.zeros(B, 3, H, W, device=dev())
    reg, clas = torch.from_numpy(g["reg"]).to(dev()), torch.from_numpy(g["clas"]).to(dev())
    variants = [("default", ()), ("topk50_max100", (0.05, 0.5, None, 50, 100)),
                ("thr02_ov03_max7", (0.2, 0.3, None, 1000, 7)), ("thr001_max1000", (0.01, 0.5, None, 1000, 1000))]
    for name, args in variants:
        PB, PC, CS = BBoxPredictor()(img, reg, clas, anchors, *args)   # positional, like Learner.py:374-376
        assert len(PB) == len(PC) == len(CS) == B
        for i in range(B):
            n = int(g[name + "_counts"][i])
            assert len(PB[i]) == len(PC[i]) == len(CS[i]) == n
            if n == 0:
                assert PB[i] == [] and PC[i] == [] and CS[i] == []
                continue
            assert isinstance(PB[i][0], np.ndarray) and PB[i][0].dtype == np.float32 and PB[i][0].shape == (4,)
            assert isinstance(PC[i][0], np.int64) and isinstance(CS[i][0], np.float32)
            assert np.array_equal(np.array(PC[i]), g[name + "_classes"][i, :n])
            assert np.array_equal(np.array(CS[i]), g[name + "_scores"][i, :n])
            syn.assert_boxes_close(np.stack(PB[i]), g[name + "_boxes"][i, :n])
            _ = PB[i][0] * 0.5   # callers rescale boxes with list_mult (Learner.py:377-381)


def test_postproc_empty_image():
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    H, W, C, B = 64, 64, 20, 2
    anchors = make_anchors(H, W)
    A = anchors.shape[0]
    clas = torch.full((B, A, C), 0.01, device=dev())
    clas[1, 5, 3] = 0.9
    reg = torch.zeros(B, A, 4, device=dev())
    PB, PC, CS = BBoxPredictor()(torch.zeros(B, 3, H, W, device=dev()), reg, clas, anchors)
    assert PB[0] == [] and PC[0] == [] and CS[0] == []
    assert len(PB[1]) == 1 and PC[1][0] == 3 and CS[1][0] == np.float32(0.9)
    # score exactly at the threshold is dropped (strict >, retinanet.py:760)
    clas[1, 5, 3] = 0.05
    PB, _, _ = BBoxPredictor()(torch.zeros(B, 3, H, W, device=dev()), reg, clas, anchors)
    assert PB[1] == []


NMS_VARIANTS = [("all", dict(top_k=3000, max_boxes=100000)), ("default", dict()),
                ("ov07_topk500_max50", dict(max_overlap=0.7, top_k=500, max_boxes=50)),
                ("rel", dict(rel_thresh=[0.3, 0.6], max_boxes=1000)),
                ("inc", dict(inc=[0.9, [1, 3]], max_boxes=1000)),
                ("dup", dict(dup=[0.4, [(0, 1), (1, 0), (2, 3)]], max_boxes=1000)),
                ("rel_inc_dup", dict(rel_thresh=[0.2, 0.5], inc=[0.8, [2]], dup=[0.5, [(0, 1), (3, 4)]],
                                     top_k=2000, max_boxes=60))]


@pytest.mark.parametrize("variant,kw", NMS_VARIANTS)
@pytest.mark.parametrize("on_device", [True, False])
def test_nms_golden(golden_dir, variant, kw, on_device):
    """nms() on caller-provided boxes against the reference's output, incl. the optional host stages."""
    from neuralnetworklibrary_b200.retinanet import nms
    g = np.load(os.path.join(golden_dir, "nms_boxes.npz"))
    boxes, classes, scores = (torch.from_numpy(g[k]) for k in ("boxes", "classes", "scores"))
    if on_device:
        boxes, classes, scores = boxes.to(dev()), classes.to(dev()), scores.to(dev())
    rb, rc, rs = nms(boxes, classes, scores, **kw)
    assert len(rb) == len(g[variant + "_scores"])
    assert np.array_equal(np.array(rs, np.float32), g[variant + "_scores"])
    assert np.array_equal(np.array(rc, np.int64), g[variant + "_classes"])
    assert np.array_equal(np.stack(rb), g[variant + "_boxes"])


def test_nms_empty_and_single():
    from neuralnetworklibrary_b200.retinanet import nms
    assert nms(torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64), torch.zeros(0)) == ([], [], [])
    b, c, s = nms(torch.tensor([[0., 0., 5., 5.]]), torch.tensor([2]), torch.tensor([0.7]))
    assert len(b) == 1 and c[0] == 2 and s[0] == np.float32(0.7)
    # identical boxes, same class -> one survivor; different class -> both
    bx = torch.tensor([[0., 0., 5., 5.], [0., 0., 5., 5.]])
    assert len(nms(bx, torch.tensor([1, 1]), torch.tensor([0.6, 0.7]))[0]) == 1
    assert len(nms(bx, torch.tensor([1, 2]), torch.tensor([0.6, 0.7]))[0]) == 2
    # IoU exactly at max_overlap is kept (strict >, retinanet.py:592)
    bx = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 5.]])
    assert len(nms(bx, torch.tensor([1, 1]), torch.tensor([0.9, 0.8]), max_overlap=0.5)[0]) == 2


@pytest.mark.parametrize("H,W,C,B,seed", [(800, 1344, 80, 2, 1004), (800, 1333, 80, 2, 1007)])
def test_postproc_full_size(H, W, C, B, seed):
    """BASELINE.json COCO post-processing shape (A = 201600, 80 classes) at a batch the oracle finishes
    in seconds, plus size-independent properties of the result."""
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    anchors = make_anchors(H, W)
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=-6.0)
    out = BBoxPredictor().predict_arrays(H, W, reg.to(dev()), clas.to(dev()), anchors, max_boxes=100)
    po = orc.postproc(clas.numpy(), reg.numpy(), an, H, W, max_boxes=100)
    compare(out, po)
    for i in range(B):
        n = out["counts"][i]
        s, b, c = out["scores"][i, :n], out["boxes"][i, :n], out["classes"][i, :n]
        assert (np.diff(s) <= 0).all() and (s > np.float32(0.05)).all()           # sorted, thresholded
        assert (b[:, 0] >= 0).all() and (b[:, 2] <= W).all() and (b[:, 3] <= H).all()
        # idempotence: NMS of the survivors keeps all of them
        keep = orc.nms(b, c, s, max_boxes=1000)
        assert len(keep) == n


def test_nms_batch_ragged_images_match_single_calls():
    """rn_nms_batch (one launch sequence for L images, survivors gathered on the device) against nms() image by image and the
    oracle: empty images, a one-box image, an image with more candidates than top_k, top_k < max_boxes."""
    from neuralnetworklibrary_b200.retinanet import decode_nms_buffer, nms, nms_batch_device
    rng = np.random.RandomState(21)
    sizes = [0, 1, 37, 0, 900, 5, 2600]
    boxes, classes, scores = [], [], []
    for n in sizes:
        xy = rng.uniform(0, 500, (n, 2)); wh = rng.uniform(8, 150, (n, 2))
        boxes.append(np.concatenate([xy, xy + wh], 1).astype(np.float32))
        classes.append(rng.randint(0, 5, n).astype(np.int64))
        scores.append(rng.permutation(np.linspace(0.05, 0.99, max(n, 1))[:n]).astype(np.float32))   # tie-free
    offs = np.zeros(len(sizes) + 1, np.int32)
    offs[1:] = np.cumsum(sizes)
    d = lambda a, dt: torch.from_numpy(np.concatenate(a).astype(dt)).to(dev())
    for top_k, max_keep in ((1000, 20), (50, 50), (2048, 300)):
        buf, K = nms_batch_device(d(boxes, np.float32).view(-1, 4), d(classes, np.int64), d(scores, np.float32),
                                  torch.from_numpy(offs).to(dev()), 0.5, top_k, max_keep)
        kb, kc, ks, ki, cnt = decode_nms_buffer(buf.cpu().numpy(), len(sizes), K)
        for l, n in enumerate(sizes):
            keep = orc.nms(boxes[l], classes[l], scores[l], top_k=top_k, max_boxes=max_keep) if n else np.zeros(0, np.int32)
            assert int(cnt[l]) == len(keep)
            assert np.array_equal(ki[l, :len(keep)], keep)
            assert np.array_equal(kb[l, :len(keep)], boxes[l][keep]) and np.array_equal(ks[l, :len(keep)], scores[l][keep])
            rb, rc, rs = nms(boxes[l], classes[l], scores[l], top_k=top_k, max_boxes=max_keep) if n else ([], [], [])
            assert len(rb) == len(keep) and (not len(keep) or np.array_equal(np.array(rc), classes[l][keep]))
