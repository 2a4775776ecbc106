"""GPU parity against THE REFERENCE ITSELF running on CUDA (its intended mode; SURVEY.md section 8c names it the primary
oracle).  The unmodified reference sources are imported from oracle/_ref (staged by oracle/stage_ref.py, SHA-256
manifest checked below) -- or from /root/reference where that exists -- and executed on cuda:0 with torch's own kernels;
this package's drop-ins run on the same tensors through the C ABI.

Bars: anchors bitwise; pos/neg index sets and matches bit-exact; candidate / keep lists, classes, scores AND decoded box
coordinates bit-exact (both sides evaluate expf with the device libm); loss, dclas and dreg PURE rtol 1e-5 with atol 0 and
identical zero patterns (both sides evaluate logf with the device libm, so the scaled tolerance needed against the
host-libm C oracle is not used here).  The observed maxima are printed (run with -s) and asserted."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import ref_shim, stage_ref
from tests import ref_runner as ref
from tests import synth as syn

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.available(), reason="reference sources not staged (oracle/stage_ref.py)")]
RTOL = 1e-5
DEV = "cuda:0"


def dev():
    return torch.device(DEV)


def our_anchors(H, W):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    return AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))


def max_rel(actual, expected):
    nz = expected != 0
    if not nz.any():
        return 0.0
    return float((np.abs(actual[nz].astype(np.float64) - expected[nz]) / np.abs(expected[nz])).max())


def test_staged_reference_is_unmodified():
    if not ref_shim.is_staged_copy():
        pytest.skip("running against the checkout itself")
    assert stage_ref.verify() >= 20


@pytest.mark.parametrize("H,W", [(64, 64), (100, 167), (512, 512), (800, 1333), (800, 1344), (608, 1024)])
def test_anchor_generator_vs_reference(H, W):
    want = ref.anchors(H, W, device=DEV)
    got = our_anchors(H, W).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("seed,H,W,M", [(11, 128, 160, 6), (12, 512, 512, 10), (13, 800, 1333, 20), (14, 100, 167, 40)])
def test_match_anchors_objects_vs_reference(seed, H, W, M):
    """pos_idxs / neg_idxs / matches of one image, the reference's return convention (Vision.py:1474-1511), incl. an image
    without objects and a duplicated box (torch.max on CUDA must also return the first maximal index)."""
    from neuralnetworklibrary_b200.vision import match_anchors_objects
    anchors = our_anchors(H, W)
    gb, gc = syn.make_targets(3, M, H, W, 20, seed=seed, min_side=10.0, max_frac=0.7)
    gb[0, 1] = gb[0, 0]
    for i in range(3):
        objs = gb[i][gc[i] >= 0]
        pos, neg, mt = match_anchors_objects(objs.to(dev()), anchors)
        rp, rn_, rm = ref.assign(anchors.cpu().numpy(), objs.numpy(), None, device=DEV)
        assert np.array_equal(pos.cpu().numpy(), rp) and np.array_equal(neg.cpu().numpy(), rn_)
        assert np.array_equal(mt.cpu().numpy(), rm)


LOSS_CASES = [  # seed, H, W, C, B, M, kwargs
    (21, 128, 160, 20, 2, 6, {}),
    (22, 200, 336, 80, 3, 10, dict(beta=0.3, alpha=0.4)),
    (23, 512, 512, 20, 2, 10, {}),                       # BASELINE configs[0] shape
    (24, 96, 128, 7, 2, 5, dict(gamma=1.5)),             # C % 4 != 0, general gamma
    (25, 800, 1333, 80, 2, 20, {}),                      # COCO as named: A = 200 700 (grid widths 167/84/42/21/11)
    (26, 800, 1344, 80, 2, 20, {}),                      # COCO as the collater pads it: A = 201 600
]


@pytest.mark.parametrize("seed,H,W,C,B,M,kw", LOSS_CASES)
def test_ssd_loss_vs_reference(seed, H, W, C, B, M, kw):
    from neuralnetworklibrary_b200.vision import SSD_loss, SSD_ClasLoss, SSD_RegLoss
    anchors = our_anchors(H, W)
    an = anchors.cpu().numpy()
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=12.0, max_frac=0.6)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=seed, edge_cases=64)
    r = ref.loss(an, clas, reg, gb, gc, device=DEV, **kw)
    cd, rd = clas.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
    f = SSD_loss(**kw)
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    got3 = np.array([loss.item(), SSD_RegLoss(f)(None, None).item(), SSD_ClasLoss(f)(None, None).item()], np.float32)
    dclas, dreg = cd.grad.cpu().numpy(), rd.grad.cpu().numpy()
    e3, ec, er = max_rel(got3, r["out3"]), max_rel(dclas, r["dclas"]), max_rel(dreg, r["dreg"])
    print("ssd_loss vs reference on CUDA %dx%d C=%d: max rel err out3 %.2e dclas %.2e dreg %.2e" % (H, W, C, e3, ec, er))
    np.testing.assert_allclose(got3, r["out3"], rtol=RTOL, atol=0)
    syn.assert_rel(dclas, r["dclas"], what="dclas")
    syn.assert_rel(dreg, r["dreg"], what="dreg")       # pure rtol 1e-5, atol 0, identical zero pattern


def test_ssd_loss_no_objects_vs_reference():
    """A batch in which no image has an object (M = 1, all padding): every anchor background, reg_loss 0."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B = 128, 128, 20, 2
    anchors = our_anchors(H, W)
    an = anchors.cpu().numpy()
    gb, gc = torch.full((B, 1, 4), -1.0), torch.full((B, 1), -1, dtype=torch.int64)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=5)
    r = ref.loss(an, clas, reg, gb, gc, device=DEV)
    for _ in range(20):   # repeated: the assignment of an all-padding batch must never race the loss kernel
        cd, rd = clas.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
        f = SSD_loss()
        loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
        loss.backward()
        assert int(f.last_assignment[1].sum()) == 0
        np.testing.assert_allclose(np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32), r["out3"],
                                   rtol=RTOL, atol=0)
        syn.assert_rel(cd.grad.cpu().numpy(), r["dclas"], what="dclas")
        assert not rd.grad.any()


POST_CASES = [  # seed, H, W, C, B, mu, kwargs
    (31, 128, 160, 20, 2, -5.0, {}),
    (32, 256, 320, 80, 3, -5.5, dict(thresh=0.1, max_overlap=0.4, top_k=300, max_boxes=50)),
    (33, 96, 96, 7, 2, -4.0, dict(thresh=0.1, max_overlap=0.3)),
    (34, 800, 1333, 80, 1, -6.0, {}),
    (35, 800, 1344, 80, 1, -6.0, dict(max_boxes=100)),
    (36, 128, 160, 20, 2, -5.0, dict(rel_thresh=[0.3, 0.6])),
    (37, 128, 160, 20, 2, -4.0, dict(inc=[0.7, [1, 2]], max_boxes=50)),
    (38, 128, 160, 20, 2, -4.0, dict(dup=[0.4, [(0, 1), (1, 0), (2, 3)]], max_boxes=50)),
]


@pytest.mark.parametrize("seed,H,W,C,B,mu,kw", POST_CASES)
def test_bbox_predictor_vs_reference(seed, H, W, C, B, mu, kw):
    """BBoxPredictor.__call__ with the reference's 11 positional arguments; the three list-of-lists returned must equal
    the reference's element for element (np.ndarray[4] float32 boxes BITWISE, np.int64 classes, np.float32 scores)."""
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    anchors = our_anchors(H, W)
    an = anchors.cpu().numpy()
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=mu, clusters=8)
    args = (kw.get("thresh", 0.05), kw.get("max_overlap", 0.5), kw.get("rel_thresh"), kw.get("top_k", 1000),
            kw.get("max_boxes", 20), kw.get("dup"), kw.get("inc"))
    rb, rc, rs = ref.postproc(clas, reg, an, H, W, device=DEV, **kw)
    img = torch.zeros(B, 3, H, W, device=dev())
    gb_, gc_, gs_ = BBoxPredictor()(img, reg.to(dev()), clas.to(dev()), anchors, *args)
    assert len(gb_) == len(rb) == B
    total = 0
    for i in range(B):
        assert len(gb_[i]) == len(rb[i]) and len(gc_[i]) == len(rc[i]) and len(gs_[i]) == len(rs[i])
        total += len(rb[i])
        if len(rb[i]) == 0:
            continue
        assert np.array_equal(np.array(gc_[i]), np.array(rc[i])) and np.array(gc_[i]).dtype == np.array(rc[i]).dtype
        assert np.array_equal(np.array(gs_[i], np.float32), np.array(rs[i], np.float32))
        assert np.array_equal(np.stack(gb_[i]).view(np.uint32), np.stack(rb[i]).view(np.uint32)), "decoded boxes differ bitwise"
        assert type(gb_[i][0]) is type(rb[i][0]) and gb_[i][0].dtype == rb[i][0].dtype
    assert total > 0


@pytest.mark.parametrize("kw", [{}, dict(max_overlap=0.3, top_k=200, max_boxes=40), dict(rel_thresh=[0.2, 0.5]),
                                dict(inc=[0.6, [0]], max_boxes=100), dict(dup=[0.3, [(0, 1), (1, 2)]], max_boxes=100)])
def test_nms_vs_reference(kw):
    """nms() on caller-provided tensors (the TTA call site, Vision.py:2118) against the reference's nms on CUDA tensors."""
    from neuralnetworklibrary_b200.retinanet import nms
    g = torch.Generator().manual_seed(77)
    n = 1500
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 80 + 4
    boxes = torch.cat([xy, xy + wh], dim=1)
    classes = torch.randint(0, 4, (n,), generator=g)
    scores = torch.rand(n, generator=g).clamp_min(1e-3)
    scores = torch.unique(scores)[:n]          # tie-free (the reference's sort is unstable, retinanet.py:573)
    boxes, classes = boxes[:len(scores)], classes[:len(scores)]
    perm = torch.randperm(len(scores), generator=g)
    scores = scores[perm]
    rb, rc, rs = ref.nms(boxes, classes, scores, device=DEV, **kw)
    gb_, gc_, gs_ = nms(boxes.to(dev()), classes.to(dev()), scores.to(dev()), **kw)
    assert len(gb_) == len(rb) > 0
    assert np.array_equal(np.array(gc_), np.array(rc)) and np.array_equal(np.array(gs_, np.float32), np.array(rs, np.float32))
    assert np.array_equal(np.stack(gb_).view(np.uint32), np.stack(rb).view(np.uint32))


def test_compute_max_overlaps_vs_reference():
    """ComputeMaxOverlaps metric (Vision.py:1666-1694) on the same batch."""
    from neuralnetworklibrary_b200.vision import ComputeMaxOverlaps
    _, vis = ref_shim.load()
    H, W, B, M = 160, 224, 3, 8
    anchors = our_anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, 20, seed=9)
    ours, theirs = ComputeMaxOverlaps(), vis.ComputeMaxOverlaps()
    a = ours([anchors, None, None], [gb.to(dev()), gc.to(dev())])
    b = theirs([anchors.clone(), None, None], [gb.to(dev()), gc.to(dev())])
    np.testing.assert_allclose(float(a), float(b), rtol=1e-6)
    assert np.array_equal(np.array(ours.max_overlaps, np.float32), np.array([float(v) for v in theirs.max_overlaps], np.float32))
