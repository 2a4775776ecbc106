"""The drop-ins at the reference's OWN call sites: the unmodified reference code (oracle/_ref or /root/reference) runs
Learner.predict, Learner.train1minibatch and ImageLearner.TTA_bbox twice -- once with its own AnchorGenerator /
BBoxPredictor / nms / SSD_loss, once after the substitution INTEGRATION.md section 2 describes -- and the results must
agree: detections element for element (boxes bitwise), loss rtol 1e-5, parameter gradients to accumulated-rounding level.

The CPU test (not gpu) runs the reference-only arm of the same harness so that the harness itself is exercised where there
is no GPU; the `gpu` tests are the comparison proper (the reference then runs on CUDA, its intended mode)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from tests import callsite_harness as H

needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not staged (oracle/stage_ref.py)")
C = 12
IMG_H, IMG_W, N_IMG = 96, 128, 3


def _predict(vis, learner, **kw):
    with torch.no_grad():
        return learner.predict("val", **kw)


def _assert_same_predictions(a, b):
    assert len(a) == len(b)
    total = 0
    for (ab, ac, as_), (bb, bc, bs) in zip(a, b):
        assert len(ab) == len(bb) and len(ac) == len(bc) and len(as_) == len(bs)
        total += len(ab)
        if len(ab) == 0:
            continue
        assert np.array_equal(np.array(ac), np.array(bc))
        assert np.array_equal(np.array(as_, np.float32), np.array(bs, np.float32))
        assert np.array_equal(np.stack(ab).view(np.uint32), np.stack(bb).view(np.uint32))
    return total


@needs_ref
def test_harness_runs_reference_on_cpu(monkeypatch, tmp_path):
    """No GPU needed: Learner.predict / train1minibatch / TTA_bbox of the unmodified reference run inside the harness."""
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu tests on a GPU box")
    rn, vis = H.patch_environment(monkeypatch)
    imgs = H.FakeImages(N_IMG, IMG_H, IMG_W)
    data = H.FakeData(imgs)
    model = H.build_model(vis, C)
    learner = H.make_learner(vis, tmp_path, model, data, vis.SSD_loss())
    preds = _predict(vis, learner, thresh=0.04, max_boxes=30)
    assert len(preds) == N_IMG and sum(len(p[0]) for p in preds) > 0
    x = torch.cat([imgs.batch(j) for j in range(2)])
    gb = torch.tensor([[[10., 12., 60., 70.], [30., 20., 90., 64.]], [[5., 8., 100., 90.], [-1., -1., -1., -1.]]])
    gc = torch.tensor([[1, 4], [7, -1]])
    model.train()
    loss = learner.train1minibatch(x, [gb, gc], 1e-3, 0.9)
    assert np.isfinite(loss) and loss > 0
    H.patch_tta_inputs(monkeypatch, vis, imgs)
    tta = learner.TTA_bbox("val", [H.FakeTransform(N_IMG, False, 1), H.FakeTransform(N_IMG, True, 2)], thresh=0.04)
    assert len(tta) == N_IMG


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(thresh=0.04, max_boxes=30), dict(thresh=0.03, max_overlap=0.4, top_k=200, max_boxes=100),
                                dict(thresh=0.04, rel_thresh=[0.3, 0.6], max_boxes=50)])
def test_learner_predict_with_dropins(monkeypatch, tmp_path, kw):
    """General/Learner.py:349-381 unchanged: model forward -> self.AnchorGenerator(x) (Vision.py:1469) ->
    self.model.BBoxPredictor(x_batch, reg, clas, anchors, thresh, max_overlap, rel_thresh, top_k, max_boxes, dup, inc)
    -> list_mult(PredBoxes, 1/scale)."""
    from neuralnetworklibrary_b200 import retinanet as ours
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    rn, vis = H.patch_environment(monkeypatch)
    imgs = H.FakeImages(N_IMG, IMG_H, IMG_W)
    data = H.FakeData(imgs)
    model = H.build_model(vis, C)
    learner = H.make_learner(vis, tmp_path, model, data, vis.SSD_loss())
    want = _predict(vis, learner, **kw)
    H.swap_predictors(learner.model, ours)
    assert not any(k.startswith(("AnchorGenerator", "BBoxPredictor")) for k in learner.model.state_dict())   # checkpoints keep loading
    got = _predict(vis, learner, **kw)
    assert _assert_same_predictions(got, want) > 0
    a = learner.model.AnchorGenerator(imgs.batch(0).cuda())
    assert a.is_cuda and a.dtype == torch.float32 and a.shape[1] == 4


@needs_ref
@pytest.mark.gpu
def test_train1minibatch_with_dropin_loss(monkeypatch, tmp_path):
    """General/Learner.py:490-516 unchanged: y_pred = model(x); loss = self.loss_func(y_pred, y_batch); loss.backward();
    optimizer.step(); loss.item() -- with ImageLearner(..., loss_func=SSD_loss()) of this package, and SSD_RegLoss /
    SSD_ClasLoss reading the attributes it stores (Vision.py:1643, :1646-1663)."""
    from neuralnetworklibrary_b200 import retinanet as ours
    from neuralnetworklibrary_b200 import vision as ours_v
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    rn, vis = H.patch_environment(monkeypatch)
    imgs = H.FakeImages(4, IMG_H, IMG_W)
    data = H.FakeData(imgs, bs=4)
    base = H.build_model(vis, C)
    x = torch.cat([imgs.batch(j) for j in range(4)])
    gb = torch.tensor([[[10., 12., 60., 70.], [30., 20., 90., 64.], [2., 2., 40., 30.]],
                       [[5., 8., 100., 90.], [-1., -1., -1., -1.], [-1., -1., -1., -1.]],
                       [[-1., -1., -1., -1.], [-1., -1., -1., -1.], [-1., -1., -1., -1.]],
                       [[64., 32., 120., 90.], [20., 40., 52., 72.], [-1., -1., -1., -1.]]])
    gc = torch.tensor([[1, 4, 0], [7, -1, -1], [-1, -1, -1], [11, 3, -1]])
    results = []
    for arm in ("reference", "dropin"):
        model = H.clone_model(base)
        loss_func = vis.SSD_loss() if arm == "reference" else ours_v.SSD_loss()
        if arm == "dropin":
            H.swap_predictors(model, ours)
        learner = H.make_learner(vis, tmp_path / arm, model, data, loss_func)
        learner.model.train()
        loss = learner.train1minibatch(x.cuda(), [gb.cuda(), gc.cuda()], 1e-3, 0.9)
        metrics = (vis.SSD_RegLoss(loss_func), vis.SSD_ClasLoss(loss_func)) if arm == "reference" else \
            (ours_v.SSD_RegLoss(loss_func), ours_v.SSD_ClasLoss(loss_func))
        results.append((loss, float(metrics[0](None, None)), float(metrics[1](None, None)),
                        {n: p.grad.detach().clone() for n, p in learner.model.named_parameters() if p.grad is not None},
                        {n: p.detach().clone() for n, p in learner.model.named_parameters()}))
    (l0, r0, c0, g0, p0), (l1, r1, c1, g1, p1) = results
    np.testing.assert_allclose([l1, r1, c1], [l0, r0, c0], rtol=1e-5, atol=0)
    assert g0.keys() == g1.keys() and len(g0) > 10
    worst = 0.0
    for n in g0:
        scale = g0[n].abs().max().item()
        if scale == 0:
            assert not g1[n].any()
            continue
        worst = max(worst, (g1[n] - g0[n]).abs().max().item() / scale)
    print("parameter gradients, drop-in loss vs reference loss: max |diff| / max |grad| over %d tensors = %.2e" % (len(g0), worst))
    assert worst < 1e-4      # activation gradients agree to 1e-5 relative; the backbone's backward accumulates them
    for n in p0:
        assert torch.allclose(p0[n], p1[n], rtol=1e-4, atol=1e-7)


@needs_ref
@pytest.mark.gpu
def test_tta_bbox_with_dropins(monkeypatch, tmp_path):
    """Applications/Vision.py:2036-2121 unchanged: five passes through self.model.BBoxPredictor, the un-transform, then
    TEN(boxes) / TEN(classes) / TEN(scores) and vmods.retinanet.nms(...) on the merged predictions (Vision.py:2104-2119)."""
    from neuralnetworklibrary_b200 import retinanet as ours
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    rn, vis = H.patch_environment(monkeypatch)
    imgs = H.FakeImages(N_IMG, IMG_H, IMG_W)
    data = H.FakeData(imgs)
    model = H.build_model(vis, C)
    learner = H.make_learner(vis, tmp_path, model, data, vis.SSD_loss())
    tfms = [H.FakeTransform(N_IMG, False, 1), H.FakeTransform(N_IMG, True, 2)]
    H.patch_tta_inputs(monkeypatch, vis, imgs)
    with torch.no_grad():
        want = learner.TTA_bbox("val", tfms, thresh=0.04, max_boxes=30)
    H.swap_predictors(learner.model, ours)
    H.swap_nms(monkeypatch, rn, ours)
    H.patch_tta_inputs(monkeypatch, vis, imgs)
    with torch.no_grad():
        got = learner.TTA_bbox("val", tfms, thresh=0.04, max_boxes=30)
    assert _assert_same_predictions(got, want) > 0
