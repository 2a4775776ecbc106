"""Error behaviour of the drop-in wrappers (bad arguments raise like the reference's argument checks do,
General/Learner.py:339-340; there is no CPU fallback) and size limits (many ground-truth boxes, K != 9)."""
import numpy as np
import pytest
import torch

from neuralnetworklibrary_b200 import _lib
from tests import synth as syn
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_cpu_tensors_are_rejected():
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors = AnchorGenerator()(torch.zeros(1, 3, 64, 64, device=dev()))
    A = anchors.shape[0]
    with pytest.raises(_lib.RetinaB200Error):
        SSD_loss()([anchors, torch.zeros(1, A, 4), torch.full((1, A, 20), 0.01)],
                   [torch.zeros(1, 1, 4, device=dev()), torch.zeros(1, 1, dtype=torch.int64, device=dev())])
    with pytest.raises(_lib.RetinaB200Error):
        BBoxPredictor().predict_arrays(64, 64, torch.zeros(1, A, 4), torch.full((1, A, 20), 0.01), anchors)


def test_bad_arguments_raise_value_error():
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor, nms
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors = AnchorGenerator()(torch.zeros(1, 3, 64, 64, device=dev()))
    A = anchors.shape[0]
    reg, clas = torch.zeros(2, A, 4, device=dev()), torch.full((2, A, 20), 0.01, device=dev())
    with pytest.raises(ValueError):
        BBoxPredictor().predict_arrays(64, 64, reg, clas, anchors, top_k=5000)          # > RN_MAX_TOP_K
    with pytest.raises(ValueError):
        BBoxPredictor().predict_arrays(64, 64, reg[:, :-1], clas, anchors)               # shape mismatch
    with pytest.raises(ValueError):
        nms(torch.zeros(4, 4), torch.zeros(3, dtype=torch.int64), torch.zeros(4))
    with pytest.raises(ValueError):
        SSD_loss()([anchors, reg, clas], [torch.zeros(3, 1, 4, device=dev()), torch.zeros(3, 1, dtype=torch.int64, device=dev())])
    with pytest.raises(ValueError):   # a geometry-tagged table handed over for another image size falls back to the
        wrong = AnchorGenerator()(torch.zeros(1, 3, 96, 96, device=dev()))   # table path and then fails the shape check
        SSD_loss()([wrong, reg, clas], [torch.zeros(2, 1, 4, device=dev()), torch.zeros(2, 1, dtype=torch.int64, device=dev())])


@pytest.mark.parametrize("M", [257, 1000])
def test_many_ground_truth_boxes(M):
    """M far beyond the 32-row warp chunk and the usual padding width."""
    from neuralnetworklibrary_b200.vision import assign_batch
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    H, W, B = 128, 160, 2
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, 20, seed=M, force_empty_and_full=False, min_side=4.0, max_frac=0.5)
    matches, npos, miou = assign_batch(anchors, gb.to(dev()), gc.to(dev()), want_iou=True)
    for i in range(B):
        m, n, iou = orc.assign(an, gb[i].numpy(), gc[i].numpy())
        assert np.array_equal(matches[i].cpu().numpy(), m) and npos[i].item() == n
        assert np.array_equal(miou[i].cpu().numpy().view(np.uint32), iou.view(np.uint32))


@pytest.mark.parametrize("ratios,scales", [([1.0], [1.0]), ([0.5, 1, 2, 3], [1.0, 1.5]), ([0.5, 2], [1, 1.26, 1.59, 2.0, 2.5, 3.0, 3.5, 4.0])])
def test_other_anchor_sets(ratios, scales):
    """K = 1, 8, 16 base boxes per cell (the generic kernels; K = 9 has the register-resident specialisation)."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B, M = 96, 128, 8, 2, 5
    anchors = AnchorGenerator(ratios, scales)(torch.zeros(1, 3, H, W, device=dev()))
    an = orc.anchors(H, W, ratios, scales)
    assert np.array_equal(anchors.cpu().numpy(), an)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=7, min_side=10.0, max_frac=0.7)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=7)
    f = SSD_loss(keep_matches=True)
    cd, rd = clas.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True)
    assert np.array_equal(f.last_assignment[0].cpu().numpy(), o["matches"])
    np.testing.assert_allclose(loss.item(), o["out3"][0], rtol=1e-5)
    syn.assert_rel(cd.grad.cpu().numpy(), o["dclas"])
    syn.assert_dreg_close(rd.grad.cpu().numpy(), o["dreg"])
    ci, ri = syn.make_infer_activations(B, an.shape[0], C, seed=8, anchors=an, mu=-4.0, clusters=4)
    out = BBoxPredictor().predict_arrays(H, W, ri.to(dev()), ci.to(dev()), anchors)
    po = orc.postproc(ci.numpy(), ri.numpy(), an, H, W)
    assert np.array_equal(out["counts"], po["counts"])
    for i, n in enumerate(po["counts"]):
        assert np.array_equal(out["anchor_idx"][i, :n], po["anchor_idx"][i, :n])


def test_empty_image_shard_and_target_validation():
    """An empty shard (more ranks than images) contributes zero and empty gradients; validate_targets=True raises the
    reference's IndexError for a category >= C (Vision.py:1593)."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, M = 96, 128, 8, 4
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    A = anchors.shape[0]
    cd = torch.zeros((0, A, C), device=dev(), requires_grad=True)
    rd = torch.zeros((0, A, 4), device=dev(), requires_grad=True)
    f = SSD_loss(global_batch=5)
    loss = f([anchors, rd, cd], [torch.zeros((0, M, 4), device=dev()), torch.zeros((0, M), dtype=torch.int64, device=dev())])
    loss.backward()
    assert loss.item() == 0.0 and f.reg_loss.item() == 0.0 and f.clas_loss.item() == 0.0
    assert cd.grad.shape == (0, A, C) and rd.grad.shape == (0, A, 4)
    gb, gc = syn.make_targets(2, M, H, W, C, seed=3)
    clas, reg = syn.make_train_activations(2, A, C, seed=3)
    gc[0, 0] = C
    with pytest.raises(IndexError):
        SSD_loss(validate_targets=True)([anchors, reg.to(dev()), clas.to(dev())], [gb.to(dev()), gc.to(dev())])
