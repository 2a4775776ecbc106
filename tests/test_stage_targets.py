"""Target staging: the bounding-box half of the reference's AspectRatioCollater
(Vision.py:770-785, :798-812).  CPU: the numpy oracle vs the reference-generated golden; GPU: the kernel vs
both, bit for bit."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.collater_cases()


def test_oracle_matches_reference_collater(golden_dir):
    g = np.load(os.path.join(golden_dir, "collater_targets.npz"))
    for k, (boxes, cats, scales, rs, rj, cj) in enumerate(_cases()):
        bp, cp = orc.stage_targets(boxes, cats, scales, rs, rj, cj)
        assert np.array_equal(bp, g["case%d_boxes" % k]) and bp.dtype == np.float32
        assert np.array_equal(cp, g["case%d_cats" % k]) and cp.dtype == np.int64


@pytest.mark.gpu
def test_kernel_matches_reference_collater(golden_dir):
    from neuralnetworklibrary_b200.vision import SSD_loss, stage_targets
    g = np.load(os.path.join(golden_dir, "collater_targets.npz"))
    for k, (boxes, cats, scales, rs, rj, cj) in enumerate(_cases()):
        BBoxes, Cats = stage_targets(boxes, cats, scales, rs, rj, cj)
        assert BBoxes.is_cuda and BBoxes.dtype == torch.float32 and Cats.dtype == torch.int64
        assert np.array_equal(BBoxes.cpu().numpy(), g["case%d_boxes" % k])
        assert np.array_equal(Cats.cpu().numpy(), g["case%d_cats" % k])
    # the staged tensors feed the loss directly
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    dev = torch.device("cuda:0")
    anchors = AnchorGenerator()(torch.zeros(1, 3, 256, 320, device=dev))
    boxes, cats, scales, rs, rj, cj = _cases()[0]
    BBoxes, Cats = stage_targets(boxes, cats, scales, rs, rj, cj)
    B, A = BBoxes.shape[0], anchors.shape[0]
    loss = SSD_loss()([anchors, torch.zeros(B, A, 4, device=dev), torch.full((B, A, 20), 0.01, device=dev)], [BBoxes, Cats])
    assert torch.isfinite(loss).item()


@pytest.mark.gpu
def test_merge_tta_predictions_equals_nms_of_union():
    from neuralnetworklibrary_b200.vision import merge_tta_predictions
    rng = np.random.RandomState(3)
    passes = []
    for p in range(5):
        per_image = []
        for l in range(3):
            n = int(rng.randint(0, 12)) if l != 1 else 0
            xy = rng.uniform(0, 200, (n, 2))
            wh = rng.uniform(10, 80, (n, 2))
            per_image.append([list(np.concatenate([xy, xy + wh], 1).astype(np.float32)),
                              list(rng.randint(0, 3, n).astype(np.int64)), list(rng.uniform(0.05, 1, n).astype(np.float32))])
        passes.append(per_image)
    merged = merge_tta_predictions(passes, max_boxes=50)
    for l in range(3):
        b = [x for p in passes for x in p[l][0]]
        c = [x for p in passes for x in p[l][1]]
        s = [x for p in passes for x in p[l][2]]
        if not b:
            assert merged[l] == [[], [], []]
            continue
        keep = orc.nms(np.stack(b), np.array(c), np.array(s), max_boxes=50)
        assert np.array_equal(np.array(merged[l][2], np.float32), np.array(s, np.float32)[keep])
        assert np.array_equal(np.stack(merged[l][0]), np.stack(b)[keep])


def _image_cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.image_cases()


def test_oracle_matches_reference_collater_images(golden_dir):
    """Pixel half of the collater (Vision.py:775-777, :786, :790-796): numpy oracle vs the reference-generated golden."""
    g = np.load(os.path.join(golden_dir, "collater_images.npz"))
    for k, (imgs, rj, cj) in enumerate(_image_cases()):
        out = orc.stage_images(imgs, rj, cj)
        assert out.dtype == np.float32 and np.array_equal(out, g["case%d" % k])


@pytest.mark.gpu
def test_kernel_matches_reference_collater_images(golden_dir):
    from neuralnetworklibrary_b200.vision import stage_images
    g = np.load(os.path.join(golden_dir, "collater_images.npz"))
    for k, (imgs, rj, cj) in enumerate(_image_cases()):
        out = stage_images(imgs, rj, cj)
        assert out.is_cuda and out.dtype == torch.float32
        assert np.array_equal(out.cpu().numpy(), g["case%d" % k])
    # a COCO-sized batch against the oracle (the reference's own collater needs cv2 and seconds per batch)
    rng = np.random.RandomState(3)
    imgs = [rng.rand(int(rng.randint(600, 801)), int(rng.randint(900, 1334)), 3).astype(np.float32) for _ in range(4)]
    out = stage_images(imgs, 7, 13)
    assert out.shape[2] % 32 == 0 and out.shape[3] % 32 == 0
    assert np.array_equal(out.cpu().numpy(), orc.stage_images(imgs, 7, 13))
    with pytest.raises(ValueError):
        stage_images([])
