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


@pytest.mark.gpu
def test_stage_images_uint8_upload_and_normalisation():
    """Extension: 8-bit images are uploaded as bytes and converted (optionally normalised) on the device."""
    from neuralnetworklibrary_b200.vision import stage_images
    rng = np.random.RandomState(11)
    imgs = [rng.randint(0, 256, size=(int(rng.randint(40, 90)), int(rng.randint(50, 130)), 3)).astype(np.uint8) for _ in range(3)]
    plain = stage_images(imgs, 3, 5)
    assert np.array_equal(plain.cpu().numpy(), orc.stage_images([im.astype(np.float32) for im in imgs], 3, 5))
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    got = stage_images(imgs, 3, 5, mean=mean, std=std).cpu().numpy()
    want = orc.stage_images([((im.astype(np.float32) / np.float32(255)) - mean) / std for im in imgs], 3, 5)
    # padding stays exactly 0, pixels equal the float32 restatement bit for bit
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        stage_images([im.astype(np.float32) for im in imgs], mean=mean, std=std)


@pytest.mark.gpu
def test_merge_tta_predictions_device_untransform():
    """merge_tta_predictions(transforms=...) = the reference's host un-transform (Vision.py:2091-2097, float64 like NumPy >= 2
    evaluates it for int64 jitter values, rounded to float32 by TEN) followed by the merge without transforms."""
    from neuralnetworklibrary_b200.vision import merge_tta_predictions
    rng = np.random.RandomState(7)
    L, NP = 4, 5
    passes, transforms, host_passes = [], [], []
    for i in range(NP):
        per_image, per_tf, per_host = [], [], []
        for l in range(L):
            n = int(rng.randint(0, 15)) if not (l == 2 and i < 3) else 0
            xy = rng.uniform(0, 300, (n, 2))
            wh = rng.uniform(10, 120, (n, 2))
            boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
            classes = list(rng.randint(0, 3, n).astype(np.int64))
            scores = list(rng.uniform(0.05, 1, n).astype(np.float32))
            tf = dict(row_jit=np.int64(rng.randint(0, 9)), col_jit=np.int64(rng.randint(0, 9)), rand_scale=np.float64(rng.uniform(0.9, 1.1)),
                      scale=float(rng.uniform(0.4, 1.2)), flip=int(rng.randint(0, 2)), cols=int(rng.randint(300, 700)))
            per_image.append([list(boxes), classes, scores])
            per_tf.append(tf)
            hb = []
            if n:   # the reference's lines, on the host
                b = np.array(list(boxes))
                b = np.array([b[:, 0] - tf["col_jit"], b[:, 1] - tf["row_jit"], b[:, 2] - tf["col_jit"], b[:, 3] - tf["row_jit"]]).T
                b = (1 / (tf["rand_scale"] * tf["scale"])) * b
                if i > 0 and tf["flip"] == 1:
                    b = np.array([tf["cols"] - b[:, 2], b[:, 1], tf["cols"] - b[:, 0], b[:, 3]]).T
                hb = list(b.astype(np.float32))
            per_host.append([hb, classes, scores])
        passes.append(per_image)
        transforms.append(per_tf)
        host_passes.append(per_host)
    for kw in (dict(max_boxes=50), dict(max_boxes=50, rel_thresh=[0.2, 0.5])):
        got = merge_tta_predictions(passes, transforms=transforms, **kw)
        want = merge_tta_predictions(host_passes, **kw)
        assert len(got) == len(want) == L
        for g_, w_ in zip(got, want):
            assert len(g_[0]) == len(w_[0])
            if len(g_[0]):
                assert np.array_equal(np.stack(g_[0]).view(np.uint32), np.stack(w_[0]).view(np.uint32))
                assert np.array_equal(np.array(g_[1]), np.array(w_[1])) and np.array_equal(np.array(g_[2]), np.array(w_[2]))
