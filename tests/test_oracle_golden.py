"""Pins the CPU oracle (oracle/retina_oracle.c) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  Integer / index results must be identical; fp32 results
within rtol 1e-5 (BASELINE.json north_star)."""
import hashlib
import os

import numpy as np
import pytest

from tests import synth as syn
from oracle import oracle as orc

RTOL = 1e-5


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


assert_rel = syn.assert_rel


def test_anchor_tables_bitwise(golden_dir):
    g = _load(golden_dir, "anchors.npz")
    for key in ("full_64x64", "full_100x167"):
        H, W = (int(v) for v in key.split("_")[1].split("x"))
        a = orc.anchors(H, W)
        assert a.dtype == np.float32
        assert np.array_equal(a.view(np.uint32), g[key].view(np.uint32))
    for (H, W), n, h in zip(g["shapes"], g["counts"], g["sha256"]):
        a = orc.anchors(int(H), int(W))
        assert a.shape == (int(n), 4)
        assert orc.num_anchors(int(H), int(W)) == int(n)
        assert _sha(a) == str(h)
    a = orc.anchors(512, 512)
    assert np.array_equal(a[:12], g["a512_head"]) and np.array_equal(a[-3:], g["a512_tail"])


def test_known_anchor_counts():
    # SURVEY.md section 8c
    assert orc.num_anchors(512, 512) == 49104
    assert orc.num_anchors(800, 1333) == 200700
    assert orc.num_anchors(800, 1344) == 201600
    assert orc.num_anchors(608, 1024) == 116712
    assert orc.num_anchors(608, 1216) == 138618
    assert syn.num_anchors(800, 1344) == 201600


@pytest.mark.parametrize("variant,kw", [("default", {}), ("beta03_alpha04", dict(beta=0.3, alpha=0.4))])
def test_loss_small(golden_dir, variant, kw):
    g = _load(golden_dir, "loss_small.npz")
    o = orc.loss(g["anchors"], g["clas"], g["reg"], g["gt_boxes"], g["gt_cats"], want_matches=True, **kw)
    assert np.array_equal(o["matches"], g["matches"])
    assert np.array_equal(o["npos"], (g["matches"] >= 0).sum(1))
    np.testing.assert_allclose(o["out3"], g[variant + "_out3"], rtol=RTOL, atol=0)
    assert_rel(o["dclas"], g[variant + "_dclas"])
    assert_rel(o["dreg"], g[variant + "_dreg"])


def test_loss_cfg1(golden_dir):
    """BASELINE.json configs[0] (B=2, 512x512, 20 classes, <=10 GT / image)."""
    g = _load(golden_dir, "loss_cfg1.npz")
    H, W, C, B, M = (int(g[k]) for k in "HWCBM")
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=int(g["seed"]))
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=int(g["seed"]))
    if _sha(clas.numpy()) != str(g["sha_clas"]) or _sha(reg.numpy()) != str(g["sha_reg"]):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    assert np.array_equal(gb.numpy(), g["gt_boxes"]) and np.array_equal(gc.numpy(), g["gt_cats"])
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True)
    assert np.array_equal(o["matches"].astype(np.int8), g["matches"])
    np.testing.assert_allclose(o["out3"], g["out3"], rtol=RTOL, atol=0)
    assert_rel(o["dclas"].reshape(-1)[g["sample_idx"]], g["sample_dclas"])
    assert_rel(o["dclas"].reshape(-1, C)[g["pos_rows"]], g["pos_dclas"])
    assert_rel(o["dreg"].reshape(-1, 4)[g["pos_rows"]], g["pos_dreg"])
    assert np.count_nonzero(np.abs(o["dreg"]).sum(-1)) == int(g["dreg_nonzero_rows"])
    np.testing.assert_allclose(np.abs(o["dclas"].astype(np.float64)).sum(), float(g["dclas_abs_sum"]), rtol=1e-6)


POST_VARIANTS = [("default", dict()),
                 ("topk50_max100", dict(top_k=50, max_boxes=100)),
                 ("thr02_ov03_max7", dict(thresh=0.2, max_overlap=0.3, max_boxes=7)),
                 ("thr001_max1000", dict(thresh=0.01, max_boxes=1000))]


@pytest.mark.parametrize("variant,kw", POST_VARIANTS)
def test_postproc_small(golden_dir, variant, kw):
    g = _load(golden_dir, "postproc_small.npz")
    o = orc.postproc(g["clas"], g["reg"], g["anchors"], int(g["H"]), int(g["W"]), **kw)
    assert np.array_equal(o["counts"], g[variant + "_counts"])
    if kw.get("thresh", 0.05) >= 0.05:
        assert o["counts"][2] == 0  # the image with every score below 0.05
    for i, n in enumerate(o["counts"]):
        assert np.array_equal(o["classes"][i, :n], g[variant + "_classes"][i, :n])
        assert np.array_equal(o["scores"][i, :n], g[variant + "_scores"][i, :n])
        np.testing.assert_allclose(o["boxes"][i, :n], g[variant + "_boxes"][i, :n], rtol=RTOL, atol=0)
        # anchor_idx is consistent with the class-max scores
        assert np.array_equal(g["clas"][i][o["anchor_idx"][i, :n]].max(1), o["scores"][i, :n])


@pytest.mark.parametrize("variant,kw", [("all", dict(top_k=3000, max_boxes=100000)), ("default", dict()),
                                        ("ov07_topk500_max50", dict(max_overlap=0.7, top_k=500, max_boxes=50))])
def test_nms_boxes(golden_dir, variant, kw):
    g = _load(golden_dir, "nms_boxes.npz")
    keep = orc.nms(g["boxes"], g["classes"], g["scores"], **kw)
    assert np.array_equal(g["boxes"][keep], g[variant + "_boxes"])
    assert np.array_equal(g["classes"][keep], g[variant + "_classes"])
    assert np.array_equal(g["scores"][keep], g[variant + "_scores"])


def test_assign_edge_cases():
    an = orc.anchors(64, 64)
    A = an.shape[0]
    # no objects: everything negative (reference Vision.py:1498-1501)
    m, npos, _ = orc.assign(an, np.zeros((0, 4), np.float32), np.zeros(0, np.int64))
    assert npos == 0 and (m == orc.MATCH_NEG).all()
    # all padding
    m, npos, _ = orc.assign(an, -np.ones((4, 4), np.float32), -np.ones(4, np.int64))
    assert npos == 0 and (m == orc.MATCH_NEG).all()
    # a GT equal to an anchor: IoU exactly 1 there; duplicated GT -> first index wins
    k = 3 + 9 * 5
    gt = np.stack([an[k], an[k], an[100]])
    m, npos, miou = orc.assign(an, gt, np.array([1, 2, 3], np.int64))
    assert m[k] == 0 and miou[k] == 1.0 and m[100] == 2
    # zero-area GT never matches
    z = np.array([[10, 10, 10, 30]], np.float32)
    m, npos, miou = orc.assign(an, z, np.array([0], np.int64))
    assert npos == 0 and (miou == 0).all()
    # thresholds are strict on both sides (Vision.py:1506-1507): IoU == 0.5 is ignored, == 0.4 too
    box = np.array([[0, 0, 10, 10]], np.float32)
    half = np.array([[0, 0, 10, 5]], np.float32)      # IoU 0.5 exactly
    m, _, miou = orc.assign(box, half, np.array([0], np.int64))
    assert miou[0] == 0.5 and m[0] == orc.MATCH_IGNORE
    forty = np.array([[0, 0, 10, 4]], np.float32)     # IoU 0.4 exactly
    m, _, miou = orc.assign(box, forty, np.array([0], np.int64))
    assert miou[0] == np.float32(0.4) and m[0] == orc.MATCH_IGNORE
    assert A == 774
