"""GPU parity of the level-tensor loss (rn_loss_levels): SSD_loss called with the heads'
NCHW conv outputs per pyramid level instead of the permuted + concatenated [bs, A, 4|C] tensors.  The checker is the
CPU oracle on the flat tensors, with the layout map of retinanet.py:215-217, :289-295 / Vision.py:1467-1468 restated
in oracle.heads_to_flat (pinned against the unmodified reference heads in tests/test_oracle_vs_reference.py).
Assignments bit-exact; loss and gradients rtol 1e-5 with identical zero patterns (dreg: the scaled tolerance of
tests/synth.py)."""
import numpy as np
import pytest
import torch

from tests import synth as syn
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def dev():
    return torch.device("cuda:0")


def _setup(seed, H, W, C, B, M, logits=False):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import level_shapes
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=10.0, max_frac=0.7)
    g = torch.Generator().manual_seed(seed)
    if logits:
        flat_c = torch.randn(B, an.shape[0], C, generator=g) * 1.5 - 4.6
        flat_c.view(-1)[:64] = torch.linspace(-12, 12, 64)
        flat_r = torch.randn(B, an.shape[0], 4, generator=g) * 0.5
    else:
        flat_c, flat_r = syn.make_train_activations(B, an.shape[0], C, seed=seed, edge_cases=64)
    cs, rs = level_shapes(H, W, 9, C), level_shapes(H, W, 9, 4)
    clas_lv = [torch.from_numpy(x) for x in orc.flat_to_heads(flat_c.numpy(), cs)]
    reg_lv = [torch.from_numpy(x) for x in orc.flat_to_heads(flat_r.numpy(), rs)]
    assert np.array_equal(orc.heads_to_flat([t.numpy() for t in clas_lv], C), flat_c.numpy())
    return anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs


# shapes: all levels V=4 (128x128: P = 256 .. 1); mixed V (100x167: 2171 odd, 546 even ...; 800x1344-like 200x336)
@pytest.mark.parametrize("seed,H,W,C,B,M,kw", [
    (301, 128, 128, 20, 2, 6, {}),
    (302, 100, 167, 80, 2, 5, {}),
    (303, 200, 336, 80, 2, 8, dict(beta=0.3, alpha=0.4)),
    (304, 96, 160, 7, 3, 4, {}),
    (305, 64, 96, 12, 2, 4, dict(gamma=1.5)),
    (306, 256, 256, 91, 1, 10, {}),
    (307, 800, 1333, 80, 1, 20, {}),     # COCO as named (A = 200 700): grid widths 167/84/42/21/11 -> the V=1 and V=2 plane paths at full size
])
def test_levels_loss_matches_oracle(seed, H, W, C, B, M, kw):
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs = _setup(seed, H, W, C, B, M)
    cd = [t.to(dev()).requires_grad_(True) for t in clas_lv]
    rd = [t.to(dev()).requires_grad_(True) for t in reg_lv]
    f = SSD_loss(**kw)
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    o = orc.loss(an, flat_c.numpy(), flat_r.numpy(), gb.numpy(), gc.numpy(), want_matches=True, **kw)
    matches, npos = f.last_assignment
    assert np.array_equal(matches.cpu().numpy(), o["matches"])
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got3, o["out3"], rtol=RTOL, atol=0)
    syn.assert_rel(orc.heads_to_flat([t.grad.cpu().numpy() for t in cd], C), o["dclas"], what="dclas")
    syn.assert_dreg_close(orc.heads_to_flat([t.grad.cpu().numpy() for t in rd], 4), o["dreg"])
    # the flat path of this library on the permuted tensors: same assignment, loss within rtol
    with torch.no_grad():
        l_flat = SSD_loss(**kw)([anchors, flat_r.to(dev()), flat_c.to(dev())], [gb.to(dev()), gc.to(dev())])
    np.testing.assert_allclose(l_flat.item(), loss.item(), rtol=RTOL)


@pytest.mark.parametrize("seed,H,W,C,B,M", [(311, 128, 160, 80, 2, 6), (312, 100, 167, 20, 2, 5), (313, 96, 96, 7, 2, 4),
                                                 (314, 800, 1333, 80, 1, 20)])
def test_levels_loss_from_logits(seed, H, W, C, B, M):
    """Sigmoid fused (as rn_loss_logits): the oracle is fed the kernel's own probabilities, see
    tests/test_gpu_assign_loss.py::test_loss_from_logits for why."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an, gb, gc, flat_z, flat_r, z_lv, reg_lv, cs, rs = _setup(seed, H, W, C, B, M, logits=True)
    zd = [t.to(dev()).requires_grad_(True) for t in z_lv]
    rd = [t.to(dev()).requires_grad_(True) for t in reg_lv]
    f = SSD_loss(from_logits=True, keep_probs=True)
    loss = f([anchors, rd, zd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    y = orc.heads_to_flat([t.cpu().numpy() for t in f.last_probs], C)
    exact = 1.0 / (1.0 + np.exp(-flat_z.numpy().astype(np.float64)))
    ulp = np.abs(y.view(np.int32).astype(np.int64) - exact.astype(np.float32).view(np.int32).astype(np.int64))
    assert ulp.max() <= 6
    o = orc.loss(an, y, flat_r.numpy(), gb.numpy(), gc.numpy(), want_matches=True)
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got3, o["out3"], rtol=RTOL, atol=0)
    syn.assert_rel(orc.heads_to_flat([t.grad.cpu().numpy() for t in zd], C), (o["dclas"] * (np.float32(1) - y)) * y, what="dlogits")
    syn.assert_dreg_close(orc.heads_to_flat([t.grad.cpu().numpy() for t in rd], 4), o["dreg"])


def test_levels_no_grad_and_upstream_scale():
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs = _setup(321, 96, 128, 20, 2, 5)
    tgt = [gb.to(dev()), gc.to(dev())]
    f = SSD_loss()
    with torch.no_grad():
        l0 = f([anchors, [t.to(dev()) for t in reg_lv], [t.to(dev()) for t in clas_lv]], tgt)
    cd = [t.to(dev()).requires_grad_(True) for t in clas_lv]
    rd = [t.to(dev()).requires_grad_(True) for t in reg_lv]
    l1 = f([anchors, rd, cd], tgt)
    assert l0.item() == l1.item()
    (l1 * 3.0).backward()
    o = orc.loss(an, flat_c.numpy(), flat_r.numpy(), gb.numpy(), gc.numpy())
    syn.assert_rel(orc.heads_to_flat([t.grad.cpu().numpy() for t in cd], 20), o["dclas"] * np.float32(3), what="dclas*3")


def test_levels_rejects_wrong_shapes():
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs = _setup(322, 64, 64, 4, 1, 2)
    tgt = [gb.to(dev()), gc.to(dev())]
    bad = [t.to(dev()) for t in clas_lv]
    bad[2] = bad[2][:, :, :-1].contiguous()
    with pytest.raises(ValueError):
        SSD_loss()([anchors, [t.to(dev()) for t in reg_lv], bad], tgt)
    with pytest.raises(ValueError):
        SSD_loss()([anchors.clone(), [t.to(dev()) for t in reg_lv], [t.to(dev()) for t in clas_lv]], tgt)


def test_levels_all_images_empty_and_other_anchor_sets():
    """No ground truth at all (every anchor background, reg loss 0, Vision.py:1498-1501 / :1603) and an anchor set with
    K != 9 (2 ratios x 2 scales): the level layout follows K."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import SSD_loss, level_shapes
    H, W, C, B = 96, 136, 20, 2
    # (a) empty ground truth, default anchors
    anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs = _setup(331, H, W, C, B, 3)
    gb[:] = -1
    gc[:] = -1
    cd = [t.to(dev()).requires_grad_(True) for t in clas_lv]
    rd = [t.to(dev()).requires_grad_(True) for t in reg_lv]
    f = SSD_loss()
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    o = orc.loss(an, flat_c.numpy(), flat_r.numpy(), gb.numpy(), gc.numpy())
    np.testing.assert_allclose(np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32), o["out3"], rtol=RTOL, atol=0)
    assert f.reg_loss.item() == 0.0 and all(float(t.grad.abs().sum()) == 0.0 for t in rd)
    syn.assert_rel(orc.heads_to_flat([t.grad.cpu().numpy() for t in cd], C), o["dclas"], what="dclas")
    # (b) K = 4
    ratios, scales = [0.5, 2.0], [1.0, 1.5]
    ag = AnchorGenerator(ratios, scales)
    anchors4 = ag(torch.zeros(1, 3, H, W, device=dev()))
    an4 = orc.anchors(H, W, ratios, scales)
    assert np.array_equal(anchors4.cpu().numpy(), an4)
    gb4, gc4 = syn.make_targets(B, 5, H, W, C, seed=332, min_side=10.0, max_frac=0.7)
    fc, fr = syn.make_train_activations(B, an4.shape[0], C, seed=332, edge_cases=32)
    cl4 = [torch.from_numpy(x).to(dev()).requires_grad_(True) for x in orc.flat_to_heads(fc.numpy(), level_shapes(H, W, 4, C))]
    rg4 = [torch.from_numpy(x).to(dev()).requires_grad_(True) for x in orc.flat_to_heads(fr.numpy(), level_shapes(H, W, 4, 4))]
    f4 = SSD_loss()
    l4 = f4([anchors4, rg4, cl4], [gb4.to(dev()), gc4.to(dev())])
    l4.backward()
    o4 = orc.loss(an4, fc.numpy(), fr.numpy(), gb4.numpy(), gc4.numpy(), want_matches=True)
    assert np.array_equal(f4.last_assignment[0].cpu().numpy(), o4["matches"])
    np.testing.assert_allclose(np.array([l4.item(), f4.reg_loss.item(), f4.clas_loss.item()], np.float32), o4["out3"], rtol=RTOL, atol=0)
    syn.assert_rel(orc.heads_to_flat([t.grad.cpu().numpy() for t in cl4], C), o4["dclas"], what="dclas K=4")
    syn.assert_dreg_close(orc.heads_to_flat([t.grad.cpu().numpy() for t in rg4], 4), o4["dreg"])


@pytest.mark.parametrize("seed,H,W,C,B,mu,kw", [(341, 128, 160, 12, 2, -5.0, {}), (342, 100, 167, 80, 2, -5.5, dict(thresh=0.1, max_overlap=0.4, top_k=300, max_boxes=50)),
                                                (343, 256, 320, 20, 3, -5.0, dict(top_k=64, max_boxes=64)), (344, 800, 1344, 80, 1, -6.0, {}),
                                                (345, 800, 1333, 80, 1, -6.0, {})])
def test_predictor_on_level_tensors(seed, H, W, C, B, mu, kw):
    """rn_postproc_levels (BBoxPredictor called with the heads' per-level NCHW tensors) against the flat path of this
    library and the oracle: candidate counts, keep indices, classes, scores and boxes identical.  With logits in, the
    flat path is fed torch.sigmoid(logits) -- the fused sigmoid uses the same operations, so everything stays bit-equal."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import level_shapes
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=mu, clusters=6)
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    bp = BBoxPredictor()
    flat = bp.predict_arrays(H, W, reg.to(dev()), clas.to(dev()), anchors, **kw)
    cl = [torch.from_numpy(x).to(dev()) for x in orc.flat_to_heads(clas.numpy(), level_shapes(H, W, 9, C))]
    rl = [torch.from_numpy(x).to(dev()) for x in orc.flat_to_heads(reg.numpy(), level_shapes(H, W, 9, 4))]
    lv = bp.predict_arrays(H, W, rl, cl, anchors, **kw)
    po = orc.postproc(clas.numpy(), reg.numpy(), an, H, W, **kw)
    assert np.array_equal(lv["counts"], po["counts"]) and np.array_equal(lv["n_candidates"], po["n_candidates"])
    for key in ("counts", "n_candidates"):
        assert np.array_equal(flat[key], lv[key])
    for i in range(B):
        n = int(flat["counts"][i])
        for key in ("anchor_idx", "classes", "scores"):
            assert np.array_equal(flat[key][i, :n], lv[key][i, :n]), key
        assert np.array_equal(flat["boxes"][i, :n].view(np.uint32), lv["boxes"][i, :n].view(np.uint32))
        assert np.array_equal(lv["anchor_idx"][i, :n], po["anchor_idx"][i, :n])
    assert int(flat["counts"].sum()) > 0
    # logits in
    z = [torch.randn_like(t) * 1.5 + mu for t in cl]
    for t in z:                                     # plant confident clusters so that NMS has something to do
        t.view(-1)[:: max(1, t.numel() // 97)] += 9.0
    # probability ties between different logits: saturated logits (sigmoid == 1.0) and neighbours one ulp apart, with
    # the smaller logit at the LOWER class index -- torch.max over probabilities keeps that lower index
    z0 = z[0].view(B, 9, C, -1)
    z0[:, 0, 1, 0], z0[:, 0, C - 1, 0] = 20.0, 25.0
    z0[:, 1, 2, 1], z0[:, 1, 3, 1] = 17.5, 18.5
    z0[:, 2, 0, 2] = 3.0
    z0[:, 2, 4, 2] = torch.nextafter(torch.tensor(3.0), torch.tensor(4.0)).item()
    bz = BBoxPredictor()
    bz.from_logits = True
    lz = bz.predict_arrays(H, W, rl, z, anchors, **kw)
    _, probs = bz.flatten_levels(rl, z)             # torch.sigmoid + the reference's layout ops
    fz = bp.predict_arrays(H, W, reg.to(dev()), probs, anchors, **kw)
    assert np.array_equal(fz["counts"], lz["counts"]) and np.array_equal(fz["n_candidates"], lz["n_candidates"])
    for i in range(B):
        n = int(fz["counts"][i])
        for key in ("anchor_idx", "classes", "scores"):
            assert np.array_equal(fz[key][i, :n], lz[key][i, :n]), key
        assert np.array_equal(fz["boxes"][i, :n].view(np.uint32), lz["boxes"][i, :n].view(np.uint32))
    assert int(lz["counts"].sum()) > 0


def test_predictor_call_accepts_level_tensors():
    """The list-returning __call__ (what Learner.predict consumes) with level tensors."""
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import level_shapes
    H, W, C, B = 128, 160, 12, 2
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=345, anchors=an, mu=-5.0, clusters=6)
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    img = torch.zeros(B, 3, H, W, device=dev())
    bp = BBoxPredictor()
    flat = bp(img, reg.to(dev()), clas.to(dev()), anchors)
    cl = [torch.from_numpy(x).to(dev()) for x in orc.flat_to_heads(clas.numpy(), level_shapes(H, W, 9, C))]
    rl = [torch.from_numpy(x).to(dev()) for x in orc.flat_to_heads(reg.numpy(), level_shapes(H, W, 9, 4))]
    lv = bp(img, rl, cl, anchors, 0.05, 0.5, [0.3, 0.6], 1000, 20, None, None)   # with a host stage (rel_thresh)
    fl = bp(img, reg.to(dev()), clas.to(dev()), anchors, 0.05, 0.5, [0.3, 0.6], 1000, 20, None, None)
    for a, b in zip(fl, lv):
        for x, y in zip(a, b):
            assert len(x) == len(y) and all(np.array_equal(u, v) for u, v in zip(x, y))
    assert sum(len(x) for x in flat[0]) > 0


def test_levels_capture_replay():
    """SSD_loss.capture() with per-level lists: graph replays reproduce the eager result and follow the static inputs."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an, gb, gc, flat_c, flat_r, clas_lv, reg_lv, cs, rs = _setup(351, 128, 160, 20, 2, 6)
    cd, rd = [t.to(dev()) for t in clas_lv], [t.to(dev()) for t in reg_lv]
    gbd, gcd = gb.to(dev()), gc.to(dev())
    f = SSD_loss()
    step = f.capture([anchors, rd, cd], [gbd, gcd])
    step.replay()
    ce = [t.clone().requires_grad_(True) for t in cd]
    re_ = [t.clone().requires_grad_(True) for t in rd]
    le = f([anchors, re_, ce], [gbd, gcd])
    le.backward()
    assert step.loss.item() == le.item()
    for a, b in zip(step.dclas_levels, ce):
        assert torch.equal(a, b.grad)
    for a, b in zip(step.dreg_levels, re_):
        assert torch.equal(a, b.grad)
    cd[0].mul_(0.5)                      # new data in the static inputs
    step.replay()
    with torch.no_grad():
        l2 = f([anchors, rd, cd], [gbd, gcd])
    assert step.loss.item() == l2.item() and step.loss.item() != le.item()


def test_predictor_levels_no_detections_and_bad_shapes():
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    from neuralnetworklibrary_b200.vision import level_shapes
    H, W, C, B = 96, 128, 8, 2
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    cl = [torch.full((B,) + shp, 0.01, device=dev()) for shp in level_shapes(H, W, 9, C)]
    rl = [torch.zeros((B,) + shp, device=dev()) for shp in level_shapes(H, W, 9, 4)]
    img = torch.zeros(B, 3, H, W, device=dev())
    boxes, classes, scores = BBoxPredictor()(img, rl, cl, anchors)          # nothing over the threshold
    assert boxes == [[], []] and classes == [[], []] and scores == [[], []]
    bad = list(cl)
    bad[1] = bad[1][:, :, :, :-1].contiguous()
    with pytest.raises(ValueError):
        BBoxPredictor()(img, rl, bad, anchors)
    with pytest.raises(ValueError):
        BBoxPredictor()(img, rl, cl, anchors.clone())                      # a copy carries no geometry tag
