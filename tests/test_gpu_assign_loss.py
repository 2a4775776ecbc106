"""GPU parity: anchors, assignment and the fused loss (forward + backward) through the drop-in wrappers
(-> C ABI -> sm_100a kernels) against the CPU oracle and the reference-generated golden fixtures.
Bit-exact for anchors / assignments; rtol 1e-5 (atol 0, identical zero patterns) for loss and gradients."""
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


rel_check = syn.assert_rel
dreg_check = syn.assert_dreg_close


def make_anchors(H, W, table_mode=False):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    a = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))
    if table_mode:
        a = a.clone()   # a copy carries no geometry tag -> kernels read the table
    return a


@pytest.mark.parametrize("H,W", [(64, 64), (100, 167), (33, 47), (512, 512), (800, 1333), (800, 1344), (608, 1024)])
def test_anchor_table_bitwise(H, W):
    a = make_anchors(H, W).cpu().numpy()
    want = orc.anchors(H, W)
    assert a.shape == want.shape and a.dtype == np.float32
    assert np.array_equal(a.view(np.uint32), want.view(np.uint32))


def test_anchor_table_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "anchors.npz"))
    assert np.array_equal(make_anchors(64, 64).cpu().numpy(), g["full_64x64"])
    assert np.array_equal(make_anchors(100, 167).cpu().numpy(), g["full_100x167"])


def test_anchor_generator_custom_set():
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    ratios, scales = [0.4, 1.0, 2.5, 3.0], [1.0, 1.3]
    a = AnchorGenerator(ratios, scales)(torch.zeros(2, 3, 96, 80, device=dev())).cpu().numpy()
    assert np.array_equal(a, orc.anchors(96, 80, ratios, scales))


ASSIGN_CASES = [  # seed, H, W, C, B, M
    (1001, 512, 512, 20, 2, 10),
    (31, 100, 167, 20, 3, 7),
    (32, 128, 128, 80, 4, 1),
    (33, 256, 320, 80, 2, 100),
    (34, 64, 64, 5, 5, 33),
]


@pytest.mark.parametrize("table_mode", [False, True])
@pytest.mark.parametrize("seed,H,W,C,B,M", ASSIGN_CASES)
def test_assign_bitexact(seed, H, W, C, B, M, table_mode):
    from neuralnetworklibrary_b200.vision import assign_batch
    anchors = make_anchors(H, W, table_mode)
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=10.0, max_frac=0.7)
    if M >= 3 and B >= 2:   # padding in the middle + duplicated box (argmax tie -> first index)
        gc[1, 0] = -1
        gb[1, 0] = -1
        gb[0, 2] = gb[0, 1]
    matches, npos, miou = assign_batch(anchors, gb.to(dev()), gc.to(dev()), want_iou=True)
    matches, npos, miou = matches.cpu().numpy(), npos.cpu().numpy(), miou.cpu().numpy()
    for i in range(B):
        m, n, iou = orc.assign(an, gb[i].numpy(), gc[i].numpy())
        assert np.array_equal(matches[i], m)
        assert npos[i] == n
        assert np.array_equal(miou[i].view(np.uint32), iou.view(np.uint32))


def test_assign_edge_cases():
    from neuralnetworklibrary_b200.vision import assign_batch, match_anchors_objects
    anchors = make_anchors(64, 64)
    an = anchors.cpu().numpy()
    A = an.shape[0]
    # all padding / M = 1 padding only: everything negative
    gb = -torch.ones(2, 4, 4)
    gc = -torch.ones(2, 4, dtype=torch.int64)
    m, npos, _ = assign_batch(anchors, gb.to(dev()), gc.to(dev()))
    assert (m.cpu().numpy() == -1).all() and (npos.cpu().numpy() == 0).all()
    # GT equal to an anchor, duplicated GT, zero-area GT; thresholds strict on both sides
    k = 3 + 9 * 5
    objs = torch.tensor(np.stack([an[k], an[k], an[100], [10, 10, 10, 30]]))
    pos, neg, mt = match_anchors_objects(objs.to(dev()), anchors)
    mo, _, miou = orc.assign(an, objs.numpy(), np.zeros(4, np.int64))
    assert mt[k].item() == 0 and mt[100].item() == 2
    assert np.array_equal(pos.cpu().numpy(), np.nonzero(mo >= 0)[0])
    assert np.array_equal(neg.cpu().numpy(), np.nonzero(mo == -1)[0])
    assert np.array_equal(mt.cpu().numpy(), np.where(mo >= 0, mo, -1))
    # no objects at all (reference Vision.py:1498-1501)
    pos, neg, mt = match_anchors_objects(torch.zeros(0, 4), anchors)
    assert pos.numel() == 0 and neg.numel() == A and (mt == -1).all()
    # exact-threshold IoUs with a one-anchor table
    box = torch.tensor([[0., 0., 10., 10.]], device=dev())
    for other, expect in (([0., 0., 10., 5.], -2), ([0., 0., 10., 4.], -2), ([0., 0., 10., 5.001], 0), ([0., 0., 10., 3.999], -1)):
        m, _, _ = assign_batch(box, torch.tensor([[other]], device=dev()), torch.zeros(1, 1, dtype=torch.int64, device=dev()))
        assert m.item() == expect, (other, m.item())


def test_max_overlaps_metric():
    from neuralnetworklibrary_b200.vision import ComputeMaxOverlaps
    H, W, B, M = 128, 160, 3, 6
    anchors = make_anchors(H, W)
    an = anchors.cpu().numpy()
    gb, gc = syn.make_targets(B, M, H, W, 20, seed=41, min_side=10.0, max_frac=0.7)
    metric = ComputeMaxOverlaps()
    val = metric([anchors], [gb.to(dev()), gc.to(dev())])
    means, allv = [], []
    for i in range(B):
        v = gc[i] >= 0
        if v.sum() == 0:
            continue
        best = []
        for j in np.nonzero(v.numpy())[0]:
            _, _, iou = orc.assign(an, gb[i, j:j + 1].numpy(), np.zeros(1, np.int64))
            best.append(iou.max())
        means.append(np.mean(np.array(best, np.float32)))
        allv += best
    assert np.array_equal(np.array(metric.max_overlaps, np.float32), np.array(allv, np.float32))
    np.testing.assert_allclose(val.item(), np.array(means).mean(), rtol=1e-6)


def run_loss(anchors, clas, reg, gb, gc, upstream=None, **kw):
    from neuralnetworklibrary_b200.vision import SSD_loss
    f = SSD_loss(keep_matches=True, **kw)   # the assignment checked below is the fused step's own output
    cd = clas.to(dev()).requires_grad_(True)
    rd = reg.to(dev()).requires_grad_(True)
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.requires_grad
    (loss if upstream is None else loss * upstream).backward()
    out3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    matches, npos = f.last_assignment
    return out3, cd.grad.cpu().numpy(), rd.grad.cpu().numpy(), matches.cpu().numpy(), npos.cpu().numpy()


LOSS_CASES = [  # seed, H, W, C, B, M, kwargs
    (1001, 512, 512, 20, 2, 10, {}),                        # BASELINE configs[0]
    (51, 128, 160, 80, 3, 8, {}),                           # C=80 vector path
    (52, 100, 167, 20, 2, 5, dict(beta=0.3, alpha=0.4)),
    (53, 96, 96, 7, 3, 4, {}),                              # C % 4 != 0 scalar path
    (54, 128, 128, 12, 2, 6, {}),                           # generic vector path
    (55, 128, 128, 20, 2, 6, dict(gamma=1.5)),              # general gamma
    (56, 64, 64, 80, 4, 3, dict(gamma=3.0, alpha=0.5)),
]


@pytest.mark.parametrize("table_mode", [False, True])
@pytest.mark.parametrize("seed,H,W,C,B,M,kw", LOSS_CASES)
def test_loss_forward_backward(seed, H, W, C, B, M, kw, table_mode):
    anchors = make_anchors(H, W, table_mode)
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=10.0, max_frac=0.7)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=seed, edge_cases=128)
    out3, dclas, dreg, matches, npos = run_loss(anchors, clas, reg, gb, gc, **kw)
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True, **kw)
    assert np.array_equal(matches, o["matches"]) and np.array_equal(npos, o["npos"])
    np.testing.assert_allclose(out3, o["out3"], rtol=RTOL, atol=0)
    rel_check(dclas, o["dclas"])
    dreg_check(dreg, o["dreg"])


def test_loss_golden_small(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_small.npz"))
    anchors = make_anchors(int(g["H"]), int(g["W"]))
    for variant, kw in (("default", {}), ("beta03_alpha04", dict(beta=0.3, alpha=0.4))):
        out3, dclas, dreg, matches, _ = run_loss(anchors, torch.from_numpy(g["clas"]), torch.from_numpy(g["reg"]),
                                                 torch.from_numpy(g["gt_boxes"]), torch.from_numpy(g["gt_cats"]), **kw)
        assert np.array_equal(matches, g["matches"])
        np.testing.assert_allclose(out3, g[variant + "_out3"], rtol=RTOL, atol=0)
        rel_check(dclas, g[variant + "_dclas"])
        dreg_check(dreg, g[variant + "_dreg"])


def test_loss_upstream_gradient_and_no_grad():
    from neuralnetworklibrary_b200.vision import SSD_ClasLoss, SSD_RegLoss, SSD_loss
    H, W, C, B, M = 96, 128, 20, 2, 5
    anchors = make_anchors(H, W)
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=61, min_side=10.0, max_frac=0.7)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=61)
    base = run_loss(anchors, clas, reg, gb, gc)
    scaled = run_loss(anchors, clas, reg, gb, gc, upstream=3.0)
    np.testing.assert_allclose(scaled[1], base[1] * np.float32(3.0), rtol=1e-6, atol=0)
    np.testing.assert_allclose(scaled[2], base[2] * np.float32(3.0), rtol=1e-6, atol=0)
    # forward only (the reference's evaluate path runs under no_grad)
    f = SSD_loss()
    with torch.no_grad():
        loss = f([anchors, reg.to(dev()), clas.to(dev())], [gb.to(dev()), gc.to(dev())])
    assert not loss.requires_grad
    np.testing.assert_allclose(loss.item(), base[0][0], rtol=1e-6)
    assert SSD_RegLoss(f)(None, None).item() == base[0][1] and SSD_ClasLoss(f)(None, None).item() == base[0][2]
    # determinism: bit-identical across repeated calls
    again = run_loss(anchors, clas, reg, gb, gc)
    assert np.array_equal(again[0], base[0]) and np.array_equal(again[1], base[1]) and np.array_equal(again[2], base[2])


def test_loss_sharded_partials_add_up():
    """Image shards with B_global = full batch give additive loss shares and identical per-image grads."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B, M = 96, 128, 20, 4, 5
    anchors = make_anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=62, min_side=10.0, max_frac=0.7)
    clas, reg = syn.make_train_activations(B, anchors.shape[0], C, seed=62)
    full = run_loss(anchors, clas, reg, gb, gc)
    parts, grads = [], []
    for sl in (slice(0, 3), slice(3, 4)):   # uneven shards: remainder to the low rank
        f = SSD_loss(global_batch=B)
        cd = clas[sl].to(dev()).requires_grad_(True)
        rd = reg[sl].to(dev()).requires_grad_(True)
        loss = f([anchors, rd, cd], [gb[sl].to(dev()), gc[sl].to(dev())])
        loss.backward()
        parts.append(np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float64))
        grads.append((cd.grad.cpu().numpy(), rd.grad.cpu().numpy()))
    np.testing.assert_allclose(parts[0] + parts[1], full[0], rtol=1e-6)
    assert np.array_equal(np.concatenate([g[0] for g in grads]), full[1])
    assert np.array_equal(np.concatenate([g[1] for g in grads]), full[2])


@pytest.mark.parametrize("H,W,C,B,M,seed", [(800, 1344, 80, 2, 20, 1002), (512, 512, 20, 4, 10, 1003), (800, 1333, 80, 2, 20, 1006)])
def test_loss_full_size(H, W, C, B, M, seed):
    """BASELINE.json shapes (COCO 800x1344 and, as named, 800x1333 = 200 700 anchors / Pascal 512x512) at a batch the oracle finishes in seconds."""
    anchors = make_anchors(H, W)
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=seed)
    out3, dclas, dreg, matches, npos = run_loss(anchors, clas, reg, gb, gc)
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True)
    assert np.array_equal(matches, o["matches"]) and np.array_equal(npos, o["npos"])
    np.testing.assert_allclose(out3, o["out3"], rtol=RTOL, atol=0)
    rel_check(dclas, o["dclas"])
    dreg_check(dreg, o["dreg"])


def test_captured_step_matches_eager():
    """SSD_loss.capture(): the CUDA-graph replay gives bit-identical results to the call-by-call path,
    also after new data is copied into the static input tensors."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B, M = 128, 160, 80, 3, 6
    anchors = make_anchors(H, W)
    A = anchors.shape[0]
    data = []
    for seed in (81, 82):
        gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=10.0, max_frac=0.7)
        clas, reg = syn.make_train_activations(B, A, C, seed=seed)
        data.append((clas, reg, gb, gc))
    clas_s, reg_s = data[0][0].to(dev()), data[0][1].to(dev())
    gb_s, gc_s = data[0][2].to(dev()), data[0][3].to(dev())
    cap = SSD_loss(keep_matches=True).capture([anchors, reg_s, clas_s], [gb_s, gc_s])
    for clas, reg, gb, gc in data:
        clas_s.copy_(clas.to(dev()))
        reg_s.copy_(reg.to(dev()))
        gb_s.copy_(gb.to(dev()))
        gc_s.copy_(gc.to(dev()))
        cap.replay()
        torch.cuda.synchronize()
        out3, dclas, dreg, matches, npos = run_loss(anchors, clas, reg, gb, gc)
        assert np.array_equal(cap.out3.cpu().numpy(), out3)
        assert np.array_equal(cap.dclas.cpu().numpy(), dclas) and np.array_equal(cap.dreg.cpu().numpy(), dreg)
        assert np.array_equal(cap.matches.cpu().numpy(), matches) and np.array_equal(cap.npos.cpu().numpy(), npos)


@pytest.mark.parametrize("seed,H,W,C,B,M,kw", [(91, 128, 160, 80, 2, 6, {}), (92, 100, 167, 20, 3, 5, dict(beta=0.3, alpha=0.4)),
                                               (93, 96, 96, 7, 2, 4, {}), (94, 64, 96, 12, 2, 4, dict(gamma=1.5))])
def test_loss_from_logits(seed, H, W, C, B, M, kw):
    """the head's sigmoid fused into the loss kernel (rn_loss_logits).

    The reference's focal term contains 1-(1-p), quantised to ulp(1): one ulp of p = sigmoid(z) (CUDA expf vs
    the host libm) moves it by up to 6e-4 relative at p = 1e-4.  The check therefore feeds the oracle the
    kernel's own probabilities (probs_out) and chains sigmoid's backward, grad*(1-y)*y, in fp32; the fused
    sigmoid itself must be within 6 ulp of the correctly rounded value."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    anchors, an = make_anchors(H, W), orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=10.0, max_frac=0.7)
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, an.shape[0], C, generator=g) * 1.5 - 4.6
    logits.view(-1)[:64] = torch.linspace(-12, 12, 64)       # both clamp ends and the ill-conditioned large-p side
    reg = torch.randn(B, an.shape[0], 4, generator=g) * 0.5
    f = SSD_loss(from_logits=True, keep_probs=True, **kw)
    zd, rd = logits.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
    loss = f([anchors, rd, zd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    y = f.last_probs.cpu().numpy()
    exact = 1.0 / (1.0 + np.exp(-logits.numpy().astype(np.float64)))
    ulp = np.abs(y.view(np.int32).astype(np.int64) - exact.astype(np.float32).view(np.int32).astype(np.int64))
    assert ulp.max() <= 6      # MUFU.EX2 (2 ulp) + product residual + add + reciprocal
    o = orc.loss(an, y, reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True, **kw)
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got3, o["out3"], rtol=RTOL, atol=0)
    rel_check(zd.grad.cpu().numpy(), (o["dclas"] * (np.float32(1) - y)) * y)
    dreg_check(rd.grad.cpu().numpy(), o["dreg"])
    # and the probability path on the same probabilities gives the identical loss value
    f2 = SSD_loss(**kw)
    with torch.no_grad():
        l2 = f2([anchors, reg.to(dev()), f.last_probs], [gb.to(dev()), gc.to(dev())])
    assert l2.item() == loss.item()
