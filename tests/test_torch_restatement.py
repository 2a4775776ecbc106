"""The torch restatement of the reference (tests/torch_restatement.py) as a second checker.

* build container: pinned against the UNMODIFIED reference on CPU, bit for bit (skipped where /root/reference is absent);
* anywhere: against the reference-generated golden fixtures;
* GPU box: run with torch's CUDA kernels ("the reference running on CUDA") against this library's
  kernels -- assignments, candidate counts, keep indices, classes, scores AND decoded box coordinates bit-exact (both
  sides evaluate expf with the device libm), loss / gradients rtol 1e-5."""
import os

import numpy as np
import pytest
import torch

from tests import synth as syn
from oracle import oracle as orc
from tests import ref_runner as ref
from tests import torch_restatement as tr

RTOL = 1e-5


def _loss_inputs(seed, H, W, C, B, M):
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=12.0, max_frac=0.6)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=seed, edge_cases=64)
    return an, gb, gc, clas, reg


def _restated_loss(an, clas, reg, gb, gc, device="cpu", **kw):
    cl = clas.clone().to(device).requires_grad_(True)
    rg = reg.clone().to(device).requires_grad_(True)
    loss, reg_loss, clas_loss, matches = tr.ssd_loss(torch.as_tensor(an).to(device), rg, cl, gb.to(device), gc.to(device), **kw)
    loss.backward()
    dreg = rg.grad if rg.grad is not None else torch.zeros_like(rg)
    return (np.array([loss.item(), reg_loss.item(), clas_loss.item()], np.float32), cl.grad.cpu().numpy(),
            dreg.cpu().numpy(), matches.cpu().numpy())


@pytest.mark.skipif(not ref.available(), reason="reference checkout not present")
@pytest.mark.parametrize("seed,H,W,C,B,M,kw", [(41, 128, 128, 20, 2, 5, {}), (42, 96, 160, 8, 3, 12, dict(beta=0.3, alpha=0.4, gamma=1.5))])
def test_restated_loss_equals_reference_bitwise(seed, H, W, C, B, M, kw):
    an, gb, gc, clas, reg = _loss_inputs(seed, H, W, C, B, M)
    out3, dclas, dreg, matches = _restated_loss(an, clas, reg, gb, gc, **kw)
    r = ref.loss(an, clas, reg, gb, gc, **kw)
    assert np.array_equal(out3, r["out3"])
    assert np.array_equal(dclas, r["dclas"]) and np.array_equal(dreg, r["dreg"])
    for i in range(B):
        v = gc[i] >= 0
        pos, neg, mt = ref.assign(an, gb[i][v].numpy(), gc[i][v].numpy())
        assert np.array_equal(np.nonzero(matches[i] >= 0)[0], pos) and np.array_equal(np.nonzero(matches[i] == -1)[0], neg)
        assert np.array_equal(np.where(matches[i] >= 0, matches[i], -1), mt)


@pytest.mark.skipif(not ref.available(), reason="reference checkout not present")
@pytest.mark.parametrize("seed,kw", [(51, {}), (52, dict(thresh=0.2, max_overlap=0.3, top_k=50, max_boxes=7))])
def test_restated_predict_equals_reference_bitwise(seed, kw):
    H, W, C, B = 128, 160, 12, 2
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=-5.0, clusters=6)
    mine = tr.predict(H, W, reg, clas, torch.as_tensor(an), **kw)
    rb, rc, rs = ref.postproc(clas, reg, an, H, W, **kw)
    for i in range(B):
        assert len(rb[i]) == len(mine[i]["boxes"])
        if len(rb[i]):
            assert np.array_equal(np.stack(rb[i]), mine[i]["boxes"])
            assert np.array_equal(np.array(rc[i]), mine[i]["classes"]) and np.array_equal(np.array(rs[i], np.float32), mine[i]["scores"])


def test_restated_loss_vs_golden(golden_dir):
    """BASELINE.json configs[0] against the reference-generated fixture (runs on any host; 1e-6 covers a different
    vector libm on another CPU)."""
    g = np.load(os.path.join(golden_dir, "loss_cfg1.npz"))
    H, W, C, B, M = (int(g[k]) for k in "HWCBM")
    an = orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=int(g["seed"]))
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=int(g["seed"]))
    if not np.array_equal(gb.numpy(), g["gt_boxes"]):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    out3, dclas, dreg, matches = _restated_loss(an, clas, reg, gb, gc)
    assert np.array_equal(matches.astype(np.int8), g["matches"])
    np.testing.assert_allclose(out3, g["out3"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(dclas.reshape(-1, C)[g["pos_rows"]], g["pos_dclas"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(dreg.reshape(-1, 4)[g["pos_rows"]], g["pos_dreg"], rtol=1e-6, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,H,W,C,B,M,kw", [(61, 128, 160, 20, 2, 6, {}), (62, 200, 336, 80, 2, 10, dict(beta=0.3, alpha=0.4)),
                                               (63, 512, 512, 20, 2, 10, {})])
def test_kernels_vs_restatement_on_cuda_loss(seed, H, W, C, B, M, kw):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    from neuralnetworklibrary_b200.vision import SSD_loss
    dev = torch.device("cuda:0")
    an, gb, gc, clas, reg = _loss_inputs(seed, H, W, C, B, M)
    gb[0, 1] = gb[0, 0]   # duplicated ground-truth box: torch.max on CUDA must also pick the first index
    out3, dclas, dreg, matches = _restated_loss(an, clas, reg, gb, gc, device=dev, **kw)
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
    cd, rd = clas.to(dev).requires_grad_(True), reg.to(dev).requires_grad_(True)
    f = SSD_loss(keep_matches=True, **kw)
    loss = f([anchors, rd, cd], [gb.to(dev), gc.to(dev)])
    loss.backward()
    got_m, _ = f.last_assignment
    assert np.array_equal(got_m.cpu().numpy().astype(np.int64), matches)
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got3, out3, rtol=RTOL, atol=0)
    syn.assert_rel(cd.grad.cpu().numpy(), dclas, what="dclas")
    # both sides evaluate logf with the device libm here, so the scaled tolerance that the host-libm C oracle needs does
    # not apply: PURE rtol 1e-5, atol 0, identical zero pattern
    syn.assert_rel(rd.grad.cpu().numpy(), dreg, what="dreg")


@pytest.mark.gpu
@pytest.mark.parametrize("seed,H,W,C,B,mu,kw", [(71, 128, 160, 20, 2, -5.0, {}), (72, 256, 320, 80, 3, -5.5, dict(thresh=0.1, max_overlap=0.4, top_k=300, max_boxes=50)),
                                                (73, 800, 1344, 80, 1, -6.0, {})])
def test_kernels_vs_restatement_on_cuda_predict(seed, H, W, C, B, mu, kw):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
    dev = torch.device("cuda:0")
    an = orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=seed, anchors=an, mu=mu, clusters=8)
    clas_d, reg_d = clas.to(dev), reg.to(dev)
    want = tr.predict(H, W, reg_d, clas_d, torch.as_tensor(an).to(dev), **kw)
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
    got = BBoxPredictor().predict_arrays(H, W, reg_d, clas_d, anchors, **kw)
    for i in range(B):
        n = len(want[i]["anchor_idx"])
        assert int(got["counts"][i]) == n and int(got["n_candidates"][i]) == want[i]["n_candidates"]
        assert np.array_equal(got["anchor_idx"][i, :n], want[i]["anchor_idx"])
        assert np.array_equal(got["classes"][i, :n], want[i]["classes"])
        assert np.array_equal(got["scores"][i, :n], want[i]["scores"])
        assert np.array_equal(got["boxes"][i, :n].view(np.uint32), want[i]["boxes"].view(np.uint32)), "decoded boxes differ bitwise"
