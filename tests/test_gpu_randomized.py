"""Randomised parity sweep on the GPU: random image shapes (incl. non-multiples of 8 / 128), class counts,
batch sizes, ground-truth counts, thresholds and loss weights, CUDA path vs the CPU oracle.  Seeds are
fixed, so the sweep is reproducible."""
import numpy as np
import pytest
import torch

from tests import synth as syn
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _anchors(H, W):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    return AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))


@pytest.mark.parametrize("case", range(10))
def test_random_loss_configs(case):
    from neuralnetworklibrary_b200.vision import SSD_loss
    rng = np.random.RandomState(1000 + case)
    H, W = int(rng.randint(33, 300)), int(rng.randint(33, 300))
    C = int(rng.choice([1, 3, 4, 8, 20, 21, 80, 91]))
    B, M = int(rng.randint(1, 5)), int(rng.randint(1, 40))
    kw = dict(beta=float(rng.uniform(0.1, 0.9)), alpha=float(rng.uniform(0.1, 0.9)),
              gamma=float(rng.choice([2.0, 2.0, 1.0, 0.5, 2.5])))
    anchors, an = _anchors(H, W), orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=case, force_empty_and_full=bool(rng.randint(2)), min_side=6.0,
                              max_frac=0.9)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=case, mu=float(rng.uniform(-5, -1)), edge_cases=64)
    f = SSD_loss(keep_matches=bool(case % 2), **kw)   # odd cases: the step's own matches; even: rn_assign on demand
    cd, rd = clas.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
    loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
    loss.backward()
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True, **kw)
    matches, npos = f.last_assignment
    assert np.array_equal(matches.cpu().numpy(), o["matches"]) and np.array_equal(npos.cpu().numpy(), o["npos"])
    got3 = np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32)
    np.testing.assert_allclose(got3, o["out3"], rtol=syn.RTOL, atol=0)
    syn.assert_rel(cd.grad.cpu().numpy(), o["dclas"], what="dclas")
    syn.assert_dreg_close(rd.grad.cpu().numpy(), o["dreg"])


@pytest.mark.parametrize("case", range(10))
def test_random_postproc_configs(case):
    from neuralnetworklibrary_b200.retinanet import BBoxPredictor
    rng = np.random.RandomState(2000 + case)
    H, W = int(rng.randint(33, 300)), int(rng.randint(33, 300))
    C = int(rng.choice([1, 3, 4, 8, 20, 21, 80, 91]))
    B = int(rng.randint(1, 4))
    top_k = int(rng.choice([1, 7, 100, 1000, 4096]))
    kw = dict(thresh=float(rng.choice([0.0, 0.01, 0.05, 0.3])), max_overlap=float(rng.choice([0.0, 0.3, 0.5, 0.9])),
              top_k=top_k, max_boxes=int(min(top_k, rng.choice([1, 20, 300, 4096]))))
    mean = [float(v) for v in rng.uniform(-0.05, 0.05, 4)] if case % 2 else [0., 0., 0., 0.]
    std = [0.1, 0.1, 0.2, 0.2] if case % 3 else [0.2, 0.15, 0.1, 0.3]
    anchors, an = _anchors(H, W), orc.anchors(H, W)
    clas, reg = syn.make_infer_activations(B, an.shape[0], C, seed=case, anchors=an, mu=float(rng.uniform(-6, -2)),
                                           clusters=6)
    out = BBoxPredictor(mean, std).predict_arrays(H, W, reg.to(dev()), clas.to(dev()), anchors, **kw)
    po = orc.postproc(clas.numpy(), reg.numpy(), an, H, W, mean=mean, std=std, **kw)
    assert np.array_equal(out["counts"], po["counts"]) and np.array_equal(out["n_candidates"], po["n_candidates"])
    for i, n in enumerate(po["counts"]):
        assert np.array_equal(out["anchor_idx"][i, :n], po["anchor_idx"][i, :n])
        assert np.array_equal(out["classes"][i, :n], po["classes"][i, :n])
        assert np.array_equal(out["scores"][i, :n], po["scores"][i, :n])
        syn.assert_boxes_close(out["boxes"][i, :n], po["boxes"][i, :n])
