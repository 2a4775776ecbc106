"""OPT-IN BUILD (-DRN_EXPERIMENTAL; skipped otherwise).  The fused training step (rn_loss_step after rn_set_option("step_fused", 1) -> rn_step_kernel: assignment + loss
forward/backward + final reduction in one persistent launch) against the separate kernels of the same library (the default)
and against the CPU oracle.  Assignments and positive counts bit-exact, gradients BIT-IDENTICAL to the separate kernels (same
element arithmetic), the three loss scalars within 1e-6 of them (the partial sums are grouped differently) and rtol 1e-5 of
the oracle.  Also: the zero-initialised state buffer is left zeroed (self-cleaning byte map), calls of different shapes share it, many boxes, boxes far outside
the image, an image without objects in the middle of a batch, and the poisoned loss for a category >= C."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import synth as syn

def _experimental():
    from neuralnetworklibrary_b200 import _lib
    return _lib.has_experimental()


# the two alternative implementations are opt-in at BUILD time (they measured slower than the default chain):
#   RN_EXTRA_NVCC_FLAGS=-DRN_EXPERIMENTAL python -c "from neuralnetworklibrary_b200 import _lib; _lib.build_library(force=True)"
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not _experimental(), reason="library built without -DRN_EXPERIMENTAL (opt-in step variants)")]
RTOL = 1e-5


def dev():
    return torch.device("cuda:0")


def anchors_for(H, W):
    from neuralnetworklibrary_b200.retinanet import AnchorGenerator
    return AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev()))


def run(anchors, clas, reg, gb, gc, unfused=False, chain4=False, **kw):
    """unfused=False: the opt-in persistent kernel; unfused=True: the opt-in three-kernel byte-map chain; chain4=True: the
    default, the kernels of rn_assign + rn_loss (int32 matches)."""
    from neuralnetworklibrary_b200 import _lib
    from neuralnetworklibrary_b200.vision import SSD_loss
    with _lib.option("step_fused", 0 if (unfused or chain4) else 1), _lib.option("step_bytemap", 1 if (unfused and not chain4) else 0):
        f = SSD_loss(keep_matches=True, **kw)
        cd, rd = clas.to(dev()).requires_grad_(True), reg.to(dev()).requires_grad_(True)
        loss = f([anchors, rd, cd], [gb.to(dev()), gc.to(dev())])
        loss.backward()
        torch.cuda.synchronize()
        m, n = f.last_assignment
        return (np.array([loss.item(), f.reg_loss.item(), f.clas_loss.item()], np.float32), cd.grad.cpu().numpy(),
                rd.grad.cpu().numpy(), m.cpu().numpy(), n.cpu().numpy())


CASES = [  # seed, H, W, C, B, M, kwargs
    (501, 128, 160, 80, 3, 8, {}),
    (502, 100, 167, 20, 5, 10, dict(beta=0.3, alpha=0.4)),
    (503, 96, 96, 7, 3, 4, {}),                  # C % 4 != 0
    (504, 64, 96, 12, 2, 40, dict(gamma=1.5)),   # generic row width, general gamma, many slots
    (505, 512, 512, 20, 8, 10, {}),              # Pascal shape: slices span several images
    (506, 800, 1344, 80, 2, 20, {}),             # COCO shape
    (507, 33, 47, 4, 7, 3, {}),                  # tiny images: a CTA slice covers many images
    (508, 256, 320, 80, 2, 128, {}),             # the largest M of the fused step
]


@pytest.mark.parametrize("seed,H,W,C,B,M,kw", CASES)
def test_fused_equals_separate_kernels_and_oracle(seed, H, W, C, B, M, kw):
    anchors, an = anchors_for(H, W), orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=8.0, max_frac=0.8)
    if B >= 3:
        gc[1] = -1          # an image without objects in the middle of the batch
    gb[0, 0] = torch.tensor([-500.0, -400.0, -300.0, -200.0])   # a box far outside the image
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=seed, edge_cases=64)
    f3, fdc, fdr, fm, fn = run(anchors, clas, reg, gb, gc, **kw)
    u3, udc, udr, um, un = run(anchors, clas, reg, gb, gc, unfused=True, **kw)
    c3, cdc, cdr, cm, cn = run(anchors, clas, reg, gb, gc, chain4=True, **kw)
    assert np.array_equal(fm, um) and np.array_equal(fn, un) and np.array_equal(cm, um) and np.array_equal(cn, un)
    assert np.array_equal(fdc, udc), "dclas differs between the fused step and rn_loss"
    assert np.array_equal(fdr, udr), "dreg differs between the fused step and rn_loss"
    assert np.array_equal(cdc, udc) and np.array_equal(cdr, udr), "gradients differ between the byte-map and the int32 chain"
    assert np.array_equal(c3, u3), "the byte-map chain and the int32 chain run the same loss kernel on the same assignment"
    np.testing.assert_allclose(f3, u3, rtol=1e-6, atol=0)
    o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True, **kw)
    assert np.array_equal(fm, o["matches"]) and np.array_equal(fn, o["npos"])
    np.testing.assert_allclose(f3, o["out3"], rtol=RTOL, atol=0)
    syn.assert_rel(fdc, o["dclas"], what="dclas")
    syn.assert_dreg_close(fdr, o["dreg"])


def test_fused_step_is_repeatable_and_leaves_workspace_zeroed():
    """Back-to-back calls with different targets through the SAME workspace: every call equals a call through a fresh
    workspace bit for bit (run to run determinism), and the workspace's zero-initialised part is zero afterwards."""
    from neuralnetworklibrary_b200 import _lib, vision
    H, W, C, B, M = 160, 224, 20, 6, 12
    anchors, an = anchors_for(H, W), orc.anchors(H, W)
    A = an.shape[0]
    outs = []
    for rep in range(3):
        for seed in (601, 602, 603):
            gb, gc = syn.make_targets(B, M, H, W, C, seed=seed, min_side=8.0, max_frac=0.8)
            clas, reg = syn.make_train_activations(B, A, C, seed=seed)
            outs.append((seed, run(anchors, clas, reg, gb, gc, unfused=bool(rep % 2))))   # both users of the state buffer, interleaved
    first = {}
    for seed, o in outs:
        if seed in first:
            for k, (a, b) in enumerate(zip(first[seed], o)):
                if k == 0:   # the three scalars: the persistent kernel and the chain group their partial sums differently
                    np.testing.assert_allclose(a, b, rtol=1e-6, atol=0)
                else:
                    assert np.array_equal(a, b)
        else:
            first[seed] = o
    state = vision._step_state.get(_lib.load().rn_loss_step_state_bytes(B, A), dev())
    torch.cuda.synchronize()
    assert not state.any(), "the step left residue in its state buffer"


def test_fused_step_no_grad_and_logits():
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B, M = 128, 160, 80, 3, 6
    anchors, an = anchors_for(H, W), orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=611)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=611)
    from neuralnetworklibrary_b200 import _lib
    base = run(anchors, clas, reg, gb, gc)
    with _lib.option("step_fused", 1):
        with torch.no_grad():
            f = SSD_loss()
            loss = f([anchors, reg.to(dev()), clas.to(dev())], [gb.to(dev()), gc.to(dev())])
        assert loss.item() == base[0][0] and not loss.requires_grad
        # logits in: the loss value on sigmoid(logits) as the kernel computed them equals the probability path bit for bit
        logits = torch.logit(clas.clamp(1e-6, 1 - 1e-6))
        fl = SSD_loss(from_logits=True, keep_probs=True)
        ld = logits.to(dev()).requires_grad_(True)
        l1 = fl([anchors, reg.to(dev()), ld], [gb.to(dev()), gc.to(dev())])
        l1.backward()
        with torch.no_grad():
            l2 = SSD_loss()([anchors, reg.to(dev()), fl.last_probs], [gb.to(dev()), gc.to(dev())])
        assert l1.item() == l2.item() and torch.isfinite(ld.grad).all()


def test_category_out_of_range_poisons_the_loss():
    """The reference raises IndexError for a category >= C (Vision.py:1593); the fused kernel cannot raise, so the returned
    loss is NaN (loud) and the next call is clean again."""
    from neuralnetworklibrary_b200.vision import SSD_loss
    H, W, C, B, M = 96, 128, 8, 2, 4
    anchors, an = anchors_for(H, W), orc.anchors(H, W)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=621)
    clas, reg = syn.make_train_activations(B, an.shape[0], C, seed=621)
    from neuralnetworklibrary_b200 import _lib
    bad = gc.clone()
    bad[0, 0] = C
    with _lib.option("step_fused", 1), torch.no_grad():
        l_bad = SSD_loss()([anchors, reg.to(dev()), clas.to(dev())], [gb.to(dev()), bad.to(dev())])
        l_ok = SSD_loss()([anchors, reg.to(dev()), clas.to(dev())], [gb.to(dev()), gc.to(dev())])
    assert np.isnan(l_bad.item()) and np.isfinite(l_ok.item())
