"""The optional pruning stages of nms() (rel_thresh / inc / dup, reference retinanet.py:612-695) run on the
host over the NMS survivors.  Checked on CPU against the reference-generated golden vectors, with the CPU
oracle providing the survivors that the GPU kernel provides in production (tests may use oracle/)."""
import os

import numpy as np
import pytest

from neuralnetworklibrary_b200.core import ARR, list_del, list_mult
from neuralnetworklibrary_b200.retinanet import _host_stages, get_anchor_set
from oracle import oracle as orc

VARIANTS = [("rel", dict(rel_thresh=[0.3, 0.6], max_boxes=1000)),
            ("inc", dict(inc=[0.9, [1, 3]], max_boxes=1000)),
            ("dup", dict(dup=[0.4, [(0, 1), (1, 0), (2, 3)]], max_boxes=1000)),
            ("rel_inc_dup", dict(rel_thresh=[0.2, 0.5], inc=[0.8, [2]], dup=[0.5, [(0, 1), (3, 4)]],
                                 top_k=2000, max_boxes=60))]


@pytest.mark.parametrize("variant,kw", VARIANTS)
def test_host_stages_match_reference(golden_dir, variant, kw):
    g = np.load(os.path.join(golden_dir, "nms_boxes.npz"))
    top_k = kw.get("top_k", 1000)
    keep = orc.nms(g["boxes"], g["classes"], g["scores"], max_overlap=kw.get("max_overlap", 0.5), top_k=top_k,
                   max_boxes=top_k)
    b, c, s = g["boxes"][keep], g["classes"][keep], g["scores"][keep]
    sel = _host_stages(b, c, s, kw.get("rel_thresh"), kw.get("inc"), kw.get("dup"))
    m = kw["max_boxes"]
    assert np.array_equal(s[sel][:m], g[variant + "_scores"])
    assert np.array_equal(c[sel][:m], g[variant + "_classes"])
    assert np.array_equal(b[sel][:m], g[variant + "_boxes"])


def test_anchor_set_matches_oracle_base():
    base = orc.base_anchors()
    a = get_anchor_set()
    assert a.dtype == np.float64 and a.shape == (9, 4)
    for l in range(5):
        assert np.array_equal(2.0 ** (l + 5) * a, base[l])     # sizes 32..512, reference retinanet.py:480


def test_core_helpers():
    import torch
    assert list_del([0, 1, 2, 3, 4], [1, 3, 3]) == [0, 2, 4]
    assert list_mult([1.0, 2.0], 2) == [2.0, 4.0] and list_mult(3, 2) == 6
    assert ARR(torch.tensor([1.5])).dtype == np.float32
