"""Developer probe: rn_assign kernel time per ground-truth fill (run under ncu; prints nothing useful by itself).
usage: ncu --metrics gpu__time_duration.sum -k regex:rn_assign --csv python profiles/assign_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import synth as syn  # noqa: E402
from neuralnetworklibrary_b200.retinanet import AnchorGenerator  # noqa: E402
from neuralnetworklibrary_b200.vision import assign_batch  # noqa: E402

dev = torch.device("cuda:0")


def run(H, W, B, M, mode, C=80):
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
    gb, gc = syn.make_targets(B, M, H, W, C, seed=5)
    if mode == "empty":
        gc[:] = -1
        gb[:] = -1
    if mode == "full":
        g = torch.Generator().manual_seed(1)
        wh = torch.rand(B, M, 2, generator=g) * 0.35 * min(H, W) + 16
        xy = torch.rand(B, M, 2, generator=g) * (torch.tensor([W, H]) - wh)
        gb = torch.cat([xy, xy + wh], -1).float()
        gc = torch.randint(0, C, (B, M))
    gb, gc = gb.to(dev), gc.to(dev)
    for _ in range(3):
        assign_batch(anchors, gb, gc)
    torch.cuda.synchronize()


for name, (H, W, B, M) in {"coco": (800, 1344, 16, 20), "pascal": (512, 512, 32, 10)}.items():
    for mode in ("mixed", "empty", "full"):
        run(H, W, B, M, mode)
