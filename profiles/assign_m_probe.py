"""Assignment time (rn_assign: fill + sparse kernel, or the dense kernel) against the number of ground-truth boxes per image,
COCO shape, B = 16; CUDA events, 20 launches.  `fill` is the fraction of the M slots that hold a real box."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from neuralnetworklibrary_b200 import _lib
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import assign_batch
dev = torch.device("cuda:0")
H, W, B = 800, 1344, 16
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
for M, fill in ((20, 1.0), (100, 0.07), (100, 0.5), (100, 1.0), (128, 1.0)):
    g = torch.Generator().manual_seed(1)
    wh = torch.rand(B, M, 2, generator=g) * 0.35 * min(H, W) + 16
    xy = torch.rand(B, M, 2, generator=g) * (torch.tensor([W, H]) - wh)
    gb = torch.cat([xy, xy + wh], -1).float(); gc = torch.randint(0, 80, (B, M))
    drop = torch.rand(B, M, generator=g) > fill
    gc[drop] = -1; gb[drop] = -1
    gb, gc = gb.to(dev), gc.to(dev)
    out = []
    for dense in (0, 1):
        with _lib.option("assign_dense", dense):
            for _ in range(3): assign_batch(anchors, gb, gc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): m, n, _ = assign_batch(anchors, gb, gc)
            e1.record(); torch.cuda.synchronize()
            out.append((e0.elapsed_time(e1) / 20 * 1e3, m.clone(), n.clone()))
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    print("M=%3d fill=%.2f (%.1f boxes/image): sparse %6.1f us   dense %6.1f us   (identical matches, %d positives)" %
          (M, fill, float((gc >= 0).sum()) / B, out[0][0], out[1][0], int(out[0][2].sum())))
