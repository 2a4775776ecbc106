import sys, os, torch
sys.path.insert(0, os.getcwd())
from tests import synth as syn
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import assign_batch
dev = torch.device("cuda:0")
H, W, B = 800, 1344, 16
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
for M, fill in ((100, 0.07), (100, 1.0), (128, 1.0), (20, 1.0)):
    g = torch.Generator().manual_seed(1)
    wh = torch.rand(B, M, 2, generator=g) * 0.35 * min(H, W) + 16
    xy = torch.rand(B, M, 2, generator=g) * (torch.tensor([W, H]) - wh)
    gb = torch.cat([xy, xy + wh], -1).float(); gc = torch.randint(0, 80, (B, M))
    drop = torch.rand(B, M, generator=g) > fill
    gc[drop] = -1; gb[drop] = -1
    gb, gc = gb.to(dev), gc.to(dev)
    for _ in range(3): assign_batch(anchors, gb, gc)
    torch.cuda.synchronize()
