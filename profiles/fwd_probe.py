import sys, os, torch, json
sys.path.insert(0, os.getcwd())
import bench
from tests import synth as syn
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import SSD_loss, level_shapes, _launch_loss_levels, _launch_loss
dev = torch.device("cuda:0")
H, W, C, M, B = 800, 1344, 80, 20, 16
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
A = anchors.shape[0]
gb, gc = syn.make_targets(B, M, H, W, C, seed=3); gb, gc = gb.to(dev), gc.to(dev)
res = {}
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): g.replay()
    t1.record(); torch.cuda.synchronize()
    return round(t0.elapsed_time(t1) / n, 4)
for logits in (False, True):
    cfg = dict(alpha=0.25, gamma=2.0, beta=0.5, pos_thresh=0.5, neg_thresh=0.4, world_size=1, group=None, global_batch=B, from_logits=logits)
    g = torch.Generator(device=dev).manual_seed(1)
    clas = []
    for shp in level_shapes(H, W, 9, C):
        z = torch.randn((B,) + shp, generator=g, device=dev) - 4.6
        clas.append(z if logits else torch.sigmoid_(z))
    reg = [torch.randn((B,) + shp, generator=g, device=dev) * 0.5 for shp in level_shapes(H, W, 9, 4)]
    for grad in (False, True):
        res["levels_%s_%s" % ("logits" if logits else "probs", "fwdbwd" if grad else "fwd")] = timeit(lambda: _launch_loss_levels(anchors, reg, clas, gb, gc, cfg, grad))
    del clas, reg
    fc, fr = bench.device_activations(B, A, C, 5, dev, mu=-4.6, logits=logits)
    for grad in (False, True):
        res["flat_%s_%s" % ("logits" if logits else "probs", "fwdbwd" if grad else "fwd")] = timeit(lambda: _launch_loss(anchors, fr, fc, gb, gc, cfg, grad))
    del fc, fr
print(json.dumps(res))
