"""Times the fused training step (rn_loss_step) on synthetic COCO / Pascal shaped batches; library options can be set from
the command line (name=value ...).  Used for A/B runs inside one gpurun call and as the ncu target.
    python profiles/step_probe.py [coco|pascal|b256][:B] [reps] [opt=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralnetworklibrary_b200 import _lib
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import SSD_loss
from tests import synth as syn

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "coco"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lib = _lib.load()
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        _lib.check(lib.rn_set_option(k.encode(), int(v)))
    which, _, bsz = which.partition(":")   # e.g. pascal:37 = the Pascal shape with 37 images
    cfg = dict(coco=(800, 1344, 80, 20, 16), pascal=(512, 512, 20, 10, 32), b256=(800, 1344, 80, 20, 256))[which]
    H, W, C, M, B = cfg
    B = int(bsz) if bsz else B
    dev = torch.device("cuda:0")
    anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
    A = anchors.shape[0]
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for k in range(2 if which != "b256" else 1):
        clas = torch.sigmoid(torch.randn((B, A, C), generator=g, device=dev) - 4.6)
        reg = torch.randn((B, A, 4), generator=g, device=dev) * 0.5
        gb, gc = syn.make_targets(B, M, H, W, C, seed=5 + k)
        sets.append((clas, reg, gb.to(dev), gc.to(dev)))
    f = SSD_loss()
    caps = [f.capture([anchors, r, c], [gb, gc]) for c, r, gb, gc in sets]
    for i in range(5):
        caps[i % len(caps)].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        caps[i % len(caps)].replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 8 * A * (C + 4) * B
    print("%s B=%d %s: %.4f ms/step  %.1f GB/s  frac %.3f  loss %.6f" % (which, B, " ".join(sys.argv[3:]), ms, alg / ms / 1e6, alg / ms / 1e6 / 6538.6, caps[0].loss.item()))

main()
