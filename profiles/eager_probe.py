"""Host overhead of the eager drop-in API: SSD_loss()(...) + loss.backward() call by call against the CUDA-graph replay of the
same step, and a cProfile of the eager loop (where the Python time goes).
    python profiles/eager_probe.py [coco|pascal] [steps]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import SSD_loss
from tests import synth as syn

which = sys.argv[1] if len(sys.argv) > 1 else "pascal"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
H, W, C, M, B = dict(coco=(800, 1344, 80, 20, 16), pascal=(512, 512, 20, 10, 32))[which]
dev = torch.device("cuda:0")
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
A = anchors.shape[0]
g = torch.Generator(device=dev).manual_seed(1)
sets = []
for k in range(2):
    clas = torch.sigmoid(torch.randn((B, A, C), generator=g, device=dev) - 4.6).requires_grad_(True)
    reg = (torch.randn((B, A, 4), generator=g, device=dev) * 0.5).requires_grad_(True)
    gb, gc = syn.make_targets(B, M, H, W, C, seed=5 + k)
    sets.append((clas, reg, gb.to(dev), gc.to(dev)))
f = SSD_loss()

def step(k):
    clas, reg, gb, gc = sets[k % 2]
    clas.grad = None
    reg.grad = None
    loss = f([anchors, reg, clas], [gb, gc])
    loss.backward()
    return loss

def timed(fn, n):
    for k in range(10):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(n):
        fn(k)
    t_issue = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, t_issue / n * 1e6

eager_us, issue_us = timed(step, steps)
caps = [f.capture([anchors, r.detach(), c.detach()], [gb, gc]) for c, r, gb, gc in sets]
graph_us, gissue_us = timed(lambda k: caps[k % 2].replay(), steps)
print("%s: eager %.1f us/step on the device (host issues a step in %.1f us), graph replay %.1f us (host %.1f us): ratio %.3f"
      % (which, eager_us, issue_us, graph_us, gissue_us, eager_us / graph_us))
pr = cProfile.Profile()
pr.enable()
for k in range(steps):
    step(k)
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(18)
print("\n".join(l for l in out.getvalue().splitlines() if l.strip())[:6000])
