import sys, time
sys.path.insert(0, '/root/repo')
import torch, bench
from neuralnetworklibrary_b200.retinanet import AnchorGenerator, BBoxPredictor
dev = torch.device('cuda:0')
H, W, C, B = 800, 1344, 80, 64
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev)); A = anchors.shape[0]
clas, reg = bench.device_activations(B, A, C, 1004, dev, mu=-6.0)
bp = BBoxPredictor()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter(); t0.record()
    for _ in range(n): fn()
    t1.record(); c1 = time.perf_counter(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n, (c1 - c0) / n * 1e3
print('eager predict_device: gpu %.4f ms, cpu issue %.4f ms' % timeit(lambda: bp.predict_device(H, W, reg, clas, anchors)))
# graph
bp.predict_device(H, W, reg, clas, anchors); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = bp.predict_device(H, W, reg, clas, anchors)
print('graph replay: gpu %.4f ms, cpu %.4f ms' % timeit(lambda: g.replay()))
# second input set to check L2 effects (inputs 4.3 GB >> L2 anyway)
