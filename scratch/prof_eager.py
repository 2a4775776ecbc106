import sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import torch
import bench
from neuralnetworklibrary_b200.vision import SSD_loss
dev = torch.device('cuda:0')
anchors, sets = bench.make_loss_sets(bench.COCO, 16, dev, 2, 1002)
loss_fn = SSD_loss()
leaves = [(c.detach().requires_grad_(True), r.detach().requires_grad_(True), gb, gc) for c, r, gb, gc in sets]
def step(k):
    clas, reg, gb, gc = leaves[k % 2]
    clas.grad = None; reg.grad = None
    loss = loss_fn([anchors, reg, clas], [gb, gc])
    loss.backward()
    return loss
for k in range(5): step(k)
torch.cuda.synchronize()
t = time.perf_counter()
for k in range(20): step(k)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('cpu issue time per step %.3f ms, incl sync %.3f ms' % ((t1 - t) / 20 * 1e3, (t2 - t) / 20 * 1e3))
pr = cProfile.Profile(); pr.enable()
for k in range(20): step(k)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
