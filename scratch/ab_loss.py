import sys, os, time
sys.path.insert(0, '/root/repo')
import torch
import bench
from neuralnetworklibrary_b200.vision import SSD_loss
dev = torch.device('cuda:0')
anchors, sets = bench.make_loss_sets(bench.COCO, 16, dev, 2, 1002)
def run(tag, pad_bytes=0):
    pads = []
    caps = []
    for clas, reg, gb, gc in sets:
        if pad_bytes: pads.append(torch.empty(pad_bytes, dtype=torch.uint8, device=dev))
        caps.append(SSD_loss().capture([anchors, reg, clas], [gb, gc]))
    for k in range(5): caps[k % 2].replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(40): caps[k % 2].replay()
    t1.record(); torch.cuda.synchronize()
    c = caps[0]
    print('%s pad=%d: %.4f ms/step  clas@%x dclas@%x diff=%d' % (tag, pad_bytes, t0.elapsed_time(t1) / 40, sets[0][0].data_ptr(), c.dclas.data_ptr(), c.dclas.data_ptr() - sets[0][0].data_ptr()))
    del caps, pads
tag = os.environ.get('RETINA_B200_LIB', 'new')
for pad in (0, 0, 1 << 20, 3 << 20, 17 << 20, 0):
    run(tag, pad)
