"""Per-CTA %globaltimer stamps of rn_step_kernel.  Needs a library built with the opt-in step variants and -DRN_STEP_TIMING:
    RN_EXTRA_NVCC_FLAGS="-DRN_EXPERIMENTAL -DRN_STEP_TIMING" python -c "from neuralnetworklibrary_b200 import _lib; _lib.build_library(force=True)" """
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neuralnetworklibrary_b200 import _lib
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import SSD_loss
from tests import synth as syn
which = sys.argv[1] if len(sys.argv) > 1 else "coco"
H, W, C, M, B = dict(coco=(800, 1344, 80, 20, 16), pascal=(512, 512, 20, 10, 32))[which]
dev = torch.device("cuda:0")
lib = _lib.load()
anchors = AnchorGenerator()(torch.zeros(1, 3, H, W, device=dev))
A = anchors.shape[0]
g = torch.Generator(device=dev).manual_seed(1)
clas = torch.sigmoid(torch.randn((B, A, C), generator=g, device=dev) - 4.6)
reg = torch.randn((B, A, 4), generator=g, device=dev) * 0.5
gb, gc = syn.make_targets(B, M, H, W, C, seed=5)
gb, gc = gb.to(dev), gc.to(dev)
f = SSD_loss()
dbg = torch.zeros(2048 * 8, dtype=torch.int64, device=dev)
for i in range(3):
    cd, rd = clas.clone().requires_grad_(True), reg.clone().requires_grad_(True)
    f([anchors, rd, cd], [gb, gc])
import ctypes as C
raw = C.CDLL(_lib.LIB_PATH)
raw.rn_step_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
cd, rd = clas.clone().requires_grad_(True), reg.clone().requires_grad_(True)
f([anchors, rd, cd], [gb, gc])
torch.cuda.synchronize()
raw.rn_step_debug_buffer(None)
t = dbg.cpu().numpy().reshape(-1, 8).astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start", "phaseA done", "phaseB done", "final end"]
print(which, "CTAs", len(t))
for i, n in enumerate(names):
    col = t[:, i][t[:, i] > 0]
    if len(col):
        print("%-16s min %8.2f us  median %8.2f  max %8.2f   (n=%d)" % (n, (col.min() - t0) / 1e3, (np.median(col) - t0) / 1e3, (col.max() - t0) / 1e3, len(col)))
