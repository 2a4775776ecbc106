"""Developer micro-benchmarks (not the contract bench): times selected pieces of bench.py in isolation.
usage: python profiles/dev_bench.py levels|levels_logits|pascal|postproc|logits [steps] [option=value ...]   (rn_set_option)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

what = sys.argv[1].split(",") if len(sys.argv) > 1 else ["levels"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
from neuralnetworklibrary_b200 import _lib  # noqa: E402
for kv in sys.argv[3:]:
    _k, _v = kv.split("=")
    _lib.check(_lib.load().rn_set_option(_k.encode(), int(_v)))
torch.cuda.set_device(dev)
peak = bench.peaks()[0]
out = {}
COCO, PASCAL = bench.COCO, bench.PASCAL
A = 201600
for w in what:
    if w in ("levels", "levels_logits"):
        ms, _, _ = bench.time_loss_levels(COCO, 16, steps, 5, dev, w == "levels_logits")
        out[w] = {"ms": ms / steps, "frac": bench.loss_bytes(16, A, 80) / (ms / steps * 1e-3) / 1e9 / peak}
    elif w in ("pascal_levels",):
        ms, _, _ = bench.time_loss_levels(PASCAL, 32, steps, 5, dev, False)
        out[w] = {"ms": ms / steps, "frac": bench.loss_bytes(32, 49104, 20) / (ms / steps * 1e-3) / 1e9 / peak}
    elif w == "pascal":
        an, sets = bench.make_loss_sets(PASCAL, 32, dev, 4, 1003)
        ms, _, _ = bench.time_loss_graph(an, sets, max(steps, 40), 5, dev, 1)
        n = max(steps, 40)
        out[w] = {"ms": ms / n, "frac": bench.loss_bytes(32, an.shape[0], 20) / (ms / n * 1e-3) / 1e9 / peak}
    elif w in ("coco", "logits"):
        an, sets = bench.make_loss_sets(COCO, 16, dev, 2, 1005, logits=(w == "logits"))
        ms, _, _ = bench.time_loss_graph(an, sets, steps, 5, dev, 1, from_logits=(w == "logits"))
        out[w] = {"ms": ms / steps, "frac": bench.loss_bytes(16, A, 80) / (ms / steps * 1e-3) / 1e9 / peak}
    elif w in ("postproc_levels", "postproc_levels_logits"):
        ms, lay, ncand, nkept = bench.time_postproc_levels(COCO, 64, steps, 3, dev, w.endswith("logits"))
        out[w] = {"ms": ms, "layout_ops_ms": lay, "cand": ncand, "kept": nkept,
                  "frac": bench.loss_bytes(64, A, 80, grad=False) / (ms * 1e-3) / 1e9 / peak}
    elif w == "aux":
        out[w] = {k: v for k, v in bench.time_aux_kernels(dev, peak).items() if "stage_images" in k or "overlaps" in k}
    elif w == "postproc":
        pms, pwall, ncand, nkept, _ = bench.time_postproc(COCO, 64, steps, 3, dev)
        out[w] = {"ms": pms / steps, "wall_ms": pwall / steps, "frac": bench.loss_bytes(64, A, 80, grad=False) / (pms / steps * 1e-3) / 1e9 / peak}
print(" ".join(sys.argv[3:]), json.dumps(out))
