"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference is imported through oracle/ref_shim.py (CPU torch, `Tensor.cuda` no-op) and executed on
seeded synthetic inputs (tests/synth.py).  Inputs that
are cheap to store are stored next to the outputs, larger ones are re-generated from their seed and
pinned by a SHA-256 of their bytes.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import synth as syn  # noqa: E402
from tests import ref_runner as ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_anchors():
    d = {}
    for (H, W) in [(64, 64), (100, 167)]:
        d["full_%dx%d" % (H, W)] = ref.anchors(H, W)
    shapes = [(512, 512), (800, 1333), (800, 1344), (608, 1024), (608, 1216), (33, 47)]
    d["shapes"] = np.array(shapes, dtype=np.int64)
    d["counts"] = np.array([ref.anchors(H, W).shape[0] for (H, W) in shapes], dtype=np.int64)
    d["sha256"] = np.array([sha(ref.anchors(H, W)) for (H, W) in shapes])
    a = ref.anchors(512, 512)
    d["a512_head"] = a[:12]
    d["a512_tail"] = a[-3:]
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), **d)


def gen_loss_small():
    H = W = 128
    C, B, M = 20, 3, 6
    an = ref.anchors(H, W)
    A = an.shape[0]
    gb, gc = syn.make_targets(B, M, H, W, C, seed=2001, min_side=12.0, max_frac=0.6)
    gb[1, 2] = gb[1, 0]            # duplicated GT box: argmax tie -> first index
    gc[1, 2] = (gc[1, 0] + 1) % C
    clas, reg = syn.make_train_activations(B, A, C, seed=2001, edge_cases=64)
    variants = {}
    for name, kw in [("default", {}), ("beta03_alpha04", dict(beta=0.3, alpha=0.4))]:
        r = ref.loss(an, clas, reg, gb, gc, **kw)
        variants[name + "_out3"] = r["out3"]
        variants[name + "_dclas"] = r["dclas"]
        variants[name + "_dreg"] = r["dreg"]
    matches = []
    for i in range(B):
        v = gc[i] >= 0
        _, neg, m = ref.assign(an, gb[i][v], gc[i][v])
        mm = np.full(A, -2, dtype=np.int32)   # -2 = ignored (neither positive nor negative)
        mm[neg] = -1
        mm[m >= 0] = m[m >= 0]
        matches.append(mm)
    np.savez_compressed(os.path.join(OUT, "loss_small.npz"), H=H, W=W, C=C, B=B, M=M, anchors=an,
                        gt_boxes=gb.numpy(), gt_cats=gc.numpy(), clas=clas.numpy(), reg=reg.numpy(),
                        matches=np.stack(matches), **variants)


def gen_loss_cfg1():
    """BASELINE.json configs[0]: B=2, 512x512, 20 classes, <= 10 GT boxes / image."""
    H = W = 512
    C, B, M = 20, 2, 10
    an = ref.anchors(H, W)
    A = an.shape[0]
    gb, gc = syn.make_targets(B, M, H, W, C, seed=1001)
    clas, reg = syn.make_train_activations(B, A, C, seed=1001)
    r = ref.loss(an, clas, reg, gb, gc)
    matches = []
    for i in range(B):
        v = gc[i] >= 0
        _, neg, m = ref.assign(an, gb[i][v], gc[i][v])
        mm = np.full(A, -2, dtype=np.int8)
        mm[neg] = -1
        mm[m >= 0] = m[m >= 0]
        matches.append(mm)
    g = torch.Generator().manual_seed(7)
    sample = torch.randint(0, B * A * C, (20000,), generator=g).numpy()
    pos_rows = np.nonzero(np.stack(matches).reshape(-1) >= 0)[0]
    np.savez_compressed(os.path.join(OUT, "loss_cfg1.npz"), H=H, W=W, C=C, B=B, M=M, seed=1001,
                        sha_clas=sha(clas.numpy()), sha_reg=sha(reg.numpy()), sha_gt=sha(gb.numpy()),
                        gt_boxes=gb.numpy(), gt_cats=gc.numpy(), out3=r["out3"], matches=np.stack(matches),
                        sample_idx=sample, sample_dclas=r["dclas"].reshape(-1)[sample],
                        pos_rows=pos_rows, pos_dclas=r["dclas"].reshape(-1, C)[pos_rows],
                        pos_dreg=r["dreg"].reshape(-1, 4)[pos_rows],
                        dclas_abs_sum=np.abs(r["dclas"].astype(np.float64)).sum(),
                        dreg_nonzero_rows=np.count_nonzero(np.abs(r["dreg"]).sum(-1)))


def _pack(lists, K):
    B = len(lists[0])
    boxes = np.zeros((B, K, 4), np.float32)
    classes = np.zeros((B, K), np.int64)
    scores = np.zeros((B, K), np.float32)
    counts = np.zeros(B, np.int32)
    for i in range(B):
        n = len(lists[0][i])
        counts[i] = n
        if n:
            boxes[i, :n] = np.stack(lists[0][i])
            classes[i, :n] = np.array(lists[1][i])
            scores[i, :n] = np.array(lists[2][i])
    return boxes, classes, scores, counts


def gen_postproc_small():
    H, W = 128, 160
    C, B = 20, 3
    an = ref.anchors(H, W)
    A = an.shape[0]
    clas, reg = syn.make_infer_activations(B, A, C, seed=2004, anchors=an, mu=-5.0, clusters=6, per_cluster=25)
    clas[2] = clas[2] * 0.04       # image with no candidate above 0.05 -> empty lists
    d = dict(H=H, W=W, C=C, B=B, anchors=an, clas=clas.numpy(), reg=reg.numpy())
    variants = [("default", dict()),
                ("topk50_max100", dict(top_k=50, max_boxes=100)),
                ("thr02_ov03_max7", dict(thresh=0.2, max_overlap=0.3, max_boxes=7)),
                ("thr001_max1000", dict(thresh=0.01, max_boxes=1000))]
    for name, kw in variants:
        K = kw.get("max_boxes", 20)
        b, c, s, n = _pack(ref.postproc(clas, reg, an, H, W, **kw), K)
        d[name + "_boxes"], d[name + "_classes"], d[name + "_scores"], d[name + "_counts"] = b, c, s, n
    np.savez_compressed(os.path.join(OUT, "postproc_small.npz"), **d)


def gen_nms_boxes():
    g = torch.Generator().manual_seed(5)
    n = 3000
    xy = torch.rand(n, 2, generator=g) * 400
    wh = torch.rand(n, 2, generator=g) * 80 + 10
    boxes = torch.cat([xy, xy + wh], 1)
    classes = torch.randint(0, 5, (n,), generator=g)
    scores = torch.rand(n, generator=g)
    d = dict(boxes=boxes.numpy(), classes=classes.numpy(), scores=scores.numpy())
    variants = [("all", dict(top_k=3000, max_boxes=100000)),
                ("default", dict()),
                ("ov07_topk500_max50", dict(max_overlap=0.7, top_k=500, max_boxes=50)),
                ("rel", dict(rel_thresh=[0.3, 0.6], max_boxes=1000)),
                ("inc", dict(inc=[0.9, [1, 3]], max_boxes=1000)),
                ("dup", dict(dup=[0.4, [(0, 1), (1, 0), (2, 3)]], max_boxes=1000)),
                ("rel_inc_dup", dict(rel_thresh=[0.2, 0.5], inc=[0.8, [2]], dup=[0.5, [(0, 1), (3, 4)]],
                                     top_k=2000, max_boxes=60))]
    for name, kw in variants:
        rb, rc, rs = ref.nms(boxes, classes, scores, **kw)
        d[name + "_boxes"] = np.stack(rb) if len(rb) else np.zeros((0, 4), np.float32)
        d[name + "_classes"] = np.array(rc, dtype=np.int64)
        d[name + "_scores"] = np.array(rs, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "nms_boxes.npz"), **d)


def collater_cases():
    """Ragged ground truth for the collater fixture: (boxes int64 / float64, cats) per image, incl. empty images."""
    rng = np.random.RandomState(77)
    cases = []
    for bs, empty_all in ((4, False), (3, True), (5, False)):
        boxes, cats, scales = [], [], []
        for i in range(bs):
            n = 0 if (empty_all or i == 1) else int(rng.randint(1, 9))
            xy = rng.randint(0, 200, size=(n, 2))
            wh = rng.randint(5, 120, size=(n, 2))
            b = np.concatenate([xy, xy + wh], axis=1)
            boxes.append(b.astype(np.float64) if i % 2 else b.astype(np.int64))
            cats.append(rng.randint(0, 20, size=n).astype(np.int64))
            scales.append(float(rng.uniform(0.5, 2.0)))
        cases.append((boxes, cats, scales, float(rng.uniform(0.8, 1.25)), int(rng.randint(0, 17)), int(rng.randint(0, 17))))
    return cases


def gen_collater():
    """Target half of the reference's AspectRatioCollater (Vision.py:730-812) on tiny synthetic images."""
    from oracle import ref_shim
    _, vis = ref_shim.load()
    d = {}
    for k, (boxes, cats, scales, rand_scale, row_jit, col_jit) in enumerate(collater_cases()):
        batch = []
        for i in range(len(boxes)):
            img = np.zeros((8, 8, 3), dtype=np.float32)
            b = boxes[i].copy() if len(boxes[i]) else np.array([])
            batch.append((img, scales[i], rand_scale, row_jit, col_jit, b, cats[i].copy() if len(cats[i]) else np.array([])))
        _, (bp, cp) = vis.AspectRatioCollater(batch)
        d["case%d_boxes" % k] = bp.numpy()
        d["case%d_cats" % k] = cp.numpy()
    np.savez_compressed(os.path.join(OUT, "collater_targets.npz"), **d)


def image_cases():
    """Small ragged image batches for the collater's pixel half: (images HxWx3 uint8 / float32, row_jit, col_jit)."""
    rng = np.random.RandomState(55)
    cases = []
    for bs, (rj, cj) in ((3, (0, 0)), (4, (5, 11)), (1, (16, 0)), (2, (3, 30))):
        imgs = []
        for i in range(bs):
            h, w = int(rng.randint(20, 90)), int(rng.randint(20, 120))
            im = rng.randint(0, 256, size=(h, w, 3))
            imgs.append(im.astype(np.uint8) if i % 2 else (im / 255.0).astype(np.float32))
        cases.append((imgs, rj, cj))
    return cases


def gen_collater_images():
    """Pixel half of the reference's AspectRatioCollater (Vision.py:730-797) with scale = rand_scale = 1, where its
    cv2.resize to the image's own size is an exact copy: what remains is the jitter placement, transpose and padding."""
    from oracle import ref_shim
    _, vis = ref_shim.load()
    d = {}
    for k, (imgs, rj, cj) in enumerate(image_cases()):
        batch = [(im.copy(), 1.0, 1.0, rj, cj, np.array([]), np.array([])) for im in imgs]
        out, _ = vis.AspectRatioCollater(batch)
        d["case%d" % k] = out.numpy()
    np.savez_compressed(os.path.join(OUT, "collater_images.npz"), **d)


def map_cases():
    """Validation-set predictions / targets for the mAP fixture: jittered copies of the ground truth plus random
    boxes; images without predictions or without targets; exact duplicates (first-argmax tie) and tied scores."""
    rng = np.random.RandomState(91)
    cases = []
    for N, C, holes in ((12, 4, False), (9, 5, True), (30, 3, False)):
        predictions, targets = [], []
        for i in range(N):
            nt = 0 if (holes and i % 4 == 1) else int(rng.randint(1, 7))
            xy = rng.uniform(0, 300, size=(nt, 2))
            wh = rng.uniform(20, 150, size=(nt, 2))
            tb = np.concatenate([xy, xy + wh], axis=1)
            tc = rng.randint(0, C - 1 if holes else C, size=nt)   # holes: the last category has no ground truth
            targets.append([(tb[j].copy(), int(tc[j])) for j in range(nt)])
            pb, pc, ps = [], [], []
            if not (holes and i % 4 == 2):
                for j in range(nt):
                    for _ in range(int(rng.randint(0, 4))):
                        pb.append((tb[j] + rng.normal(0, 6, size=4)).astype(np.float32))
                        pc.append(np.int64(tc[j] if rng.rand() < 0.8 else rng.randint(0, C)))
                        ps.append(np.float32(rng.uniform(0.05, 1.0)))
                for _ in range(int(rng.randint(0, 5))):
                    q = rng.uniform(0, 300, size=2)
                    pb.append(np.concatenate([q, q + rng.uniform(20, 150, size=2)]).astype(np.float32))
                    pc.append(np.int64(rng.randint(0, C)))
                    ps.append(np.float32(rng.uniform(0.05, 0.6)))
                if len(pb) >= 2 and i % 3 == 0:
                    pb.append(pb[0].copy()); pc.append(pc[0]); ps.append(ps[1])   # duplicate box, tied score
            predictions.append([pb, pc, ps])
        cases.append((predictions, targets, {c: "cat%d" % c for c in range(C)}))
    return cases


def gen_map():
    """mAP1 / mAP (Vision.py:1696-1800) of the unmodified reference on map_cases()."""
    import contextlib
    import io
    import warnings
    from oracle import ref_shim
    _, vis = ref_shim.load()
    d = {}
    for k, (predictions, targets, categories) in enumerate(map_cases()):
        for name, thresholds in (("coco", vis.COCO_thresholds), ("pascal", vis.Pascal_thresholds), ("odd", [0.3, 0.62])):
            N, C = len(predictions), len(categories)
            table = np.zeros((len(thresholds), C))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for c in range(C):
                    targs = [[b for b, cc in targets[i] if cc == c] for i in range(N)]
                    preds = [[predictions[i][0][j] for j in range(len(predictions[i][0])) if predictions[i][1][j] == c]
                             for i in range(N)]
                    scores = [[predictions[i][2][j] for j in range(len(predictions[i][0])) if predictions[i][1][j] == c]
                              for i in range(N)]
                    for j, t in enumerate(thresholds):
                        table[j, c] = vis.mAP1(targs, preds, scores, t)
                with contextlib.redirect_stdout(io.StringIO()):
                    mean = vis.mAP(predictions, targets, categories, thresholds)
            assert np.array_equal(np.float64(mean), np.mean(table), equal_nan=True)
            d["case%d_%s_table" % (k, name)] = table
            d["case%d_%s_mean" % (k, name)] = np.float64(mean)
    np.savez_compressed(os.path.join(OUT, "map_scores.npz"), **d)


if __name__ == "__main__":
    if not ref.available():
        sys.exit("reference checkout not found; golden fixtures can only be regenerated where it exists")
    gen_anchors()
    gen_loss_small()
    gen_loss_cfg1()
    gen_postproc_small()
    gen_nms_boxes()
    gen_collater()
    gen_map()
    gen_collater_images()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
