"""Harness that drives the UNMODIFIED reference at the call sites of the hot path:

    General/Learner.py:286-381   Learner.predict('val')  -> model forward (Vision.py:1469 self.AnchorGenerator(x)),
                                 self.model.BBoxPredictor(x_batch, reg, clas, anchors, ...11 positional...), list_mult
    General/Learner.py:490-516   Learner.train1minibatch -> self.loss_func(y_pred, y_batch); loss.backward(); step
    Applications/Vision.py:2036-2121  ImageLearner.TTA_bbox -> BBoxPredictor per pass, un-transform, TEN(...) and
                                 vmods.retinanet.nms(...) on the merged predictions (Vision.py:2104-2119)

TEST INFRASTRUCTURE.  Nothing of the code under test is re-stated here: the harness only supplies what the reference
needs around those lines and this image lacks -- a data object (the Pascal notebook and its images are not in the
checkout), progress bars (tqdm_notebook needs ipywidgets), plt.imread (matplotlib is absent), and random-initialised
weights in place of the LFS pointer `RetinanetPretrainedCOCO.pt` that ObjectDetectionNet.__init__ torch.load()s
(Vision.py:1412 -> retinanet.py:432-434).  Each `swap_*` function below is exactly the substitution INTEGRATION.md
section 2 tells a maintainer to make.
"""
import copy

import numpy as np
import torch

from oracle import ref_shim


class FakeImages(object):
    """Deterministic stand-in for the validation images: image j of pass i is N(0,1) noise of a fixed size."""

    def __init__(self, n, H, W, seed=0):
        self.n, self.H, self.W, self.seed = n, H, W, seed
        self.images = [{"img": "img%d.jpg" % j, "scale": 0.5 + 0.25 * j} for j in range(n)]
        self.IMG_PATH = "/nonexistent/"

    def batch(self, j, pass_id=0):
        g = torch.Generator().manual_seed(self.seed + 1000 * pass_id + j)
        return torch.randn(1, 3, self.H, self.W, generator=g)


class FakeLoader(object):
    """Iterable of (x_batch, y_batch) with identity comparison (Learner.predict tests `dl == self.data.val_dl`)."""

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class FakeData(object):
    target_type = "bbox"

    def __init__(self, imgs, bs=2):
        self.bs = bs
        self.val_ds = imgs
        self.test_ds = None
        self.val_dl = FakeLoader([(imgs.batch(j), [torch.zeros(1, 1, 4), torch.zeros(1, 1, dtype=torch.int64)]) for j in range(imgs.n)])
        self.test_dl = None


class FakeTransform(object):
    """What TTA_bbox reads from a transform (Vision.py:2070-2071, :2086-2087)."""

    def __init__(self, n, augment, seed):
        self.n, self.augment, self.seed = n, augment, seed
        self.calls = 0

    def get_values(self):
        rng = np.random.RandomState(self.seed + self.calls)
        self.calls += 1
        if self.augment:
            self.row_jitter_values = list(rng.randint(0, 8, self.n))
            self.col_jitter_values = list(rng.randint(0, 8, self.n))
            self.scale_values = list(rng.uniform(0.9, 1.1, self.n))
            self.flip_values = list(rng.randint(0, 2, self.n))
        else:
            self.row_jitter_values = [0] * self.n
            self.col_jitter_values = [0] * self.n
            self.scale_values = [1.0] * self.n
            self.flip_values = [0] * self.n


def patch_environment(monkeypatch):
    """Progress bars, plt.imread and the weight file.  Returns (rn, vis)."""
    rn, vis = ref_shim.load()
    import General.Learner as learner_mod

    def passthrough(it, *a, **k):
        return it

    for mod in (learner_mod, vis):
        for name in ("PBar", "PBarPredict", "PBarTrain", "PBarEvalTrain", "PBarEvalVal", "PBarTTA"):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, passthrough)
    # random-initialised backbone instead of torch.load of the LFS pointer (retinanet.py:430-435); a shallow ResNet keeps
    # the test fast -- the backbone is out of scope, only its [anchors, reg, clas] interface matters
    monkeypatch.setattr(rn, "retinanet", lambda: rn.RetinaNet(80, rn.BasicBlock, [1, 1, 1, 1]))
    return rn, vis


def build_model(vis, num_classes, seed=0):
    """The reference's ObjectDetectionNet (Vision.py:1382-1471), unmodified, with its head's output layers re-drawn so
    that scores are not the constant prior 0.01 the reference initialises them to (Vision.py:1434-1437)."""
    torch.manual_seed(seed)
    model = vis.ObjectDetectionNet(num_classes=num_classes, feature_size=32)
    with torch.no_grad():
        model.classifier.output.weight.normal_(0.0, 0.08)
        model.classifier.output.bias.fill_(-3.0)
        model.regressor.output.weight.normal_(0.0, 0.02)
    return model


def swap_predictors(model, ours_retinanet):
    """the two attributes of the model."""
    model.AnchorGenerator = ours_retinanet.AnchorGenerator()
    model.BBoxPredictor = ours_retinanet.BBoxPredictor()
    return model


def swap_nms(monkeypatch, rn, ours_retinanet):
    """`vmods.retinanet.nms = nms` (the TTA call site, Vision.py:2118)."""
    monkeypatch.setattr(rn, "nms", ours_retinanet.nms)


def make_learner(vis, tmp_path, model, data, loss_func):
    opt = vis.Optimizer(torch.optim.SGD, model)
    return vis.ImageLearner(str(tmp_path), data, model, optimizer=opt, loss_func=loss_func)


def patch_tta_inputs(monkeypatch, vis, imgs):
    """Dataset / DataLoader / plt.imread for TTA_bbox (Vision.py:2064-2084): pass i yields imgs.batch(j, i)."""
    state = {"pass": 0}

    class _DS(object):
        def __init__(self, IMG_PATH, images, transform, target_type, ds_type):
            self.pass_id = state["pass"]
            state["pass"] += 1

        def __len__(self):
            return imgs.n

    def _DL(ds, **kw):
        return FakeLoader([(imgs.batch(j, ds.pass_id), None) for j in range(imgs.n)])

    monkeypatch.setattr(vis, "ImageDataset", _DS)
    monkeypatch.setattr(vis, "DataLoader", _DL)
    monkeypatch.setattr(vis.plt, "imread", lambda path: np.zeros((2 * imgs.H, 2 * imgs.W, 3), np.float32), raising=False)
    return state


def clone_model(model):
    return copy.deepcopy(model)
