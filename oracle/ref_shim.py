"""Import shim for the UNMODIFIED reference (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the C oracle
(oracle/retina_oracle.c) against the live reference and (b) generate the golden fixtures under
tests/golden/ (see tests/golden/make_golden.py).  /root/reference does not exist on the GPU box, so
nothing under `-m gpu`, `smoke()` or `bench.py` imports this module.

What it does:
  * stubs the third-party packages General/Core.py:6-22 and Applications/pycocotools import but this
    image lacks (matplotlib, seaborn, spacy, skimage, GPUtil, IPython.display, pycocotools._mask);
  * makes `Tensor.cuda` a no-op, because the path hard-codes `.cuda()` (General/Core.py:70,
    Applications/Vision.py:1499-1501,1592) and this container has no GPU;
  * imports Applications.VisionModels.retinanet and Applications.Vision from /root/reference.
No reference source is copied; the modules are imported where they lie.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RETINA_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """Permissive placeholder module: any non-dunder attribute resolves to a dummy callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _dummy


class _Dummy(object):
    """Callable placeholder whose attributes are placeholders too (e.g. `plt.cm.Blues`)."""

    def __call__(self, *args, **kwargs):
        return None

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self


_dummy = _Dummy()


def _install_stubs():
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.patheffects",
        "matplotlib.collections", "matplotlib.cm", "seaborn", "spacy", "spacy.symbols",
        "skimage", "skimage.io", "skimage.transform", "GPUtil", "IPython", "IPython.display",
        "pycocotools", "pycocotools._mask",
    ]
    for n in names:
        try:
            if n not in sys.modules:
                __import__(n)
        except Exception:
            sys.modules[n] = _Stub(n)
    for n in names:
        if "." in n:
            parent, child = n.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _Stub):
                setattr(sys.modules[parent], child, sys.modules[n])


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Applications"))


_loaded = None


def load():
    """Returns (retinanet_module, vision_module) of the reference, importing them on first use."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    import torch

    _install_stubs()
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    rn = importlib.import_module("Applications.VisionModels.retinanet")
    vis = importlib.import_module("Applications.Vision")
    _loaded = (rn, vis)
    return _loaded
