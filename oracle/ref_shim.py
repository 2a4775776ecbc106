"""Import shim for the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  The reference is imported from the read-only checkout at /root/reference where
that exists (the build container) and otherwise from oracle/_ref/, the byte-for-byte copy staged by
oracle/stage_ref.py (git-ignored; it travels to the GPU box with the working tree like a built .so, and
oracle/_ref/MANIFEST.json holds the SHA-256 of every file).  Used to (a) validate the C oracle
(oracle/retina_oracle.c) against the live reference, (b) generate the golden fixtures under tests/golden/
(tests/golden/make_golden.py), (c) on the GPU box: run the reference ON CUDA as the primary parity oracle
(tests/test_reference_cuda.py, tests/test_reference_callsites.py) and time it (bench.py --impl reference and
the `reference_cuda` extra).  The product package never imports this module.

What it does:
  * stubs the third-party packages General/Core.py:6-22 and Applications/pycocotools import but this
    image lacks (matplotlib, seaborn, spacy, skimage, GPUtil, IPython.display, pycocotools._mask);
  * makes `Tensor.cuda` a no-op, because the path hard-codes `.cuda()` (General/Core.py:70,
    Applications/Vision.py:1499-1501,1592) and this container has no GPU;
  * imports Applications.VisionModels.retinanet and Applications.Vision from /root/reference.
No reference source enters the repository history; the modules are imported where they lie.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")


def _pick_root():
    env = os.environ.get("RETINA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/Applications"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


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


def is_staged_copy():
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(STAGED_ROOT)


_loaded = None


class cpu_mode(object):
    """Context manager: makes `Tensor.cuda` a no-op while active, so that the reference's hard-coded `.cuda()` calls
    (General/Core.py:70, Applications/Vision.py:1499-1501,1592) keep everything on the host even on a machine that
    has a GPU -- the CPU baseline of bench.py on the GPU box.  Restores the real method on exit."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._orig
        return False


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
