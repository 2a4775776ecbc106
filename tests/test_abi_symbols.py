"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/retina_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from neuralnetworklibrary_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "retina_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = _lib.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.PROTOTYPES) == names, "binding and header disagree"


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.rn_abi_version() == 1
    assert lib.rn_num_anchors(512, 512, 9) == 49104
    assert lib.rn_num_anchors(800, 1333, 9) == 200700
    assert lib.rn_num_anchors(800, 1344, 9) == 201600
    assert lib.rn_num_anchors(0, 10, 9) == 0
    assert lib.rn_loss_workspace_bytes(16, 201600, 80) % 256 == 0
    assert lib.rn_postproc_workspace_bytes(64, 201600, 1000) >= 64 * 201600 * 8


def test_sass_is_sm100a_without_fma_contraction_in_iou():
    """The shipped cubin targets sm_100a only (no PTX JIT fallback for other archs is relied upon)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
