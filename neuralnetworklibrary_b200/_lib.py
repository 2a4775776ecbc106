"""ctypes binding of libretina_sm100.so (C ABI: include/retina_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `build_library()` with nvcc for sm_100a.
There is NO fallback: if the shared object is missing or a tensor is not on a CUDA device, the
wrappers raise.
"""
import ctypes as C
import glob
import os
import subprocess
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# RETINA_B200_LIB overrides the library path (A/B comparisons of builds); the default is the in-tree build.
LIB_PATH = os.environ.get("RETINA_B200_LIB") or os.path.join(_PKG, "libretina_sm100.so")
SOURCES = ["rn_abi.cu", "rn_assign.cu", "rn_loss.cu", "rn_loss_inst_logits.cu", "rn_step.cu", "rn_post.cu", "rn_loss_levels.cu",
           "rn_metrics.cu"]
# RN_EXTRA_NVCC_FLAGS=-DRN_EXPERIMENTAL (with a forced rebuild) adds the two alternative implementations of rn_loss_step that
# measured slower than the default (the persistent kernel and the byte-map chain): 22 more kernels
if "-DRN_EXPERIMENTAL" in os.environ.get("RN_EXTRA_NVCC_FLAGS", ""):
    SOURCES += ["rn_loss_inst_bytes.cu", "rn_loss_inst_bytes_logits.cu"]
BUILD_DIR = os.path.join(_PKG, "csrc", "_build")   # object files (git-ignored); the .so is what travels

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # never contract a*b+c behind our back; fused ops are spelled fmaf() explicitly
    # no -split-compile: with it, identical command lines produced different SASS from run to run (the partition of a
    # translation unit depends on thread timing) and the level-tensor kernels came out 2-3 % apart (profiles/r02_summary.md)
    "-Xcompiler", "-fPIC",
]

RN_OK, RN_ERR_INVALID_ARG, RN_ERR_WORKSPACE, RN_ERR_CUDA = 0, 1, 2, 3
MATCH_NEG, MATCH_IGNORE = -1, -2
MAX_TOP_K = 4096
MAX_K = 16
NUM_LEVELS = 5


class RetinaB200Error(RuntimeError):
    pass


def _headers():
    return glob.glob(os.path.join(_PKG, "csrc", "*.cuh")) + [os.path.join(_ROOT, "include", "retina_b200.h")]


def _newer(deps, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _stale():
    return _newer([os.path.join(_PKG, "csrc", s) for s in SOURCES] + _headers(), LIB_PATH)


def build_library(force=False, verbose=False):
    """Compiles every CUDA source of the package for sm_100a (one nvcc process per translation unit, in
    parallel; cross-compiles without a GPU) and links LIB_PATH."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    jobs = []
    for src in SOURCES:
        path = os.path.join(_PKG, "csrc", src)
        obj = os.path.join(BUILD_DIR, src[:-3] + ".o")
        if force or _newer([path] + _headers(), obj):
            extra = os.environ.get("RN_EXTRA_NVCC_FLAGS", "").split()   # build-time only (e.g. -DRN_STEP_TIMING for profiles/step_timing.py)
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            for _, other in jobs:
                if other.poll() is None:
                    other.kill()
            raise RetinaB200Error("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            print(out)
    objs = [os.path.join(BUILD_DIR, s[:-3] + ".o") for s in SOURCES]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RetinaB200Error("link failed:\n" + res.stdout)
    return LIB_PATH


_vp, _f32p, _f64p, _i32p, _i64p = C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p
_hf32p = C.POINTER(C.c_float)
_pp = C.POINTER(C.c_void_p)   # host array of device pointers

# name -> (restype, argtypes).  Mirrors include/retina_b200.h one to one (tests check the symbols).
PROTOTYPES = {
    "rn_last_error": (C.c_char_p, []),
    "rn_abi_version": (C.c_int, []),
    "rn_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "rn_get_option": (C.c_int, [C.c_char_p]),
    "rn_num_anchors": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "rn_anchors": (C.c_int, [C.c_int, C.c_int, _f64p, C.c_int, _f32p, _vp]),
    "rn_assign": (C.c_int, [_f32p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f32p, C.c_int,
                            C.c_float, C.c_float, _i32p, _i32p, _f32p, _vp]),
    "rn_max_overlaps": (C.c_int, [_f32p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f32p, C.c_int,
                                  _f32p, _vp]),
    "rn_stage_targets": (C.c_int, [_vp, _i64p, _i32p, _vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i64p, _vp]),
    "rn_stage_images": (C.c_int, [_f32p, _i64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _vp]),
    "rn_stage_images_u8": (C.c_int, [_vp, _i64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _hf32p, _hf32p,
                                     _f32p, _vp]),
    "rn_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rn_loss": (C.c_int, [_f32p, _f32p, _f32p, _i64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_int, _f64p, C.c_int, _f32p, C.c_double, C.c_double, C.c_double, C.c_int,
                          _f32p, _f32p, _f32p, _vp, C.c_size_t, _vp]),
    "rn_loss_logits": (C.c_int, [_f32p, _f32p, _f32p, _i64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _f64p, C.c_int, _f32p, C.c_double, C.c_double, C.c_double, C.c_int,
                                 _f32p, _f32p, _f32p, _f32p, _vp, C.c_size_t, _vp]),
    "rn_loss_step_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rn_loss_step_state_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rn_loss_step_state_init": (C.c_int, [_vp, C.c_size_t, _vp]),
    "rn_loss_step": (C.c_int, [_f32p, _f32p, _f32p, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int,
                               _f32p, C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                               _f32p, _f32p, _f32p, _f32p, _i32p, _i32p, _vp, C.c_size_t, _vp, C.c_size_t, _vp]),
    "rn_peer_exchange_bytes": (C.c_size_t, [C.c_int]),
    "rn_peer_exchange": (C.c_int, [_f32p, _pp, C.c_int, C.c_int, _vp, _vp]),
    "rn_peer_exchange_to": (C.c_int, [_f32p, _f32p, _pp, C.c_int, C.c_int, _vp, _vp]),
    "rn_scale_grads": (C.c_int, [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _vp]),
    "rn_postproc_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rn_postproc": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _f32p,
                              _hf32p, _hf32p, C.c_float, C.c_float, C.c_int, C.c_int, _f32p, _i64p, _f32p,
                              _i32p, _i32p, _i32p, _vp, C.c_size_t, _vp]),
    "rn_loss_levels_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rn_loss_levels": (C.c_int, [_pp, _pp, C.c_int, _f32p, _i64p, _i32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 _f64p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _pp, _pp, _pp, _f32p, _vp,
                                 C.c_size_t, _vp]),
    "rn_postproc_levels": (C.c_int, [_pp, _pp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _hf32p, _hf32p,
                                     C.c_float, C.c_float, C.c_int, C.c_int, _f32p, _i64p, _f32p, _i32p, _i32p, _i32p, _vp,
                                     C.c_size_t, _vp]),
    "rn_map_match": (C.c_int, [_f32p, _i32p, _i32p, _f32p, _i32p, _i32p, C.c_int, C.c_int, _f32p, C.c_int, _vp, _vp]),
    "rn_nms_batch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rn_nms_batch": (C.c_int, [_f32p, _i64p, _f32p, _i32p, C.c_int, C.c_int, _i32p, _vp, C.c_int, C.c_float, C.c_int, C.c_int,
                               _f32p, _i64p, _f32p, _i32p, _i32p, _vp, C.c_size_t, _vp]),
    "rn_map_ap_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rn_map_ap": (C.c_int, [_f32p, _i32p, _i32p, _vp, _i32p, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_size_t, _vp]),
    "rn_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rn_nms": (C.c_int, [_f32p, _i64p, _f32p, C.c_int, C.c_float, C.c_int, C.c_int, _i32p, _i32p, _vp,
                         C.c_size_t, _vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Loads the shared library (never builds implicitly on a GPU box: a missing .so is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RetinaB200Error(
                    "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc != RN_OK:
        msg = load().rn_last_error().decode("utf-8", "replace")
        if rc == RN_ERR_INVALID_ARG:
            raise ValueError(msg)  # the reference reports bad arguments as ValueError (Learner.py:339-340)
        raise RetinaB200Error("libretina_sm100 error %d: %s" % (rc, msg))


def has_experimental():
    """True when the loaded library was built with -DRN_EXPERIMENTAL (the opt-in implementations of rn_loss_step)."""
    return load().rn_get_option(b"experimental") == 1


class option(object):
    """Context manager that sets a library tuning / test switch (rn_set_option) and restores it on exit, e.g.
    `with _lib.option("assign_dense", 1): ...` forces the dense assignment kernel."""

    def __init__(self, name, value):
        self.name, self.value = name.encode(), int(value)

    def __enter__(self):
        lib = load()
        self.old = lib.rn_get_option(self.name)
        check(lib.rn_set_option(self.name, self.value))
        return self

    def __exit__(self, *exc):
        check(load().rn_set_option(self.name, max(self.old, 0)))
        return False


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    """Host array of the tensors' device pointers (None -> NULL array)."""
    if tensors is None:
        return None
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # the cudaStream_t without building a Stream object


def stream_ptr(device=None):
    """The current CUDA stream of `device` (default: the current device) as a void*.  Called once or twice per library
    call: torch.cuda.current_stream() costs ~8 us of Python per call, the raw lookup well under one."""
    if _raw_stream is not None:
        if device is None:
            idx = torch.cuda.current_device()
        elif isinstance(device, int):
            idx = device
        else:
            idx = device.index if getattr(device, "index", None) is not None else torch.cuda.current_device()
        return C.c_void_p(_raw_stream(idx))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RetinaB200Error("%s must live on a CUDA device: this path has no CPU fallback" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def base_ptr(base_np):
    """Host pointer to the float64 [5][K][4] base table (a contiguous numpy array)."""
    return base_np.ctypes.data_as(_f64p)


class Workspace(object):
    """Grow-only per-(device, stream) scratch buffer handed to the library (256-byte aligned by the
    caching allocator).  zeroed=True: the buffer is zero-filled when it is (re)allocated -- the contract of
    rn_loss_step, whose kernel leaves its workspace zeroed again after every call."""

    def __init__(self, zeroed=False):
        self._bufs = {}
        self._zeroed = zeroed

    def get(self, nbytes, device):
        key = (device.index, stream_ptr(device).value)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            alloc = torch.zeros if self._zeroed else torch.empty
            buf = alloc(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf
