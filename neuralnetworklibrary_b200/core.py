"""Boundary helpers with the semantics of the reference's General/Core.py:46-102 (TEN, ARR, list_del,
list_mult): they fix the dtypes and device the hot path sees, so the drop-in wrappers coerce inputs
exactly like the reference does."""
import numpy as np
import torch


def TEN(x, GPU=True):
    """list / ndarray / python scalar -> FloatTensor (float32) or LongTensor (int64), moved to the
    current CUDA device when GPU is True (reference Core.py:46-71)."""
    if isinstance(x, list):
        x = np.array(x)
    if isinstance(x, np.ndarray):
        if x.dtype in (np.float32, np.float64):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        elif x.dtype in (np.int32, np.int64):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64))
    elif isinstance(x, (float, np.float32, np.float64)):
        x = torch.tensor(float(x), dtype=torch.float32)
    elif isinstance(x, (int, np.int32, np.int64)):
        x = torch.tensor(int(x), dtype=torch.int64)
    if GPU:
        x = x.cuda()
    return x


def ARR(x):
    """torch tensor (any device) -> numpy array on the host (reference Core.py:73-76)."""
    return x.detach().cpu().numpy() if x.is_cuda else x.detach().numpy()


def list_del(L, idxs):
    """L without the items at the positions in idxs (reference Core.py:88-96)."""
    drop = set(int(i) for i in idxs)
    return [v for i, v in enumerate(L) if i not in drop]


def list_mult(L, c):
    """Element-wise product of a list (or a single number) with c (reference Core.py:98-102)."""
    return [v * c for v in L] if type(L) == list else L * c
