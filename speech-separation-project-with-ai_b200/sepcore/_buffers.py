"""Buffer plumbing: numpy arrays travel as host pointers (the library stages
them), torch CUDA tensors as device pointers (zero copies).  torch is imported
only when a tensor shows up -- it is plumbing for device memory and streams."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import MEM_DEVICE, MEM_HOST


def is_device_tensor(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


def as_f32_host(x):
    """C-contiguous float32 numpy view/copy of an array-like."""
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


def ptr(x):
    """void* of a numpy array or torch tensor (None -> NULL)."""
    if x is None:
        return None
    if is_device_tensor(x) or hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def mem_kind(*arrays):
    """All buffers of one call must live on the same side."""
    kinds = {is_device_tensor(a) for a in arrays if a is not None}
    if len(kinds) > 1:
        raise ValueError("mixing host arrays and CUDA tensors in one call is not supported")
    return MEM_DEVICE if kinds == {True} else MEM_HOST


def current_stream(mem, like=None):
    """cudaStream_t (as void*) to launch on: torch's current stream in device
    mode so that the caller's tensor lifetimes and ordering hold; the legacy
    default stream in host mode (the call synchronises before returning)."""
    if mem == MEM_DEVICE:
        import torch

        return C.c_void_p(torch.cuda.current_stream(like.device).cuda_stream)
    return None


def require_f32_cuda(x, name):
    import torch

    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)
    return x
