"""Utterance-level permutation-invariant MSE (uPIT) on the GPU.

Reference: pit_with_outputsize / pit_loss, uPIT_baseline.ipynb:1023-1059 (cell
28); identical copy in Raw_with_Convlayer.ipynb:338-374 (cell 12).  The
reference is a Keras loss on TensorFlow tensors; TensorFlow is not part of this
stack, so the drop-in takes numpy arrays or torch CUDA tensors with the same
(y_true, y_pred) layout and returns the same scalar (batch SUM of the selected
permutation's cost).
"""
from __future__ import annotations

import itertools

import numpy as np

from . import _lib
from ._buffers import as_f32_host, current_stream, is_device_tensor, mem_kind, ptr, require_f32_cuda


def permutations(n_src):
    """Lexicographic order used by the library; perm[c] = estimate assigned to source c."""
    return list(itertools.permutations(range(n_src)))


def pit_mse(y_true, y_pred, output_size, with_grad=False):
    """Full result of the PIT-MSE: dict(pair [B,C,C], costs [B,P], idx [B], loss, grad?).

    y_true [B, T+1, C*F] (last time row holds the valid length), y_pred [B, T, C*F].
    """
    lib = _lib.load()
    dev = is_device_tensor(y_true)
    if dev:
        import torch

        yt, yp = require_f32_cuda(y_true, "y_true"), require_f32_cuda(y_pred, "y_pred")
    else:
        yt, yp = as_f32_host(y_true), as_f32_host(y_pred)
    if yt.ndim != 3 or yp.ndim != 3:
        raise ValueError("y_true and y_pred must be rank 3")
    batch, rows, width = (int(v) for v in yt.shape)
    frames, feat = rows - 1, int(output_size)
    n_src = width // feat
    if width != n_src * feat or tuple(int(v) for v in yp.shape) != (batch, frames, width):
        raise ValueError("expected y_true [B, T+1, C*F] and y_pred [B, T, C*F] with F=%d" % feat)
    n_perm = len(permutations(n_src))
    mem = mem_kind(yt, yp)
    if dev:
        kw = dict(device=yt.device)
        pair = torch.empty((batch, n_src, n_src), dtype=torch.float64, **kw)
        costs = torch.empty((batch, n_perm), dtype=torch.float64, **kw)
        idx = torch.empty((batch,), dtype=torch.int32, **kw)
        loss = torch.empty((1,), dtype=torch.float64, **kw)
        grad = torch.empty_like(yp) if with_grad else None
    else:
        pair = np.empty((batch, n_src, n_src), dtype=np.float64)
        costs = np.empty((batch, n_perm), dtype=np.float64)
        idx = np.empty((batch,), dtype=np.int32)
        loss = np.empty((1,), dtype=np.float64)
        grad = np.empty_like(yp) if with_grad else None
    _lib.check(lib.sep_pit_mse_f32(ptr(yt), ptr(yp), batch, frames, feat, n_src, ptr(pair), ptr(costs),
                                   ptr(idx), ptr(loss), ptr(grad), mem,
                                   current_stream(mem, yt if dev else None)), "sep_pit_mse_f32")
    out = {"pair": pair, "costs": costs, "idx": idx, "loss": loss[0]}
    if with_grad:
        out["grad"] = grad
    return out


def pit_with_outputsize(output_size):
    """Closure factory with the reference's names (cell 28 :1023, :1059); the
    inner function must stay named `pit_loss` (it is the `custom_objects` key at
    uPIT_baseline.ipynb:1374)."""

    def pit_loss(y_true, y_pred):
        res = pit_mse(y_true, y_pred, output_size)["loss"]
        return res if is_device_tensor(y_true) else np.float32(res)

    return pit_loss
