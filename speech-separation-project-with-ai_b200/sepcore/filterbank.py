"""Conv1D "learned filterbank" front end of Raw_with_Convlayer on the GPU.

Reference: segmentation Raw_with_Convlayer.ipynb:83-100 (cell 2); layer
Conv1D(filters=129, kernel_size=2, activation='sigmoid', padding='same') :389
(cell 13); mask multiply :402-403.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._buffers import as_f32_host, current_stream, is_device_tensor, mem_kind, ptr, require_f32_cuda


def segment_raw(wave, seg_len=40):
    """K = ceil(N / L) non-overlapping rows, zero padded (cell 2 :83-96). Host reshape, no compute."""
    wave = np.asarray(wave)
    k = int(np.ceil(len(wave) / seg_len))
    padded = np.concatenate([wave, np.zeros(k * seg_len - len(wave), dtype=wave.dtype)])
    return np.reshape(padded, [k, seg_len])


def conv1d(x, kernel, bias=None, stride=1, padding="same", activation=None):
    """Keras Conv1D forward: x [B, K, C_in], kernel [taps, C_in, N] (Keras layout),
    bias [N] -> [B, K_out, N] float32 with the activation fused."""
    lib = _lib.load()
    dev = is_device_tensor(x)
    if dev:
        import torch

        xx = require_f32_cuda(x, "x")
        ww = require_f32_cuda(kernel, "kernel")
        bb = None if bias is None else require_f32_cuda(bias, "bias")
    else:
        xx, ww = as_f32_host(x), as_f32_host(kernel)
        bb = None if bias is None else as_f32_host(bias)
    if xx.ndim != 3 or ww.ndim != 3 or int(ww.shape[1]) != int(xx.shape[2]):
        raise ValueError("x must be [B, K, C_in] and kernel [taps, C_in, N]")
    batch, rows, c_in = (int(v) for v in xx.shape)
    taps, _, filters = (int(v) for v in ww.shape)
    if padding == "same":
        rows_out = -(-rows // stride)
    elif padding == "valid":
        rows_out = (rows - taps) // stride + 1
    else:
        raise ValueError("padding must be 'same' or 'valid'")
    mem = mem_kind(xx, ww, bb)
    if dev:
        out = torch.empty((batch, rows_out, filters), dtype=torch.float32, device=xx.device)
    else:
        out = np.empty((batch, rows_out, filters), dtype=np.float32)
    _lib.check(lib.sep_conv1d_f32(ptr(xx), ptr(ww), ptr(bb), batch, rows, c_in, taps, filters,
                                  int(stride), _lib.PAD[padding], _lib.ACT[activation], ptr(out), mem,
                                  current_stream(mem, xx if dev else None)), "sep_conv1d_f32")
    return out


def filterbank_separate(wave, enc, dec, masks, stride=8, want_code=False):
    """BASELINE config 5 (a generalisation of the reference's Conv1D front end, which has
    no decoder): frames[k] = wave[k*stride : k*stride+L]; code = relu(frames @ enc);
    est_c = overlap_add((code * mask_c) @ dec, stride).

    wave [B, n], enc [L, N], dec [N, L], masks [B, C, K, N] with K = (n - L)//stride + 1
    -> est [B, C, (K-1)*stride + L] (and code [B, K, N] if want_code).  Runs on the
    tensor cores (tcgen05, 3xTF32, TMEM accumulators); built for L=16, N=256, stride=8.
    """
    lib = _lib.load()
    dev = is_device_tensor(wave)
    if dev:
        import torch

        w, e, d, k = (require_f32_cuda(t, n) for t, n in ((wave, "wave"), (enc, "enc"), (dec, "dec"),
                                                       (masks, "masks")))
    else:
        w, e, d, k = as_f32_host(wave), as_f32_host(enc), as_f32_host(dec), as_f32_host(masks)
    if w.ndim != 2 or e.ndim != 2 or d.ndim != 2 or k.ndim != 4:
        raise ValueError("wave [B, n], enc [L, N], dec [N, L], masks [B, C, K, N] expected")
    batch, n = (int(v) for v in w.shape)
    taps, filters = (int(v) for v in e.shape)
    n_src = int(k.shape[1])
    frames = (n - taps) // stride + 1
    if tuple(int(v) for v in d.shape) != (filters, taps) or \
            tuple(int(v) for v in k.shape) != (batch, n_src, frames, filters):
        raise ValueError("shape mismatch: dec must be [N, L], masks [B, C, K=%d, N]" % frames)
    est_len = (frames - 1) * stride + taps
    mem = mem_kind(w, e, d, k)
    if dev:
        est = torch.empty((batch, n_src, est_len), dtype=torch.float32, device=w.device)
        code = torch.empty((batch, frames, filters), dtype=torch.float32, device=w.device) if want_code else None
    else:
        est = np.empty((batch, n_src, est_len), dtype=np.float32)
        code = np.empty((batch, frames, filters), dtype=np.float32) if want_code else None
    _lib.check(lib.sep_filterbank_separate_f32(ptr(w), ptr(e), ptr(d), ptr(k), batch, n_src, n, taps,
                                               filters, int(stride), ptr(est), ptr(code), mem,
                                               current_stream(mem, w if dev else None)),
               "sep_filterbank_separate_f32")
    return (est, code) if want_code else est
