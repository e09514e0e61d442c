"""Sample-format steps either side of the signal path, on the GPU.

Reference: `audiowrite` -- uPIT_baseline.ipynb:1317-1354 (cell 40); wav decoding as
`wavread` / `librosa.load` deliver it for 16-bit PCM -- metrics/evaluate_metrics.py:7-12,
parallel_stft.py:213.
"""
from __future__ import annotations

import threading

import numpy as np

from . import _lib
from ._buffers import as_f32_host, current_stream, is_device_tensor, ptr


def pcm16_to_float32(pcm, out=None):
    """int16 PCM -> float32 in [-1, 1): pcm / 32768 (what sf.read(dtype='float32') returns).
    `out` (device mode): a preallocated contiguous float32 CUDA tensor of the same shape."""
    lib = _lib.load()
    if is_device_tensor(pcm):
        import torch

        if pcm.dtype != torch.int16 or not pcm.is_contiguous():
            raise ValueError("pcm must be a contiguous int16 CUDA tensor")
        if out is None:
            out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
        elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != pcm.numel():
            raise ValueError("out must be a contiguous float32 CUDA tensor with pcm's element count")
        mem, n, stream = _lib.MEM_DEVICE, pcm.numel(), current_stream(_lib.MEM_DEVICE, pcm)
    else:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.empty(pcm.shape, dtype=np.float32)
        mem, n, stream = _lib.MEM_HOST, pcm.size, None
    _lib.check(lib.sep_pcm16_to_f32(ptr(pcm), int(n), ptr(out), mem, stream), "sep_pcm16_to_f32")
    return out


def audiowrite_int16(data, normalize=False, out=None, clipped=None):
    """The sample conversion inside `audiowrite` (cell 40 :1331-1346) for data [n] or [B, n]
    (row by row): returns (int16 array, clipped count per row -- a Python int for 1-D input).
    The arithmetic runs in the input's own precision like numpy's: float32 stays float32,
    float64 (what the reference's istft returns) stays float64; other float types are computed
    in float32.  `out` / `clipped` (device mode): preallocated int16 [B, n] / int64 [B] CUDA tensors."""
    lib = _lib.load()
    dev = is_device_tensor(data)
    if dev:
        import torch

        if data.dtype == torch.float64:
            x, fn = data.contiguous(), lib.sep_audiowrite_i16_f64
        else:
            x = data if data.dtype == torch.float32 and data.is_contiguous() else data.float().contiguous()
            fn = lib.sep_audiowrite_i16_f32
        rows = x.reshape(1, -1) if x.dim() == 1 else x
        if out is None:
            out = torch.empty(rows.shape, dtype=torch.int16, device=x.device)
        if clipped is None:
            clipped = torch.empty((rows.shape[0],), dtype=torch.int64, device=x.device)
        mem, stream = _lib.MEM_DEVICE, current_stream(_lib.MEM_DEVICE, x)
    else:
        x = np.asarray(data)
        if x.dtype == np.float64:
            x, fn = np.ascontiguousarray(x), lib.sep_audiowrite_i16_f64
        else:
            x, fn = as_f32_host(x), lib.sep_audiowrite_i16_f32
        rows = x.reshape(1, -1) if x.ndim == 1 else x
        out = np.empty(rows.shape, dtype=np.int16)
        clipped = np.empty((rows.shape[0],), dtype=np.int64)
        mem, stream = _lib.MEM_HOST, None
    if rows.ndim != 2 or rows.shape[1] < 1:
        raise ValueError("data must be [n] or [B, n] with n >= 1")
    _lib.check(fn(ptr(rows), int(rows.shape[0]), int(rows.shape[1]), int(bool(normalize)),
                  ptr(out), ptr(clipped), mem, stream), "sep_audiowrite_i16")
    if (x.dim() if dev else x.ndim) == 1:
        return out.reshape(-1), int(clipped[0])
    return out, clipped


def audiowrite(data, path, samplerate=16000, normalize=False, threaded=True):
    """Reference signature (uPIT_baseline.ipynb:1317): converts on the GPU, writes the wav with
    scipy (optionally on a thread, like the reference); returns the number of clipped samples.

    Like the reference cell: the peak of `normalize` is the maximum over the WHOLE array (a
    multi-channel [n, ch] array is normalised by one common peak and written with its shape);
    float data is scaled by 32767, integer data is not (:1339-1340) -- with normalize it is first
    promoted to float64 (:1335-1337), without it only clipped to the int16 range and cast."""
    from scipy.io.wavfile import write as wav_write

    arr = np.asarray(data)
    if arr.dtype.kind != 'f' and not normalize:
        # integer samples pass through: count, clip, cast -- a format step with no arithmetic to accelerate
        clipped = int(np.sum(arr > np.iinfo(np.int16).max))
        pcm = np.clip(arr, np.iinfo(np.int16).min, np.iinfo(np.int16).max).astype(np.int16)
    else:
        if arr.dtype.kind != 'f':
            arr = arr.astype(np.float64)
        pcm, clipped = audiowrite_int16(arr.reshape(-1), normalize)
        pcm = pcm.reshape(arr.shape)
    if clipped > 0:
        print('Warning, clipping {} samples'.format(clipped))
    if threaded:
        threading.Thread(target=wav_write, args=(path, samplerate, pcm)).start()
    else:
        wav_write(path, samplerate, pcm)
    return clipped
