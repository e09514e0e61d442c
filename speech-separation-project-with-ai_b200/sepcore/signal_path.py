"""Host side of the STFT front end / iSTFT back end, with the reference's names
and signatures; every compute call goes through libsepcore (CUDA, sm_100a).

Reference:
  segment_axis, _samples_to_stft_frames, _stft_frames_to_samples, stft
      parallel_stft.py:37-196 (duplicated in parallel_stft_single.py:39-198 and
      uPIT_baseline.ipynb cells 5-8)
  _biorthogonal_window_loopy, istft    uPIT_baseline.ipynb:1234-1307 (cells 38-39)
  feature / label math                 parallel_stft.py:262-272

numpy inputs are staged by the library (drop-in mode) and results come back in
the reference's dtypes (complex128 / float64, computed in float32 on the GPU);
torch CUDA tensors are used in place and results stay on the device as
float32/complex64 (throughput mode).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._buffers import (as_f32_host, current_stream, is_device_tensor, mem_kind, ptr,
                       require_f32_cuda)
from .plan import get_plan


# ------------------------------------------------------------------ a1
def _samples_to_stft_frames(samples, size, shift):
    """Number of STFT frames for `samples` time samples (parallel_stft.py:125-134).
    Integer ceil of (samples - size + shift) / shift (the reference's `np.int`
    no longer exists in numpy)."""
    return int(np.ceil((float(samples) - size + shift) / shift))


def _stft_frames_to_samples(frames, size, shift):
    """Time samples spanned by `frames` frames (parallel_stft.py:136-144)."""
    return frames * shift + size - shift


# ------------------------------------------------------------------ a2
def segment_axis(a, length, overlap=0, axis=None, end='cut', endvalue=0):
    """Chop `a` along `axis` into overlapping frames (parallel_stft.py:37-123).

    numpy input: host index logic only -- like the reference it returns a
    strided VIEW when it can (no compute, no copy).  torch CUDA float32 input:
    the framing copy runs on the GPU (sep_segment_axis_f32).
    """
    if is_device_tensor(a):
        return _segment_axis_device(a, length, overlap, axis, end, endvalue)
    a = np.asarray(a)
    if axis is None:
        a = np.ravel(a)
        axis = 0
    if axis < 0:
        axis += a.ndim
    if overlap >= length:
        raise ValueError("frames cannot overlap by more than 100%")
    if overlap < 0 or length <= 0:
        raise ValueError("overlap must be nonnegative and length must be positive")
    a = _fit_tail(np, a, length, overlap, axis, end, endvalue)
    n = a.shape[axis]
    if n == 0:
        raise ValueError(
            "Not enough data points to segment array in 'cut' mode; try 'pad' or 'wrap'")
    hop = length - overlap
    assert n >= length and (n - length) % hop == 0
    count = 1 + (n - length) // hop
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    step = a.strides[axis]
    shape = a.shape[:axis] + (count, length) + a.shape[axis + 1:]
    strides = a.strides[:axis] + (hop * step, step) + a.strides[axis + 1:]
    return np.lib.stride_tricks.as_strided(a, shape=shape, strides=strides, writeable=False)


def _fit_tail(xp, a, length, overlap, axis, end, endvalue):
    """Ragged-tail handling of parallel_stft.py:72-99 for numpy or torch."""
    hop = length - overlap
    n = a.shape[axis]
    if not (n < length or (n - length) % hop):
        return a
    if n > length:
        rounddown = length + ((n - length) // hop) * hop
        roundup = rounddown + hop
    else:
        roundup, rounddown = length, 0
    moved = xp.swapaxes(a, -1, axis)
    if end == 'cut':
        moved = moved[..., :rounddown]
    elif end in ('pad', 'wrap'):
        if xp is np:
            grown = np.empty(moved.shape[:-1] + (roundup,), dtype=a.dtype)
        else:
            grown = a.new_empty(tuple(moved.shape[:-1]) + (roundup,))
        grown[..., :n] = moved
        if end == 'pad':
            grown[..., n:] = endvalue
        else:
            grown[..., n:] = moved[..., :roundup - n]
        moved = grown
    return xp.swapaxes(moved, -1, axis)


def _segment_axis_device(a, length, overlap, axis, end, endvalue):
    import torch

    if axis is None:
        a = a.reshape(-1)
        axis = 0
    if axis < 0:
        axis += a.ndim
    if overlap >= length:
        raise ValueError("frames cannot overlap by more than 100%")
    if overlap < 0 or length <= 0:
        raise ValueError("overlap must be nonnegative and length must be positive")
    a = _fit_tail(torch, a, length, overlap, axis, end, endvalue)
    n = a.shape[axis]
    if n == 0:
        raise ValueError(
            "Not enough data points to segment array in 'cut' mode; try 'pad' or 'wrap'")
    hop = length - overlap
    count = 1 + (n - length) // hop
    rows = torch.movedim(a, axis, -1).contiguous().to(torch.float32)
    lead = rows.shape[:-1]
    batch = int(np.prod(lead)) if lead else 1
    out = torch.empty((batch, count, length), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    _lib.check(lib.sep_segment_axis_f32(ptr(rows), batch, n, length, overlap, ptr(out),
                                        _lib.MEM_DEVICE, current_stream(_lib.MEM_DEVICE, a)),
               "sep_segment_axis_f32")
    out = out.reshape(tuple(lead) + (count, length))
    # frames axis at `axis`, in-frame axis right after it
    nd = out.ndim
    order = list(range(nd - 2))
    order[axis:axis] = [nd - 2, nd - 1]
    return out.permute(order)


# ------------------------------------------------------------------ a3
def stft(time_signal, time_dim=None, size=1024, shift=256, window=None, fading=True,
         window_length=None):
    """Short-time Fourier transform of a multi-channel signal (parallel_stft.py:146-196).

    Same signature and semantics as the reference; `window` is the reference's
    window callable (default: symmetric Blackman, `scipy.signal.blackman` in the
    reference) or an array of `size` taps.  Returns frames at `time_dim` and
    size/2+1 bins at `time_dim + 1`.
    """
    plan = get_plan(size, shift, window, fading, window_length)
    lib = _lib.load()
    dev = is_device_tensor(time_signal)
    if dev:
        import torch

        x = time_signal
        if time_dim is None:
            time_dim = int(np.argmax(tuple(x.shape)))
        rows = torch.movedim(x, time_dim, -1).to(torch.float32).contiguous()
    else:
        x = np.asarray(time_signal)
        if time_dim is None:
            time_dim = int(np.argmax(x.shape))
        rows = as_f32_host(np.moveaxis(x, time_dim, -1))
    if time_dim < 0:
        time_dim += x.ndim
    lead = tuple(rows.shape[:-1])
    n = int(rows.shape[-1])
    batch = int(np.prod(lead)) if lead else 1
    frames = plan.frames(n)
    if dev:
        out = torch.empty((batch, frames, plan.bins), dtype=torch.complex64, device=x.device)
        mem = _lib.MEM_DEVICE
    else:
        out = np.empty((batch, frames, plan.bins), dtype=np.complex64)
        mem = _lib.MEM_HOST
    if frames > 0 and batch > 0:
        _lib.check(lib.sep_stft_f32(plan.handle, ptr(rows), batch, n, n, ptr(out), mem,
                                    current_stream(mem, x if dev else None)), "sep_stft_f32")
    out = out.reshape(lead + (frames, plan.bins))
    nd = out.ndim
    order = list(range(nd - 2))
    order[time_dim:time_dim] = [nd - 2, nd - 1]
    if dev:
        return out.permute(order)
    return np.transpose(out, order).astype(np.complex128)


# ------------------------------------------------------------------ a4
def stft_features(mix, sources=None, size=256, shift=128, window=None):
    """|X| || angle X network inputs and PSA labels (parallel_stft.py:262-272).

    mix [B, N] (or [N]); sources [B, C, N] (or [C, N]) or None.
    Returns (inputs [B, T, 2F], labels [B, T, C*F] or None) as float32 -- what the
    reference writes into its TFRecords (FloatList is float32).
    """
    plan = get_plan(size, shift, window, True)
    lib = _lib.load()
    dev = is_device_tensor(mix)
    squeeze = mix.ndim == 1
    if dev:
        import torch

        m = require_f32_cuda(mix.reshape(1, -1) if squeeze else mix, "mix")
        s = None if sources is None else require_f32_cuda(
            sources.reshape((1,) + tuple(sources.shape)) if squeeze else sources, "sources")
    else:
        m = as_f32_host(mix).reshape(1, -1) if squeeze else as_f32_host(mix)
        s = None if sources is None else as_f32_host(sources)
        if s is not None and squeeze:
            s = s.reshape((1,) + s.shape)
    batch, n = int(m.shape[0]), int(m.shape[1])
    n_src = 0 if s is None else int(s.shape[1])
    if s is not None and (tuple(s.shape[:1]) != (batch,) or int(s.shape[2]) != n):
        raise ValueError("sources must be [B, C, N] matching mix [B, N]")
    frames = plan.frames(n)
    mem = mem_kind(m, s)
    if dev:
        feats = torch.empty((batch, frames, 2 * plan.bins), dtype=torch.float32, device=m.device)
        labels = None if s is None else torch.empty((batch, frames, n_src * plan.bins),
                                                    dtype=torch.float32, device=m.device)
    else:
        feats = np.empty((batch, frames, 2 * plan.bins), dtype=np.float32)
        labels = None if s is None else np.empty((batch, frames, n_src * plan.bins), dtype=np.float32)
    _lib.check(lib.sep_stft_features_f32(plan.handle, ptr(m), ptr(s), batch, n_src, n, ptr(feats),
                                         ptr(labels), mem, current_stream(mem, m if dev else None)),
               "sep_stft_features_f32")
    if squeeze:
        feats = feats[0]
        labels = None if labels is None else labels[0]
    return feats, labels


# ------------------------------------------------------------------ a7
def _biorthogonal_window_loopy(analysis_window, shift):
    """Biorthogonal synthesis window (uPIT_baseline.ipynb:1234-1259, cell 38).
    Computed once per plan inside libsepcore (float64), including the
    reference's exclusion of the last tap from the sums (:1253)."""
    taps = np.ascontiguousarray(np.asarray(analysis_window, dtype=np.float64))
    assert np.mod(len(taps), shift) == 0
    return get_plan(len(taps), shift, taps, True).synthesis_window()


# ------------------------------------------------------------------ a8
def istft(stft_signal, size=1024, shift=256, window=None, fading=True, window_length=None):
    """Inverse STFT with overlap-add (uPIT_baseline.ipynb:1269-1307, cell 39).

    stft_signal [T, size/2+1] as in the reference (one utterance), or
    [B, T, size/2+1] for a batch.  Returns float64 [T*shift - (size-shift)]
    (numpy in) or float32 CUDA tensor (tensor in).
    """
    assert stft_signal.shape[-1] == size // 2 + 1
    assert size % shift == 0        # cell 38 :1245
    if size % 2:
        # the reference's `irfft(stft_signal[j])` (:1301, no `n`) returns 2 (F - 1) = size - 1 samples for an odd size,
        # and adding them to a `size`-long window product raises exactly this numpy error
        raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,) " % (size, size - 1))
    plan = get_plan(size, shift, window, fading, window_length)
    lib = _lib.load()
    dev = is_device_tensor(stft_signal)
    single = stft_signal.ndim == 2
    if dev:
        import torch

        spec = stft_signal.to(torch.complex64).contiguous()
        if single:
            spec = spec.reshape((1,) + tuple(spec.shape))
        raw = torch.view_as_real(spec)
    else:
        spec = np.ascontiguousarray(np.asarray(stft_signal), dtype=np.complex64)
        if single:
            spec = spec.reshape((1,) + spec.shape)
        raw = spec.view(np.float32)
    batch, frames = int(spec.shape[0]), int(spec.shape[1])
    length = plan.istft_samples(frames)
    if dev:
        out = torch.zeros((batch, length), dtype=torch.float32, device=spec.device)
        mem = _lib.MEM_DEVICE
    else:
        out = np.zeros((batch, length), dtype=np.float32)
        mem = _lib.MEM_HOST
    if frames > 0 and length > 0:
        _lib.check(lib.sep_istft_f32(plan.handle, ptr(raw), batch, frames, ptr(out), mem,
                                     current_stream(mem, spec if dev else None)), "sep_istft_f32")
    if single:
        out = out[0]
    return out if dev else out.astype(np.float64)


def recombine_istft(cleaned, phase, n_src, size=256, shift=128, window=None):
    """spec_c = cleaned_c * exp(1j * phase) (uPIT_baseline.ipynb:1385-1388, cell 41)
    followed by istft per source (:1401-1402), fused: the complex spectra are
    formed in registers.  cleaned [B, T, C*F] (the model output = mask_c * |X|,
    cell 29 :1087-1090), phase [B, T, F] -> waves [B, C, L] float32."""
    plan = get_plan(size, shift, window, True)
    lib = _lib.load()
    dev = is_device_tensor(cleaned)
    if dev:
        import torch

        c = require_f32_cuda(cleaned, "cleaned")
        p = require_f32_cuda(phase, "phase")
    else:
        c, p = as_f32_host(cleaned), as_f32_host(phase)
    batch, frames = int(c.shape[0]), int(c.shape[1])
    if tuple(c.shape) != (batch, frames, n_src * plan.bins) or tuple(p.shape) != (batch, frames, plan.bins):
        raise ValueError("cleaned must be [B, T, C*F] and phase [B, T, F]")
    length = plan.istft_samples(frames)
    mem = mem_kind(c, p)
    if dev:
        out = torch.zeros((batch, n_src, length), dtype=torch.float32, device=c.device)
    else:
        out = np.zeros((batch, n_src, length), dtype=np.float32)
    _lib.check(lib.sep_recombine_istft_f32(plan.handle, ptr(c), ptr(p), batch, n_src, frames, ptr(out),
                                           mem, current_stream(mem, c if dev else None)),
               "sep_recombine_istft_f32")
    return out
