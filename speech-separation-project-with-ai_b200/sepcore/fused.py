"""The fused hot path: STFT -> mask -> iSTFT overlap-add -> PIT-MSE -> SI-SDR/SDR
in one pass over HBM (spectra never leave the SM).

Reference chain: stft parallel_stft.py:146-196; |X| and PSA
labels :262-272; mask multiply uPIT_baseline.ipynb:1087-1088 (cell 29); phase
recombination :1385-1388 (cell 41); istft :1269-1307 (cell 39); pit_loss
:1023-1059 (cell 28); si_sdr / permute_si_sdr metrics/evaluate_metrics.py:14-34.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from ._buffers import as_f32_host, current_stream, is_device_tensor, mem_kind, ptr, require_f32_cuda
from .plan import get_plan


def score_layout(n_src):
    """Column slices of one `scores` row (see include/sepcore.h)."""
    cc, p = n_src * n_src, math.factorial(n_src)
    o = 0
    lay = {}
    for name, width in (("pit_pair", cc), ("pit_costs", p), ("pit_perm", 1), ("pit_loss", 1),
                        ("si_pair", cc), ("si_best", 1), ("si_perm", 1),
                        ("sdr_pair", cc), ("sdr_best", 1), ("sdr_perm", 1)):
        lay[name] = slice(o, o + width)
        o += width
    lay["stride"] = o
    return lay


def parse_scores(scores, n_src):
    lay = score_layout(n_src)
    batch = scores.shape[0]
    out = {}
    for name in ("pit_pair", "si_pair", "sdr_pair"):
        out[name] = scores[:, lay[name]].reshape(batch, n_src, n_src)
    out["pit_costs"] = scores[:, lay["pit_costs"]]
    for name in ("pit_perm", "pit_loss", "si_best", "si_perm", "sdr_best", "sdr_perm"):
        out[name] = scores[:, lay[name].start]
    return out


def workspace_bytes(batch, n_src, n_samples, size=256, shift=128, window=None):
    """Device scratch needed by one fused call (see `workspace=` below)."""
    import ctypes as C

    plan = get_plan(size, shift, window, True)
    out = C.c_int64()
    _lib.check(_lib.load().sep_fused_workspace_bytes(plan.handle, int(batch), int(n_src),
                                                     int(n_samples), C.byref(out)))
    return out.value


def separate_and_score(mix, masks, refs=None, frame_lengths=None, valid_samples=None,
                       size=256, shift=128, window=None, want_est=True, out=None, workspace=None, push=None):
    """One fused pass.

    mix [B, N] float32; masks [B, C, T, F] float32 with T = frames(N), F = size/2+1;
    refs [B, C, N] or None; frame_lengths [B] float32 (valid frames, the `length`
    row of y_true) or None; valid_samples [B] int32 (samples scored) or None.

    Returns dict: est [B, C, N] (if want_est), and when refs are given the parsed
    per-utterance scores plus `sums` = [sum pit_loss, sum si_best, sum sdr_best, B].
    numpy in -> numpy out (library stages copies); CUDA tensors in -> CUDA tensors out.
    `out` may carry preallocated 'est' / 'scores' / 'sums' buffers (CUDA tensors in
    device mode; numpy arrays -- e.g. views of pinned memory -- in host mode).
    `workspace` (device mode): a ZERO-FILLED uint8 CUDA tensor of `workspace_bytes(...)`
    bytes (`torch.zeros`; the library leaves it zero-filled, so one buffer can be
    reused by successive calls on one stream); with it the call allocates nothing, so
    it can be captured into a CUDA graph.
    `push` (device mode, refs given): `sepcore.distributed.PeerSums.target(slot)` -- the kernel also stores this
    call's batch sums into every rank's inbox over NVLink (per-batch reduction without a collective call).
    """
    plan = get_plan(size, shift, window, True)
    lib = _lib.load()
    dev = is_device_tensor(mix)
    if dev:
        import torch

        m = require_f32_cuda(mix, "mix")
        k = require_f32_cuda(masks, "masks")
        r = None if refs is None else require_f32_cuda(refs, "refs")
        fl = None if frame_lengths is None else require_f32_cuda(frame_lengths, "frame_lengths")
        vs = valid_samples
        if vs is not None and (vs.dtype != torch.int32 or not vs.is_contiguous()):
            raise ValueError("valid_samples must be a contiguous int32 CUDA tensor")
    else:
        m, k = as_f32_host(mix), as_f32_host(masks)
        r = None if refs is None else as_f32_host(refs)
        fl = None if frame_lengths is None else as_f32_host(frame_lengths)
        vs = None if valid_samples is None else np.ascontiguousarray(valid_samples, dtype=np.int32)
    if m.ndim != 2 or k.ndim != 4:
        raise ValueError("mix must be [B, N] and masks [B, C, T, F]")
    batch, n = int(m.shape[0]), int(m.shape[1])
    n_src = int(k.shape[1])
    frames = plan.frames(n)
    if tuple(int(v) for v in k.shape) != (batch, n_src, frames, plan.bins):
        raise ValueError("masks must be [B=%d, C, T=%d, F=%d], got %r"
                         % (batch, frames, plan.bins, tuple(k.shape)))
    if r is not None and tuple(int(v) for v in r.shape) != (batch, n_src, n):
        raise ValueError("refs must be [B, C, N]")
    stride = score_layout(n_src)["stride"]
    mem = mem_kind(m, k, r, fl, vs)
    out = out or {}
    est = scores = sums = None
    if dev:
        if want_est:
            est = out.get("est")
            if est is None:
                est = torch.empty((batch, n_src, n), dtype=torch.float32, device=m.device)
        if r is not None:
            scores = out.get("scores")
            if scores is None:
                scores = torch.empty((batch, stride), dtype=torch.float64, device=m.device)
            sums = out.get("sums")
            if sums is None:
                sums = torch.empty((4,), dtype=torch.float64, device=m.device)
    else:
        if want_est:
            est = out.get("est")
            if est is None:
                est = np.empty((batch, n_src, n), dtype=np.float32)
        if r is not None:
            scores = out.get("scores")
            if scores is None:
                scores = np.empty((batch, stride), dtype=np.float64)
            sums = out.get("sums")
            if sums is None:
                sums = np.empty((4,), dtype=np.float64)
    for name, buf, shape in (("est", est, (batch, n_src, n)), ("scores", scores, (batch, stride)),
                             ("sums", sums, (4,))):
        if buf is not None and tuple(int(v) for v in buf.shape) != shape:
            raise ValueError("out[%r] must have shape %r" % (name, shape))
    ws_ptr, ws_bytes = None, 0
    if workspace is not None and dev:
        ws_ptr, ws_bytes = ptr(workspace), int(workspace.numel() * workspace.element_size())
    if push is not None:
        if not dev or r is None:
            raise ValueError("push needs CUDA tensors and refs")
        peers, world, rank, slot, slots = push
        _lib.check(lib.sep_fused_separate_push_f32(plan.handle, ptr(m), ptr(k), ptr(r), ptr(fl), ptr(vs),
                                                   batch, n_src, n, ptr(est), ptr(scores), ptr(sums),
                                                   ws_ptr, ws_bytes, ptr(peers), int(world), int(rank), int(slot),
                                                   int(slots), current_stream(mem, m)),
                   "sep_fused_separate_push_f32")
    else:
        _lib.check(lib.sep_fused_separate_ws_f32(plan.handle, ptr(m), ptr(k), ptr(r), ptr(fl), ptr(vs),
                                                 batch, n_src, n, ptr(est), ptr(scores), ptr(sums),
                                                 ws_ptr, ws_bytes, mem,
                                                 current_stream(mem, m if dev else None)),
                   "sep_fused_separate_ws_f32")
    res = {}
    if want_est:
        res["est"] = est
    if r is not None:
        res.update(parse_scores(scores, n_src))
        res["scores"] = scores
        res["sums"] = sums
    return res
