"""SI-SDR / SDR scoring on the GPU with the reference's function names.

Reference: metrics/evaluate_metrics.py:14-92.  Signals are read once; Gram
statistics are accumulated in float64 on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._buffers import as_f32_host, current_stream, is_device_tensor, ptr


def _pack(signals):
    """Concatenates 1-D float32 signals, each start aligned to 4 elements (16 B)
    so the kernel can use 128-bit loads; returns (flat, offsets)."""
    offs, total = [], 0
    for s in signals:
        offs.append(total)
        total += (len(s) + 3) & ~3
    flat = np.zeros(max(total, 1), dtype=np.float32)
    for s, o in zip(signals, offs):
        flat[o:o + len(s)] = s
    return flat, np.asarray(offs, dtype=np.int64)


def score_batch(refs, ests, n_src=None):
    """Scores a ragged batch.

    refs / ests: lists (one entry per utterance) of [C, n_b] arrays (or lists of
    C 1-D arrays); every signal of utterance b must already be cut to the same
    length n_b (use `truncate_to_min_len`).  Returns dict with si_pair [B,C,C]
    (dB, [i][j] = SI-SDR(ref_j, est_i)), si_best [B], si_perm [B], sdr_pair,
    sdr_best, sdr_perm, sums = (sum si_best, sum sdr_best, B).
    """
    lib = _lib.load()
    batch = len(refs)
    if batch == 0 or len(ests) != batch:
        raise ValueError("refs and ests must be equally long, non-empty lists")
    r_sig, e_sig, lengths = [], [], []
    for r, e in zip(refs, ests):
        r = [as_f32_host(v).reshape(-1) for v in r]
        e = [as_f32_host(v).reshape(-1) for v in e]
        if n_src is None:
            n_src = len(r)
        n = len(r[0])
        if len(r) != n_src or len(e) != n_src or any(len(v) != n for v in r + e):
            raise ValueError("every utterance needs C references and C estimates of one length")
        r_sig += r
        e_sig += e
        lengths.append(n)
    r_flat, r_off = _pack(r_sig)
    e_flat, e_off = _pack(e_sig)
    return _score_flat(lib, r_flat, e_flat, r_off, e_off, np.asarray(lengths, dtype=np.int64),
                       batch, n_src, _lib.MEM_HOST, None)


def score_flat_device(refs_flat, ests_flat, ref_offsets, est_offsets, lengths, n_src):
    """Throughput mode: flat float32 CUDA tensors plus host offset / length arrays."""
    lib = _lib.load()
    if not (is_device_tensor(refs_flat) and is_device_tensor(ests_flat)):
        raise ValueError("score_flat_device needs CUDA tensors")
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    return _score_flat(lib, refs_flat, ests_flat,
                       np.ascontiguousarray(ref_offsets, dtype=np.int64),
                       np.ascontiguousarray(est_offsets, dtype=np.int64), lengths, len(lengths),
                       n_src, _lib.MEM_DEVICE, current_stream(_lib.MEM_DEVICE, refs_flat))


def _score_flat(lib, r_flat, e_flat, r_off, e_off, lengths, batch, n_src, mem, stream):
    width = 2 * n_src * n_src + 4
    if mem == _lib.MEM_DEVICE:
        import torch

        scores = torch.empty((batch, width), dtype=torch.float64, device=r_flat.device)
        sums = torch.empty((3,), dtype=torch.float64, device=r_flat.device)
        total_r, total_e = r_flat.numel(), e_flat.numel()
    else:
        scores = np.empty((batch, width), dtype=np.float64)
        sums = np.empty((3,), dtype=np.float64)
        total_r, total_e = r_flat.size, e_flat.size
    i64p = C.POINTER(C.c_int64)
    _lib.check(lib.sep_score_batch_f32(ptr(r_flat), ptr(e_flat), r_off.ctypes.data_as(i64p),
                                       e_off.ctypes.data_as(i64p), lengths.ctypes.data_as(i64p),
                                       batch, n_src, int(total_r), int(total_e), ptr(scores),
                                       ptr(sums), mem, stream), "sep_score_batch_f32")
    cc = n_src * n_src
    return {
        "si_pair": scores[:, :cc].reshape(batch, n_src, n_src),
        "si_best": scores[:, cc],
        "si_perm": scores[:, cc + 1],
        "sdr_pair": scores[:, cc + 2:2 * cc + 2].reshape(batch, n_src, n_src),
        "sdr_best": scores[:, 2 * cc + 2],
        "sdr_perm": scores[:, 2 * cc + 3],
        "sums": sums,
        "raw": scores,
    }


def truncate_to_min_len(ref_s1, ref_s2, est_s1, est_s2):
    """All four signals cut to min(len(ref_s1), len(est_s1)) (evaluate_metrics.py:46-48)."""
    n = min(np.size(ref_s1), np.size(est_s1))
    return ref_s1[:n], ref_s2[:n], est_s1[:n], est_s2[:n]


# ------------------------------------------------------------------ reference names
def pow_np_norm(signal):
    """Squared 2-norm (evaluate_metrics.py:14-17)."""
    return pow_norm(signal, signal)


def pow_norm(s1, s2):
    """Inner product sum(s1 * s2) (evaluate_metrics.py:19-20), float64 accumulation
    on the device; returned as float32 like the reference's float32 inputs give."""
    lib = _lib.load()
    dev = is_device_tensor(s1)
    if dev:
        import torch

        a, b = s1.reshape(-1).contiguous().float(), s2.reshape(-1).contiguous().float()
        out = torch.empty((1,), dtype=torch.float64, device=a.device)
        mem = _lib.MEM_DEVICE
    else:
        a, b = as_f32_host(s1).reshape(-1), as_f32_host(s2).reshape(-1)
        out = np.empty((1,), dtype=np.float64)
        mem = _lib.MEM_HOST
    if a.shape != b.shape:
        raise ValueError("pow_norm needs equally long signals")
    n = int(a.shape[0])
    _lib.check(lib.sep_dot_f32(ptr(a), ptr(b), n, ptr(out), mem,
                               current_stream(mem, a if dev else None)), "sep_dot_f32")
    return out[0] if dev else np.float32(out[0])


def si_sdr(original, estimated):
    """SI-SDR in dB (evaluate_metrics.py:22-26)."""
    res = score_batch([[as_f32_host(original).reshape(-1)]], [[as_f32_host(estimated).reshape(-1)]], 1)
    return np.float32(res["si_pair"][0, 0, 0])


def permute_si_sdr(ref1, ref2, est1, est2):
    """0.5 * max over the two speaker assignments (evaluate_metrics.py:28-34)."""
    res = score_batch([[ref1, ref2]], [[est1, est2]], 2)
    return np.float32(res["si_best"][0])
