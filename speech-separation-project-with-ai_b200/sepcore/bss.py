"""BSS Eval v4 on the GPU with museval's names.

Reference call site: metrics/evaluate_metrics.py:79-81
    museval.metrics.bss_eval(reference_stack, estimated_stack, window=np.inf, hop=np.inf,
                             compute_permutation=True)
museval is a third-party dependency of the reference that is neither vendored nor installable here, so
parity with museval itself is UNPINNED; the CUDA path is checked against a float64 restatement of museval's
published algorithm (the test-side CPU checker).  See csrc/bss.cu for how the criteria are computed without ever
forming the 512-tap projections in the time domain.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._buffers import as_f32_host, ptr
from .scoring import _pack


def bss_eval_batch(refs, ests, n_src=None, filters_len=512):
    """Scores a ragged batch.  refs / ests: lists (one entry per utterance) of [C, n_b] arrays (or lists of C
    1-D arrays), every signal of utterance b cut to the same length.  Returns dict:
    sdr / isr / sir / sar [B, C, C] ([jtrue][jest], dB), perm [B, C] (perm[jtrue] = jest, chosen by mean SIR),
    sdr_selected [B, C], value [B] (mean SDR of the selection with eval_sdr's NaN fallback)."""
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
    lengths = np.asarray(lengths, dtype=np.int64)
    width = int(lib.sep_bss_eval_row_width(n_src))
    rows = np.empty((batch, width), dtype=np.float64)
    i64p = C.POINTER(C.c_int64)
    _lib.check(lib.sep_bss_eval_f32(ptr(r_flat), ptr(e_flat), r_off.ctypes.data_as(i64p), e_off.ctypes.data_as(i64p),
                                    lengths.ctypes.data_as(i64p), batch, n_src, int(r_flat.size), int(e_flat.size),
                                    int(filters_len), ptr(rows), _lib.MEM_HOST, None), "sep_bss_eval_f32")
    cc = n_src * n_src
    perms = np.array(list(itertools.permutations(range(n_src))))
    return {
        "sdr": rows[:, 0:cc].reshape(batch, n_src, n_src),
        "isr": rows[:, cc:2 * cc].reshape(batch, n_src, n_src),
        "sir": rows[:, 2 * cc:3 * cc].reshape(batch, n_src, n_src),
        "sar": rows[:, 3 * cc:4 * cc].reshape(batch, n_src, n_src),
        "perm": perms[rows[:, 4 * cc].astype(np.int64)],
        "value": rows[:, 4 * cc + 1],
        "sdr_selected": rows[:, 4 * cc + 2:4 * cc + 2 + n_src],
    }


def bss_eval(reference_sources, estimated_sources, window=np.inf, hop=np.inf, compute_permutation=False,
             filters_len=512, framewise_filters=False, bsseval_sources_version=False):
    """museval.metrics.bss_eval for the configuration the reference uses: ONE window (window / hop >= the signal
    length, e.g. np.inf), time-invariant filters, images criteria, single-channel sources.

    reference_sources / estimated_sources: [nsrc, nsampl] or [nsrc, nsampl, 1].
    Returns (SDR, ISR, SIR, SAR, perm) with the metric arrays shaped [nsrc, 1] like museval's [nsrc, nwin]."""
    ref = np.asarray(reference_sources)
    est = np.asarray(estimated_sources)
    if ref.ndim == 3:
        if ref.shape[2] != 1:
            raise NotImplementedError("bss_eval: only single-channel sources are built (the reference's data is mono)")
        ref, est = ref[:, :, 0], est.reshape(est.shape[0], est.shape[1])
    if ref.ndim != 2 or ref.shape != est.shape:
        raise ValueError("reference_sources and estimated_sources must both be [nsrc, nsampl(, 1)]")
    nsampl = ref.shape[1]
    if framewise_filters or bsseval_sources_version or window < nsampl or hop < nsampl:
        raise NotImplementedError("bss_eval: only window = hop >= nsampl (one window), framewise_filters=False, "
                                  "bsseval_sources_version=False -- the reference's call -- is built")
    res = bss_eval_batch([ref], [est], ref.shape[0], filters_len)
    nsrc = ref.shape[0]
    perm = res["perm"][0] if compute_permutation else np.arange(nsrc)
    dum = np.arange(nsrc)
    pick = lambda key: res[key][0][dum, perm][:, None]
    return pick("sdr"), pick("isr"), pick("sir"), pick("sar"), perm
