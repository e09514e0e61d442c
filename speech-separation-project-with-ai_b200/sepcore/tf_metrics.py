"""Batched TF-style SI-SDR metric and loss of the reference's waveform models.

Reference: `SiSdr.update_state` / `result` / `reset_states` -- vq-vae_for_1d_data.ipynb cell 13
(:388-432); `custom_sisdr_loss` -- cell 14 (:457-469).  Layout: y_true [B, L + 1, 1] (last time
row = length, unused by the arithmetic), y_pred [B, L', 1]; label and prediction are cut to
min(L, L') (cell 13 :408-411).  Every utterance is one (reference, estimate) pair of the ragged
scoring kernel (score.cu): both tensors are read in place, once.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._buffers import as_f32_host, is_device_tensor
from .scoring import _score_flat


def sisdr_values(y_true, y_pred):
    """SI-SDR in dB of every batch element, [B] (float64)."""
    dev = is_device_tensor(y_true)
    if dev != is_device_tensor(y_pred):
        raise ValueError("y_true and y_pred must both be numpy arrays or both CUDA tensors")
    if dev:
        t = y_true if y_true.is_contiguous() else y_true.contiguous()
        p = y_pred if y_pred.is_contiguous() else y_pred.contiguous()
        t, p = t.float(), p.float()
    else:
        t, p = as_f32_host(y_true), as_f32_host(y_pred)
    if len(t.shape) != 3 or len(p.shape) != 3 or t.shape[2] != 1 or p.shape[2] != 1 or t.shape[0] != p.shape[0]:
        raise ValueError("expected y_true [B, L + 1, 1] and y_pred [B, L', 1]")
    batch, rows_t, rows_p = int(t.shape[0]), int(t.shape[1]), int(p.shape[1])
    n = min(rows_t - 1, rows_p)
    if n < 1:
        raise ValueError("empty sequences")
    r_off = np.arange(batch, dtype=np.int64) * rows_t
    e_off = np.arange(batch, dtype=np.int64) * rows_p
    lengths = np.full((batch,), n, dtype=np.int64)
    lib = _lib.load()
    if dev:
        from ._buffers import current_stream

        res = _score_flat(lib, t.reshape(-1), p.reshape(-1), r_off, e_off, lengths, batch, 1, _lib.MEM_DEVICE,
                          current_stream(_lib.MEM_DEVICE, t))
    else:
        res = _score_flat(lib, t.reshape(-1), p.reshape(-1), r_off, e_off, lengths, batch, 1, _lib.MEM_HOST, None)
    return res["si_pair"][:, 0, 0]


class SiSdr:
    """Keras-metric look-alike (name, update_state, result, reset_states)."""

    def __init__(self, name="Si-sdr", **kwargs):
        self.name = name
        self.reset_states()

    def update_state(self, y_true, y_pred, sample_weight=None):
        values = sisdr_values(y_true, y_pred)
        if sample_weight is not None:
            w = sample_weight
            if is_device_tensor(values):
                import torch

                w = torch.as_tensor(w, dtype=values.dtype, device=values.device)
            else:
                w = np.asarray(w, dtype=np.float64)
            values = values * w
        self.sdr = self.sdr + float(values.sum())
        self.count = self.count + float(values.shape[0])

    def result(self):
        return self.sdr / self.count

    def reset_states(self):
        self.sdr = 0.0
        self.count = 0.0


def custom_sisdr_loss(y_true, y_pred):
    """-mean SI-SDR over the batch (cell 14 :457-469)."""
    return -float(sisdr_values(y_true, y_pred).mean())
