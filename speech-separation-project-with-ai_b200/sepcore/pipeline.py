"""Pipelined host-to-host execution of the fused hot path.

`separate_and_score(numpy...)` (drop-in mode) is synchronous: H2D, kernels, D2H
run back to back on one stream and the call returns when the results are on
the host.  For streams of batches (dataset scoring, inference) the PCIe link is
the bottleneck, and it is full duplex: `HostPipeline` keeps `depth` batches in
flight on three streams -- copy-in, compute, copy-out -- so the H2D of batch
k+1, the kernels of batch k and the D2H of batch k-1 overlap.  Everything is
plain CUDA streams and events (torch is the plumbing); the compute step is the
same C-ABI call, sep_fused_separate_ws_f32, with caller-owned buffers.
"""
from __future__ import annotations

import numpy as np

from . import fused


class HostPipeline:
    """depth-slot pipeline for batches of fixed shape.

    submit(mix, masks, refs=None, ...) -> ticket;  result(ticket) -> dict of numpy
    arrays (views of pinned host memory owned by the slot, valid until the slot is
    reused `depth` submits later).  Inputs may be numpy arrays or CPU tensors;
    pinned inputs are copied asynchronously, pageable ones synchronously (CUDA rule).

    pcm16=True: the waveforms cross PCIe in the format they have on disk -- mixture and references
    as int16 PCM (decoded on the device exactly like wavread / librosa.load do: x / 32768), the
    estimates as the int16 samples `audiowrite(wav, path, rate, normalize=True)` would write
    (uPIT_baseline.ipynb:1403-1404), plus the per-row clipped counts: half the waveform bytes each way.

    device_masks=True: `masks` given to submit() is a CUDA tensor that is already on the device -- in the
    reference's pipeline the masks are the model's output (cell 29: `Multiply()([pred, inputs])`), produced
    where they are consumed, so only the waveforms cross PCIe.  The tensor is read in place on the compute
    stream (the caller's producer stream must have finished writing it, e.g. via torch's stream ordering).
    """

    def __init__(self, batch, n_src, n_samples, size=256, shift=128, window=None, depth=3,
                 scored=True, want_est=True, device=None, pcm16=False, device_masks=False):
        import torch

        self.torch = torch
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.kw = dict(size=size, shift=shift, window=window, want_est=want_est)
        self.scored, self.want_est, self.depth, self.pcm16 = scored, want_est, int(depth), bool(pcm16)
        self.device_masks = bool(device_masks)
        from .plan import get_plan

        plan = get_plan(size, shift, window, True)
        frames, bins = plan.frames(n_samples), plan.bins
        self.n_src = n_src
        stride = fused.score_layout(n_src)["stride"]
        nbytes = fused.workspace_bytes(batch, n_src, n_samples, size, shift, window)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.slots = []
        for _ in range(self.depth):
            s = {
                "mix": torch.empty((batch, n_samples), **f32),
                "masks": None if self.device_masks else torch.empty((batch, n_src, frames, bins), **f32),
                "refs": torch.empty((batch, n_src, n_samples), **f32) if scored else None,
                "workspace": torch.zeros(nbytes, dtype=torch.uint8, device=self.dev),
                "out": {}, "host": {},
                "ev_in": torch.cuda.Event(), "ev_done": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                "busy": False,
            }
            if self.pcm16:
                i16 = dict(dtype=torch.int16, device=self.dev)
                s["mix_i16"] = torch.empty((batch, n_samples), **i16)
                s["refs_i16"] = torch.empty((batch, n_src, n_samples), **i16) if scored else None
            if want_est:
                s["out"]["est"] = torch.empty((batch, n_src, n_samples), **f32)
                if self.pcm16:
                    s["est_i16"] = torch.empty((batch * n_src, n_samples), dtype=torch.int16, device=self.dev)
                    s["clipped"] = torch.empty((batch * n_src,), dtype=torch.int64, device=self.dev)
                    s["host"]["est"] = torch.empty((batch, n_src, n_samples), dtype=torch.int16).pin_memory()
                    s["host"]["clipped"] = torch.empty((batch, n_src), dtype=torch.int64).pin_memory()
                else:
                    s["host"]["est"] = torch.empty((batch, n_src, n_samples), dtype=torch.float32).pin_memory()
            if scored:
                s["out"]["scores"] = torch.empty((batch, stride), dtype=torch.float64, device=self.dev)
                s["out"]["sums"] = torch.empty((4,), dtype=torch.float64, device=self.dev)
                s["host"]["scores"] = torch.empty((batch, stride), dtype=torch.float64).pin_memory()
                s["host"]["sums"] = torch.empty((4,), dtype=torch.float64).pin_memory()
            self.slots.append(s)
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_run = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self.count = 0
        wave_bytes = 2 if self.pcm16 else 4
        self.h2d_bytes = self.slots[0]["mix"].numel() * wave_bytes \
            + (0 if self.device_masks else self.slots[0]["masks"].numel() * 4) \
            + (self.slots[0]["refs"].numel() * wave_bytes if scored else 0)
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.slots[0]["host"].values())

    def _as_tensor(self, x):
        return x if self.torch.is_tensor(x) else self.torch.from_numpy(np.ascontiguousarray(x))

    def submit(self, mix, masks, refs=None):
        torch = self.torch
        ticket = self.count
        slot = self.slots[ticket % self.depth]
        if slot["busy"]:
            slot["ev_out"].synchronize()          # the slot's previous results have reached the host
        slot["busy"] = True
        with torch.cuda.stream(self.s_in):
            slot["mix_i16" if self.pcm16 else "mix"].copy_(self._as_tensor(mix), non_blocking=True)
            if self.device_masks:
                if not masks.is_cuda:
                    raise ValueError("device_masks=True: masks must be a CUDA tensor")
                slot["masks"] = masks
            else:
                slot["masks"].copy_(self._as_tensor(masks), non_blocking=True)
            if self.scored:
                slot["refs_i16" if self.pcm16 else "refs"].copy_(self._as_tensor(refs), non_blocking=True)
            slot["ev_in"].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(slot["ev_in"])
            if self.pcm16:
                from . import audio_io

                audio_io.pcm16_to_float32(slot["mix_i16"], out=slot["mix"])
                if self.scored:
                    audio_io.pcm16_to_float32(slot["refs_i16"], out=slot["refs"])
            out = dict(slot["out"])
            fused.separate_and_score(slot["mix"], slot["masks"], slot["refs"], out=out,
                                     workspace=slot["workspace"], **self.kw)
            if self.pcm16 and self.want_est:
                audio_io.audiowrite_int16(slot["out"]["est"].reshape(slot["est_i16"].shape), True,
                                          out=slot["est_i16"], clipped=slot["clipped"])
            slot["ev_done"].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["ev_done"])
            for key, host in slot["host"].items():
                if self.pcm16 and key == "est":
                    host.copy_(slot["est_i16"].reshape(host.shape), non_blocking=True)
                elif key == "clipped":
                    host.copy_(slot["clipped"].reshape(host.shape), non_blocking=True)
                else:
                    host.copy_(slot["out"][key], non_blocking=True)
            slot["ev_out"].record(self.s_out)
        self.count += 1
        return ticket

    def result(self, ticket):
        if ticket < self.count - self.depth or ticket >= self.count:
            raise ValueError("ticket %d is not in flight (slots are reused after %d submits)"
                             % (ticket, self.depth))
        slot = self.slots[ticket % self.depth]
        slot["ev_out"].synchronize()
        res = {}
        if self.want_est:
            res["est"] = slot["host"]["est"].numpy()
            if self.pcm16:
                res["clipped"] = slot["host"]["clipped"].numpy()
        if self.scored:
            scores = slot["host"]["scores"].numpy()
            res.update(fused.parse_scores(scores, self.n_src))
            res["scores"] = scores
            res["sums"] = slot["host"]["sums"].numpy()
        return res

    def drain(self):
        for s in self.slots:
            if s["busy"]:
                s["ev_out"].synchronize()
