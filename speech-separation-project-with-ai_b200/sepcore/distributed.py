"""Utterance sharding across GPUs and the single collective of the path.

Every utterance is independent through STFT -> mask -> iSTFT -> PIT / SI-SDR; the
only cross-utterance operations in the reference are the batch SUM in pit_loss
(uPIT_baseline.ipynb:1055, cell 28) and the dataset MEAN in eval_si_sdr /
eval_sdr (metrics/evaluate_metrics.py:53, :90).  So: one process per GPU, a
contiguous (or load-balanced) block of utterances per rank, no data-path
collective, and ONE all-reduce(sum) of [loss_sum, si_sdr_sum, sdr_sum, n_utt]
(32 bytes, float64) per batch -- NCCL over NVLink on the GPU box, gloo in the
CPU tests.  torch.distributed is plumbing here, nothing else.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_load(lengths, world):
    """Greedy longest-first assignment balancing total samples per rank (ragged
    test sets, BASELINE config 3).  Returns a list of index arrays, each sorted
    so that per-rank reductions run in utterance order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    loads = np.zeros(world, dtype=np.int64)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        bins[r].append(int(i))
        loads[r] += lengths[i]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def all_reduce_sums(sums, group=None):
    """In-place all-reduce(sum) of the per-rank [loss_sum, si_sdr_sum, sdr_sum, n_utt]
    tensor (float64).  CUDA tensor -> NCCL, CPU tensor -> gloo."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def gather_per_utterance(values, counts, group=None):
    """All-gather of per-utterance scalars (for means that must be bit-identical
    for every world size: the caller then reduces in fixed utterance order).
    `values` is this rank's 1-D tensor, `counts` the per-rank lengths."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return values
    width = int(max(counts))
    padded = torch.zeros(width, dtype=values.dtype, device=values.device)
    padded[: values.numel()] = values
    parts = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def means_from_sums(sums):
    """(mean pit loss per utterance, mean SI-SDR, mean SDR) from the reduced sums."""
    n = float(sums[3])
    return float(sums[0]) / n, float(sums[1]) / n, float(sums[2]) / n


class _DevArray:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerSums:
    """Per-batch [loss, SI-SDR, SDR, n] sums of ALL ranks without a collective call on the step
    (sep_fused_separate_push_f32): every rank owns an inbox of `slots` x world rows in peer-visible device
    memory; the fused kernel of rank r, in the epilogue that adds its batch sums, stores them into row r of
    the step's slot in EVERY rank's inbox over NVLink and bumps that inbox's arrival counter.  Readers add the
    `world` rows after their own synchronisation -- an all-gather by one-sided puts plus a local reduction.
    The NCCL all-reduce (`all_reduce_sums`) stays as the checked baseline.

    Setup exchanges 64-byte CUDA IPC handles through torch.distributed (any backend); one process per GPU
    on one node (NVLink / NVSwitch P2P)."""

    def __init__(self, slots, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        lib = _lib.load()
        self.lib = lib
        self.slots = int(slots)
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.dev = torch.device("cuda", torch.cuda.current_device())
        nbytes = self.slots * self.world * 32 + self.slots * 8
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        _lib.check(lib.sep_peer_alloc(nbytes, C.byref(own), handle), "sep_peer_alloc")
        self._own = own.value
        handles = [handle.raw]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
        self._opened, ptrs = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self._own)
                continue
            peer = C.c_void_p()
            _lib.check(lib.sep_peer_open(h, C.byref(peer)), "sep_peer_open")
            self._opened.append(peer.value)
            ptrs.append(peer.value)
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.dev)      # the DEVICE array of inbox pointers
        self.rows = torch.as_tensor(_DevArray(self._own, (self.slots, self.world, 4), "<f8"), device=self.dev)
        self.arrived = torch.as_tensor(_DevArray(self._own + self.slots * self.world * 32, (self.slots,), "<i8"),
                                       device=self.dev)
        if self.world > 1:
            dist.barrier(group=group)            # every inbox is mapped everywhere before anyone pushes

    def target(self, slot):
        """(peer pointer array, world, rank, slot, slots) for `separate_and_score(push=...)`."""
        return (self.peer_ptrs, self.world, self.rank, int(slot) % self.slots, self.slots)

    def reduced(self):
        """[slots, 4]: the sums of every step over all ranks (valid once arrived[slot] counts `world` per use)."""
        return self.rows.sum(dim=1)

    def close(self):
        for p in self._opened:
            self.lib.sep_peer_close(p)
        self._opened = []
        if self._own:
            self.rows = self.arrived = None
            self.lib.sep_peer_free(self._own)
            self._own = None
