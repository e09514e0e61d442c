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
