"""CPU tests of the multi-GPU host logic: utterance sharding and the single
all-reduce of [loss_sum, si_sdr_sum, sdr_sum, n_utt], world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_shard_range_covers_everything():
    from sepcore import distributed as d
    for n in (0, 1, 7, 64, 3000):
        for world in (1, 2, 3, 8):
            spans = [d.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_range(10, 2, 2)


def test_shard_by_load_balances_samples():
    from sepcore import distributed as d
    rng = np.random.default_rng(4)
    lengths = rng.integers(16000, 80000, size=3000)
    for world in (2, 4, 8):
        bins = d.shard_by_load(lengths, world)
        assert sorted(np.concatenate(bins).tolist()) == list(range(3000))
        loads = [lengths[b].sum() for b in bins]
        assert (max(loads) - min(loads)) / np.mean(loads) < 0.01
        assert all(np.all(np.diff(b) > 0) for b in bins)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, per_utt, out):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from sepcore import distributed as d
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = d.shard_range(per_utt.shape[0], rank, world)
    mine = per_utt[lo:hi]
    sums = torch.tensor([mine[:, 0].sum(), mine[:, 1].sum(), mine[:, 2].sum(), float(hi - lo)],
                        dtype=torch.float64)
    d.all_reduce_sums(sums)
    gathered = d.gather_per_utterance(mine[:, 1].contiguous(),
                                      [d.shard_range(per_utt.shape[0], r, world)[1]
                                       - d.shard_range(per_utt.shape[0], r, world)[0] for r in range(world)])
    if rank == 0:
        out.put((sums.numpy().copy(), gathered.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_reduce_sums_world_size_2_gloo():
    """Same per-utterance results at 1 and 2 ranks: sums agree to float64 rounding and
    the gathered per-utterance values reproduce the single-process mean bit for bit."""
    from sepcore import distributed as d
    rng = np.random.default_rng(8)
    per_utt = torch.from_numpy(rng.standard_normal((13, 3)))          # odd count: uneven shards
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, per_utt, out)) for r in range(2)]
    for p in procs:
        p.start()
    sums, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = per_utt.numpy()
    assert np.allclose(sums[:3], want.sum(axis=0), rtol=1e-13, atol=1e-13) and sums[3] == 13
    assert np.array_equal(gathered, want[:, 1])
    assert np.mean(gathered) == np.mean(want[:, 1])
    loss, si, sdr = d.means_from_sums(sums)
    assert abs(si - want[:, 1].mean()) < 1e-13
