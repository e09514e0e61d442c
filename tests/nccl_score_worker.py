"""Worker of tests/test_multigpu.py::test_scoring_sharded_bit_identical: the ragged scorer on this rank's
`shard_by_load` block of one seeded set; all-reduce of the sums, all-gather of the per-utterance values."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import sepcore
    from sepcore import distributed as d

    out_path = sys.argv[1]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(99)
    n_utts = 40
    lens = rng.integers(3000, 40000, size=n_utts)
    refs, ests = [], []
    for i in range(n_utts):
        r = (0.1 * rng.standard_normal((2, lens[i]))).astype(np.float32)
        e = (r + 0.03 * rng.standard_normal((2, lens[i]))).astype(np.float32)
        if i % 3 == 0:
            e = e[::-1].copy()
        refs.append(r)
        ests.append(e)
    mine = d.shard_by_load(lens, world)[rank]
    res = sepcore.score_batch([refs[i] for i in mine], [ests[i] for i in mine], 2)
    sums = torch.from_numpy(np.asarray(res["sums"], dtype=np.float64)).to(dev)
    d.all_reduce_sums(sums)
    counts = [len(b) for b in d.shard_by_load(lens, world)]
    order = np.concatenate(d.shard_by_load(lens, world))
    got = {}
    for key in ("si_best", "si_perm", "sdr_best"):
        v = torch.from_numpy(np.asarray(res[key], dtype=np.float64)).to(dev)
        g = d.gather_per_utterance(v, counts).cpu().numpy()
        full = np.empty(n_utts)
        full[order] = g                                   # back to utterance order
        got[key] = full
    if rank == 0:
        np.savez(out_path, sums=sums.cpu().numpy(), **got)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
