"""Worker of tests/test_multigpu.py: one rank per GPU (torchrun), NCCL.

Every rank builds the SAME seeded utterance set, runs its `shard_range` block through the fused
path and the ragged scorer, all-reduces the [loss, SI-SDR, SDR, n] sums and all-gathers the
per-utterance values; rank 0 writes everything to the .npz given on the command line.  With
--world-check the file of a 1-rank run is compared by the test, not here."""
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
    rng = np.random.default_rng(2024)
    n_utts, n, n_src = 22, 12000, 2                      # 22 utterances: uneven blocks at 4 and 8 ranks
    refs = (0.1 * rng.standard_normal((n_utts, n_src, n))).astype(np.float32)
    mix = refs.sum(axis=1).astype(np.float32)
    frames = sepcore.get_plan(256, 128).frames(n)
    masks = rng.random((n_utts, n_src, frames, 129)).astype(np.float32)
    lo, hi = d.shard_range(n_utts, rank, world)
    counts = [d.shard_range(n_utts, r, world)[1] - d.shard_range(n_utts, r, world)[0] for r in range(world)]
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    per = {k: torch.zeros(0, dtype=torch.float64, device=dev) for k in ("pit_loss", "si_best", "sdr_best", "pit_perm")}
    if hi > lo:
        res = sepcore.separate_and_score(torch.from_numpy(mix[lo:hi]).to(dev), torch.from_numpy(masks[lo:hi]).to(dev),
                                         torch.from_numpy(refs[lo:hi]).to(dev), size=256, shift=128)
        sums = res["sums"].clone()
        per = {k: res[k].contiguous().clone() for k in per}
    d.all_reduce_sums(sums)
    gathered = {k: d.gather_per_utterance(v, counts) for k, v in per.items()}
    # the same reduction WITHOUT a collective call: the fused kernel pushes its sums into every rank's inbox
    peer = d.PeerSums(slots=4)
    if hi > lo:
        sepcore.separate_and_score(torch.from_numpy(mix[lo:hi]).to(dev), torch.from_numpy(masks[lo:hi]).to(dev),
                                   torch.from_numpy(refs[lo:hi]).to(dev), size=256, shift=128, push=peer.target(2))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                       # every rank's kernel (and its pushes) has completed
    pushed = peer.reduced()[2].cpu().numpy()
    arrived = int(peer.arrived[2].item())
    untouched = float(peer.rows[[0, 1, 3]].abs().sum().item())
    if world > 1:
        dist.barrier()
    peer.close()
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out_path, sums=sums.cpu().numpy(), push_sums=pushed, push_arrived=arrived, push_untouched=untouched,
                 **{k: v.cpu().numpy() for k, v in gathered.items()})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
