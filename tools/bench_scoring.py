"""BASELINE config 3: SI-SDR/SDR over a synthetic 3000-utterance wsj0-2mix-shaped test set
(SURVEY.md 8d): lengths 8000*U[2,10] s, refs 0.1*N(0,1), est = ref + noise at U[-5,20] dB,
half of the utterances with swapped estimates.  Prints one JSON line (device-resident
throughput of sep_score_batch_f32 and its fraction of the HBM roofline: 16 B per sample
index -> 128 000 B per audio-second).

Multi-GPU (strong scaling): `python -m torch.distributed.run --nproc-per-node N tools/bench_scoring.py`
shards the SAME 3000 utterances by load (sepcore.distributed.shard_by_load), every rank scores its shard and the
[si_sdr_sum, sdr_sum, n] rows are all-reduced once per step; time = max over ranks."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]

import numpy as np
import torch

import sepcore
from sepcore import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=3000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--check", type=int, default=8, help="utterances verified against the oracle")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    from sepcore.distributed import shard_by_load
    dist.init_process_group("nccl", device_id=dev)

rng = np.random.default_rng(4)
all_lengths = (8000 * rng.uniform(2, 10, size=args.utts)).astype(np.int64)
lengths = all_lengths if world == 1 else all_lengths[shard_by_load(all_lengths, world)[rank]]
n_mine = len(lengths)
offs, total = [], 0
for n in lengths:
    for _ in range(2):
        offs.append(total)
        total += (int(n) + 3) & ~3
offs = np.asarray(offs, dtype=np.int64)
gen = torch.Generator(device=dev).manual_seed(5 + rank)
refs = 0.1 * torch.randn(total, device=dev, generator=gen)
snr = torch.from_numpy(rng.uniform(-5, 20, size=n_mine)).to(dev)
noise = 0.1 * torch.randn(total, device=dev, generator=gen)
ests = refs.clone()
scale = torch.zeros(total, device=dev)
for b, n in enumerate(lengths):
    for c in range(2):
        o = offs[2 * b + c]
        scale[o:o + n] = 10 ** (-float(snr[b]) / 20)
ests += scale * noise
est_offs = offs.copy()
swap = np.arange(n_mine) % 2 == 1
est_offs[0::2][swap], est_offs[1::2][swap] = offs[1::2][swap], offs[0::2][swap]
del scale, noise
torch.cuda.synchronize()

res = sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
torch.cuda.synchronize()
# parity spot check against the oracle (float32 reference arithmetic)
from oracle import signal_path as oracle
worst = 0.0
for b in range(min(args.check, n_mine)):
    n = int(lengths[b])
    r = [refs[offs[2 * b + c]:offs[2 * b + c] + n].cpu().numpy() for c in range(2)]
    e = [ests[est_offs[2 * b + c]:est_offs[2 * b + c] + n].cpu().numpy() for c in range(2)]
    v, p, _, _ = oracle.permute_si_sdr_detail(r[0], r[1], e[0], e[1])
    assert int(res["si_perm"][b].item()) == p, (b, p)
    worst = max(worst, abs(float(res["si_best"][b].item()) - float(v)))
assert worst < 0.01, worst

_lib.profile_enable(False)
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = sepcore.launch_count()
for _ in range(3):
    sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
torch.cuda.synchronize()
launches = (sepcore.launch_count() - l0) // 3
if world > 1:
    dist.all_reduce(res["sums"].clone())          # warms NCCL up
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.all_reduce(torch.zeros(1, device=dev))   # device-side rendezvous in front of the start event
start.record()
for _ in range(args.steps):
    out = sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
    if world > 1:
        dist.all_reduce(out["sums"])              # dataset sums over all shards (evaluate_metrics.py:53,90)
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / args.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res = out
# the dominant kernel alone (score_chunk_kernel), bracketed by CUDA events inside the library
_lib.profile_enable(True)
for _ in range(args.steps):
    sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
torch.cuda.synchronize()
k_ms, k_cnt = _lib.profile_collect()
_lib.profile_enable(False)
k_ms /= max(k_cnt, 1)
audio_s = float(all_lengths.sum()) / 8000.0          # the whole set: all ranks together
bytes_alg = 16.0 * float(lengths.sum())               # this rank's shard (kernel roofline)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
line = {
    "metric": "audio-sec/sec SI-SDR+SDR scoring (evaluate_metrics)", "value": audio_s / (ms * 1e-3),
    "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms, "scaling": "strong",
    "dtype": "f32 in, f64 accumulate",
    "config": {"workload": "cfg3: %d utterances, 2-10 s @ 8 kHz, 2 refs + 2 ests (%.2f GB > L2)"
                           % (args.utts, 16.0 * float(all_lengths.sum()) / 1e9), "launches_per_step": launches,
               "sharding": "by load, %d utterances on rank 0" % n_mine},
    "roofline": {"bound": "hbm", "achieved": bytes_alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / peak, "kernel": "score_chunk_kernel<2>",
                 "kernel_ms": k_ms, "launches_timed": k_cnt, "bytes_per_launch": bytes_alg,
                 "achieved_whole_call": 16.0 * float(all_lengths.sum()) / world / (ms * 1e-3) / 1e9,
                 "frac_whole_call": 16.0 * float(all_lengths.sum()) / world / (ms * 1e-3) / 1e9 / peak,
                 "note": "kernel: CUDA events around score_chunk_kernel; whole call adds finalize + sums + the host "
                         "metadata upload (offsets, lengths, chunk table) of every call"},
    "check": {"oracle_utts": args.check, "max_abs_db_err": worst,
              "mean_si_sdr_db": float(res["sums"][0].item() / res["sums"][2].item())},
}
if rank == 0:
    print(json.dumps(line))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
