"""BASELINE config 5: learned conv encoder/decoder filterbank (N=256, L=16, stride 8), C=2,
64 x 4 s @ 8 kHz.  One JSON line: audio-s/s, HBM roofline fraction (algorithmic bytes per
utterance = 4n + 4*C*K*256 (masks) + 4*C*est_len), tensor throughput as secondary.
Under torchrun every rank runs its own batches (weak scaling, no collective on the data path); the
timed loop is bracketed by barriers and the slowest rank's time counts."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import numpy as np, torch
import sepcore
from sepcore import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--sets", type=int, default=2)
args = ap.parse_args()
n, C, taps, filters, stride = 32000, 2, 16, 256, 8
K = (n - taps) // stride + 1
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
gen = torch.Generator(device=dev).manual_seed(1 + rank)
enc = 0.25 * torch.randn((taps, filters), device=dev, generator=gen)
dec = 0.06 * torch.randn((filters, taps), device=dev, generator=gen)
sets = [(0.1 * torch.randn((args.batch, n), device=dev, generator=gen),
         torch.rand((args.batch, C, K, filters), device=dev, generator=gen)) for _ in range(args.sets)]
for w, m in sets:
    est = sepcore.filterbank_separate(w, enc, dec, m, stride=stride)
torch.cuda.synchronize()
# spot check one utterance against the oracle
from oracle import signal_path as oracle
_, want = oracle.filterbank_separate(sets[-1][0][0].cpu().numpy(), enc.cpu().numpy(), dec.cpu().numpy(),
                                     sets[-1][1][0].cpu().numpy(), stride)
err = float(np.max(np.abs(est[0].cpu().numpy() - want)) / np.max(np.abs(want)))
assert err < 1e-4, err
_lib.profile_enable(True)
for s in range(args.steps):
    w, m = sets[s % args.sets]
    sepcore.filterbank_separate(w, enc, dec, m, stride=stride)
torch.cuda.synchronize()
ms, cnt = _lib.profile_collect()
_lib.profile_enable(False)
ms /= cnt
# whole loop between CUDA events (all ranks start together; the slowest rank counts)
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.all_reduce(torch.zeros(1, device=dev))
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for s in range(args.steps):
    w, m = sets[s % args.sets]
    sepcore.filterbank_separate(w, enc, dec, m, stride=stride)
stop.record()
torch.cuda.synchronize()
loop_ms = start.elapsed_time(stop) / args.steps
if world > 1:
    t = torch.tensor([loop_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loop_ms = float(t.item())
est_len = (K - 1) * stride + taps
bytes_utt = 4 * n + 4 * C * K * filters + 4 * C * est_len
flop_utt = 2 * K * taps * filters * (1 + C)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
ach = bytes_utt * args.batch / (ms * 1e-3) / 1e9
line = json.dumps({
    "metric": "audio-sec/sec conv filterbank encode->mask->decode (tcgen05)",
    "value": world * args.batch * 4.0 / (loop_ms * 1e-3), "unit": "audio-s/s", "n_gpus": world, "scaling": "weak",
    "ms_per_step": loop_ms, "ms_per_launch": ms, "dtype": "tf32x3 (fp32-accurate), f32 accumulate",
    "config": {"workload": "cfg5: %d x 4 s @ 8 kHz, N=256, L=16, stride 8, C=2, fp32 masks (%.0f MB per set, %d sets)"
                           % (args.batch, bytes_utt * args.batch / 1e6, args.sets)},
    "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                 "useful_tflops": flop_utt * args.batch / (ms * 1e-3) / 1e12,
                 "issued_tf32_tflops": 3 * flop_utt * args.batch / (ms * 1e-3) / 1e12},
    "check": {"max_rel_err_vs_oracle": err}})
if rank == 0:
    print(line)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
