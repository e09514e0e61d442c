#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the fused separation hot path on B200.

Workload (BASELINE.json configs[1]): uPIT 2-speaker step = STFT -> mask -> iSTFT
overlap-add -> PSA labels + PIT-MSE -> SI-SDR/SDR on a synthetic batch of 64 x
4 s 8 kHz mixtures (N = 32000, C = 2, fp32), reference-default Blackman 256/128
(--size/--shift/--window select the other parameterisations).
A "step" is one fused pass over one batch.  One process per GPU; every rank runs
its own batches (weak scaling, utterance sharding); the path's only collective
is the all-reduce of the per-batch [loss, SI-SDR, SDR, n] sums.

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each number is made.

  value      device-resident throughput: K steps replayed from a CUDA graph,
             CUDA events on the launching stream, max over ranks
  roofline   dominant kernel's algorithmic bytes / its own event-bracketed
             duration (sep_profile_*), against MEASURED_PEAKS.json hbm_gbs
  e2e        same metric through the public host-pointer API: pinned host
             buffers, H2D + kernels + D2H inside the timed region
  cpu_baseline / --impl reference
             the oracle port (numpy restatement of the reference's algorithm;
             the Python reference itself cannot travel to the GPU box) on all
             host cores via multiprocessing
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "speech-separation-project-with-ai_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

SAMPLE_RATE = 8000
N_SETS = 6                      # distinct buffer sets rotated through (6 x 57.5 MB > 126 MB L2)
METRIC = "audio-sec/sec STFT->mask->iSTFT->PIT (fused signal path)"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="sepcore", choices=["sepcore", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--sources", type=int, default=2)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--shift", type=int, default=128)
    ap.add_argument("--window", default="blackman", choices=["blackman", "hann", "hamming"])
    ap.add_argument("--streams", type=int, default=6,
                    help="independent steps captured round-robin on this many streams inside the graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def window_fn(name):
    from scipy.signal import windows

    return {"blackman": windows.blackman, "hann": windows.hann, "hamming": windows.hamming}[name]


def geometry(args):
    n = int(round(args.seconds * SAMPLE_RATE))
    pad = args.size - args.shift
    frames = int(math.ceil((n + 2 * pad - args.size + args.shift) / args.shift))
    bins = args.size // 2 + 1
    c = args.sources
    # algorithmic bytes per utterance: read mix 4N + masks
    # 4CTF + refs 4CN, write estimates 4CN
    bytes_per_utt = 4 * n + 4 * c * frames * bins + 4 * c * n + 4 * c * n
    return n, frames, bins, bytes_per_utt


def workload_name(args):
    return ("cfg2: uPIT %d-spk STFT->mask->iSTFT->PIT-MSE+SI-SDR, %d x %g s @ 8 kHz, %s %d/%d, fp32"
            % (args.sources, args.batch, args.seconds, args.window, args.size, args.shift))


def make_set(args, seed):
    """Synthetic wsj0-2mix-shaped batch: refs 0.1*N(0,1), mix = sum, masks U[0,1)."""
    n, frames, bins, _ = geometry(args)
    rng = np.random.default_rng(seed)
    refs = (0.1 * rng.standard_normal((args.batch, args.sources, n), dtype=np.float32))
    mix = refs.sum(axis=1, dtype=np.float32)
    masks = rng.random((args.batch, args.sources, frames, bins), dtype=np.float32)
    return {"mix": mix, "refs": refs, "masks": masks}


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(job):
    mix, refs, masks, size, shift, wname = job
    from oracle import signal_path as oracle

    res = oracle.separate_and_score(mix, refs, masks, size=size, shift=shift, window=window_fn(wname))
    # the reference chain ends with permute_si_sdr on the waveforms
    if refs.shape[0] == 2:
        e = res["ests"][:, :mix.shape[0]].astype(np.float32)
        oracle.permute_si_sdr(refs[0], refs[1], e[0], e[1])
    return float(res["pit"]["loss"])


def _pool(cores):
    import multiprocessing as mp

    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    return mp.get_context("fork").Pool(cores)


def cpu_throughput(args, steps, warmup, budget_s):
    """Oracle port over all host cores.  Each step processes `sample` utterances of
    the workload (bounded so that the whole run stays within `budget_s`)."""
    cores = os.cpu_count() or 1
    data = make_set(args, seed=1)
    jobs = [(data["mix"][b], data["refs"][b], data["masks"][b], args.size, args.shift, args.window)
            for b in range(args.batch)]
    t0 = time.perf_counter()
    _cpu_worker(jobs[0])
    t_utt = time.perf_counter() - t0
    per_step = budget_s / max(1, steps + warmup)
    sample = int(per_step / t_utt * cores * 0.7)
    sample = max(min(cores, args.batch), min(args.batch, sample))
    with _pool(cores) as pool:
        chunk = max(1, sample // (cores * 2))
        for _ in range(warmup):
            pool.map(_cpu_worker, jobs[:sample], chunksize=chunk)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker, jobs[:sample], chunksize=chunk)
        dt = time.perf_counter() - t0
    audio = steps * sample * args.seconds
    return {"value": audio / dt, "cores": cores, "sample_utts": sample, "steps": steps,
            "ms_per_step": 1e3 * dt / steps, "t_utt_1core_ms": 1e3 * t_utt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_throughput(args, args.steps, args.warmup, budget_s=150.0)
    sample = ("%d of the %d utterances of the batch per step, %d steps, multiprocessing.Pool(%d), "
              "1 BLAS thread per worker" % (r["sample_utts"], args.batch, r["steps"], r["cores"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "oracle port of the reference's numpy path "
                   "(the Python reference cannot travel to the GPU box); host cores only"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU works."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reason_bits, self.stop_flag, self.ok = [], 0, threading.Event(), False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                util = self.nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                if util > 0:
                    self.reason_bits |= bits
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        reasons = [name for bit, name in self.REASONS.items()
                   if self.reason_bits & bit and name not in ("gpu_idle",)]
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
def run_sepcore(args):
    import torch
    import torch.distributed as dist

    import sepcore
    from sepcore import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, frames, bins, bytes_per_utt = geometry(args)
    win = window_fn(args.window)
    kw = dict(size=args.size, shift=args.shift, window=win)

    host_sets = [make_set(args, seed=1000 * rank + i) for i in range(N_SETS)]
    dev_sets = [{k: torch.from_numpy(v).to(dev) for k, v in s.items()} for s in host_sets]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput: CUDA-graph replay, K steps exactly ----
    block = min(args.steps, 1024)
    n_blocks, tail = divmod(args.steps, block)
    graph = sepcore.GraphedSeparator(dev_sets, block, streams=args.streams, **kw)
    tail_graph = sepcore.GraphedSeparator(dev_sets, tail, streams=args.streams, **kw) if tail else None
    warm = sepcore.GraphedSeparator(dev_sets, max(args.warmup, 3), streams=args.streams, **kw)
    warm.replay()
    if world > 1:
        dist.all_reduce(warm.sums)           # warms NCCL up too
    sampler = ClockSampler(local)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        # device-side rendezvous right in front of the start event: the ranks leave the host barrier at slightly
        # different times, and the all-reduce inside the timed region would charge that skew to every rank
        dist.all_reduce(torch.zeros(1, device=dev))
    start.record()
    for _ in range(n_blocks):
        graph.replay()
        if world > 1:
            dist.all_reduce(graph.sums)      # per-batch sums, one bucket per replay
    if tail_graph is not None:
        tail_graph.replay()
        if world > 1:
            dist.all_reduce(tail_graph.sums)
    stop.record()
    torch.cuda.synchronize()
    ms_total = start.elapsed_time(stop)
    barrier()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps * args.batch * args.seconds / (ms_total * 1e-3)
    sums = graph.sums[0].cpu().numpy().tolist()

    # ---- roofline leg: the dominant kernel alone, bracketed by events in the library ----
    prof_steps = min(args.steps, 200)
    ws = torch.zeros(sepcore.workspace_bytes(args.batch, args.sources, n, args.size, args.shift, win),
                     dtype=torch.uint8, device=dev)
    outs = {"est": torch.empty((args.batch, args.sources, n), device=dev),
            "scores": torch.empty((args.batch, sepcore.score_layout(args.sources)["stride"]),
                                  dtype=torch.float64, device=dev),
            "sums": torch.empty(4, dtype=torch.float64, device=dev)}
    l0 = sepcore.launch_count()
    sepcore.separate_and_score(dev_sets[0]["mix"], dev_sets[0]["masks"], dev_sets[0]["refs"], out=outs,
                               workspace=ws, **kw)
    launches_per_step = sepcore.launch_count() - l0
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for s in range(prof_steps):
        d = dev_sets[s % N_SETS]
        sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], out=outs, workspace=ws, **kw)
    torch.cuda.synchronize()
    kernel_ms, bracketed = _lib.profile_collect()
    _lib.profile_enable(False)
    kernel_ms_avg = kernel_ms / max(bracketed, 1)
    bytes_per_launch = bytes_per_utt * args.batch
    achieved = bytes_per_launch / (kernel_ms_avg * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("%d_%d_c%d_b%d" % (args.size, args.shift, args.sources, args.batch))

    # ---- end to end: public host-pointer API, pinned buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pin = [{k: torch.from_numpy(v).pin_memory() for k, v in s.items()} for s in host_sets]
        stride = sepcore.score_layout(args.sources)["stride"]
        pin_out = {"est": torch.empty((args.batch, args.sources, n)).pin_memory(),
                   "scores": torch.empty((args.batch, stride), dtype=torch.float64).pin_memory(),
                   "sums": torch.empty(4, dtype=torch.float64).pin_memory()}
        np_out = {k: v.numpy() for k, v in pin_out.items()}
        np_in = [{k: v.numpy() for k, v in s.items()} for s in pin]
        e2e_steps = max(3, min(args.steps, 100))
        pipe = sepcore.HostPipeline(args.batch, args.sources, n, depth=3, **kw)
        for s in range(3):
            pipe.submit(pin[s]["mix"], pin[s]["masks"], pin[s]["refs"])
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        loss, tickets = 0.0, []
        for s in range(e2e_steps):
            d = pin[s % N_SETS]
            tickets.append(pipe.submit(d["mix"], d["masks"], d["refs"]))   # H2D + kernels + D2H of step s
            if s >= 2:
                loss += float(pipe.result(tickets[s - 2])["sums"][0])       # step s-2's loss, read on the host
        for t in tickets[max(e2e_steps - 2, 0):]:
            loss += float(pipe.result(t)["sums"][0])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        # the synchronous drop-in call on the same buffers, for reference
        sync_steps = max(3, min(args.steps, 30))
        for s in range(3):
            sepcore.separate_and_score(np_in[s]["mix"], np_in[s]["masks"], np_in[s]["refs"], out=np_out, **kw)
        t0 = time.perf_counter()
        for s in range(sync_steps):
            d = np_in[s % N_SETS]
            res = sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], out=np_out, **kw)
            _ = float(res["sums"][0])
        dt_sync = (time.perf_counter() - t0) / sync_steps
        # the same pipeline with the waveforms in their on-disk format (int16 PCM in, audiowrite's int16 out)
        pin16 = [{"mix": torch.from_numpy(np.round(np.clip(np_in[i]["mix"], -1, 1) * 32767).astype(np.int16)).pin_memory(),
                  "refs": torch.from_numpy(np.round(np.clip(np_in[i]["refs"], -1, 1) * 32767).astype(np.int16)).pin_memory(),
                  "masks": pin[i]["masks"]} for i in range(N_SETS)]
        pipe16 = sepcore.HostPipeline(args.batch, args.sources, n, depth=3, pcm16=True, **kw)
        for s in range(3):
            pipe16.submit(pin16[s]["mix"], pin16[s]["masks"], pin16[s]["refs"])
        pipe16.drain()
        barrier()
        t0 = time.perf_counter()
        loss16, tickets = 0.0, []
        for s in range(e2e_steps):
            d = pin16[s % N_SETS]
            tickets.append(pipe16.submit(d["mix"], d["masks"], d["refs"]))
            if s >= 2:
                loss16 += float(pipe16.result(tickets[s - 2])["sums"][0])
        for t in tickets[max(e2e_steps - 2, 0):]:
            loss16 += float(pipe16.result(t)["sums"][0])
        torch.cuda.synchronize()
        dt16 = time.perf_counter() - t0
        t = torch.tensor([dt16], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt16 = float(t.item())
        e2e = {"value": world * e2e_steps * args.batch * args.seconds / dt, "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "sepcore.HostPipeline.submit/result on pinned host buffers (3 slots; copy-in, "
                      "compute, copy-out streams) -> sep_fused_separate_ws_f32",
               "sync_call_ms_per_step": 1e3 * dt_sync,
               "sync_call": "sepcore.separate_and_score(numpy views of pinned memory), SEP_MEM_HOST",
               "loss_sum": loss,
               "pcm16": {"value": world * e2e_steps * args.batch * args.seconds / dt16, "unit": UNIT,
                         "h2d_bytes_per_step": pipe16.h2d_bytes, "d2h_bytes_per_step": pipe16.d2h_bytes,
                         "ms_per_step": 1e3 * dt16 / e2e_steps, "loss_sum": loss16,
                         "api": "sepcore.HostPipeline(pcm16=True): int16 PCM mixture / references in (decoded on the "
                                "device like wavread), audiowrite's peak-normalised int16 estimates out"}}
    clocks = sampler.summary()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_throughput(args, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "%d utterances of the cfg batch x 3 steps, oracle port, multiprocessing.Pool(%d)"
                         % (r["sample_utts"], r["cores"])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "batch_per_gpu": args.batch,
                       "samples_per_utt": n, "frames": frames, "bins": bins, "sources": args.sources,
                       "l2": "rotating %d distinct buffer sets (%.0f MB > 126 MB L2)"
                             % (N_SETS, N_SETS * bytes_per_launch / 1e6),
                       "launch": "CUDA graph replay (%d-step graphs, independent steps round-robin on %d "
                                 "streams inside the graph)" % (block, graph.n_streams),
                       "parallelism": "utterance-sharded x%d, all-reduce of per-batch sums bucketed per replay"
                                      % world},
            # The timed region is the graph-replayed loop: one launch of the dominant kernel per step, steps of
            # different streams overlapping (its CTAs move in as the previous launch's retire), the two small
            # finalisation kernels hidden underneath.  achieved = algorithmic bytes per launch / (timed region /
            # launches).  "alone" is the same kernel launched eagerly, one at a time, events around the launch.
            "roofline": {"bound": "hbm", "achieved": bytes_per_launch / (ms_total / args.steps * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s",
                         "frac": bytes_per_launch / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "wstrip256_kernel (256/{128,64}, C=2) / strip256_kernel (C=1) / strip512_kernel (512/128) / "
                                   "tile or generic kernel otherwise; CUDA events around the timed region on the "
                                   "launching stream",
                         "kernel_ms": ms_total / args.steps, "bytes_per_launch": bytes_per_launch,
                         "launches_timed": args.steps,
                         "alone": {"kernel_ms": kernel_ms_avg, "achieved": achieved, "frac": achieved / peak,
                                   "launches_timed": bracketed,
                                   "how": "eager launches, one at a time, CUDA events around each launch on its stream "
                                          "(includes ~4 us of launch / event overhead; ncu: see profiles/)"}},
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps * world),
            "clocks": clocks, "check": {"pit_loss_sum": sums[0], "si_sdr_sum": sums[1], "n": sums[3]},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_sepcore(args)


if __name__ == "__main__":
    main()
