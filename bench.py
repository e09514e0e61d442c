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

  value      device-resident throughput: a CUDA graph of --steps fused steps,
             replayed `timing.replays` times so that the timed region is >= 20 ms
             whatever --steps is; CUDA events on the launching stream, max over ranks
  roofline   algorithmic bytes per launch / time per launch, against
             MEASURED_PEAKS.json hbm_gbs; `ms_per_step_overlapped` is the loop figure
             (launches of several streams share the SMs), `alone` the same kernel
             launched one at a time between events
  e2e        same metric through the public host-pointer API: pinned host
             buffers, H2D + kernels + D2H inside the timed region (+ int16 PCM and
             device-resident-mask variants)
  extra      the other BASELINE configs, each with its own clocks / roofline:
             cfg2 at Hann 256/64, cfg3 scoring, cfg4 (3 spk, 512/128), cfg5 tcgen05
             filterbank, a14 Conv1D at the reference shape, per-batch all-reduce (N > 1)
  cpu_baseline / --impl reference
             the UNMODIFIED reference functions (oracle/_ref: byte copies of
             parallel_stft.py, evaluate_metrics.py, notebook cells 38-39) on the host
             cores via multiprocessing (kind "reference"; pit_loss is TensorFlow code and
             is restated) -- or the numpy port when oracle/_ref is absent (kind "port")
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
METRIC = "audio-sec/sec STFT->mask->iSTFT->PIT (fused signal path)"
UNIT = "audio-s/s"
MIN_TIMED_MS = 20.0             # the graph is replayed until the timed region is at least this long
L2_BYTES = 126e6


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
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


def window_fn(name):
    from scipy.signal import windows

    return {"blackman": windows.blackman, "hann": windows.hann, "hamming": windows.hamming}[name]


class Workload:
    """One fused-path configuration (cfg2 / cfg2-Hann / cfg4)."""

    def __init__(self, name, batch, seconds, sources, size, shift, window):
        self.name, self.batch, self.seconds, self.sources = name, batch, seconds, sources
        self.size, self.shift, self.window = size, shift, window
        self.n = int(round(seconds * SAMPLE_RATE))
        pad = size - shift
        self.frames = int(math.ceil((self.n + 2 * pad - size + shift) / shift))
        self.bins = size // 2 + 1
        c = sources
        # algorithmic bytes per utterance: read mix 4N + masks 4CTF + refs 4CN,
        # write estimates 4CN
        self.bytes_per_utt = 4 * self.n + 4 * c * self.frames * self.bins + 4 * c * self.n + 4 * c * self.n
        self.bytes_per_launch = self.bytes_per_utt * batch
        self.n_sets = 6                         # 6 x 57.5 MB (cfg2) > 126 MB L2; 1, 2, 3 or 6 graph lanes divide it

    def kw(self):
        return dict(size=self.size, shift=self.shift, window=window_fn(self.window))

    def describe(self):
        return ("%s: uPIT %d-spk STFT->mask->iSTFT->PIT-MSE+SI-SDR, %d x %g s @ 8 kHz, %s %d/%d, fp32"
                % (self.name, self.sources, self.batch, self.seconds, self.window, self.size, self.shift))

    def config(self):
        """The workload description -- identical in the sepcore and the reference arm."""
        return {"workload": self.describe(), "batch_per_gpu": self.batch, "samples_per_utt": self.n,
                "frames": self.frames, "bins": self.bins, "sources": self.sources, "size": self.size,
                "shift": self.shift, "window": self.window,
                "l2": "inputs larger than L2: %d distinct buffer sets rotated (%.0f MB > 126 MB L2)"
                      % (self.n_sets, self.n_sets * self.bytes_per_launch / 1e6),
                "replays": "the timed region holds timing.replays x --steps steps (>= %g ms), captured back to back in "
                           "CUDA graphs of up to 1024 steps; ms_per_step = timed region / (replays * steps)" % MIN_TIMED_MS}

    def make_set(self, seed):
        """Synthetic wsj0-2mix-shaped batch: refs 0.1*N(0,1), mix = sum, masks U[0,1)."""
        rng = np.random.default_rng(seed)
        refs = (0.1 * rng.standard_normal((self.batch, self.sources, self.n), dtype=np.float32))
        mix = refs.sum(axis=1, dtype=np.float32)
        masks = rng.random((self.batch, self.sources, self.frames, self.bins), dtype=np.float32)
        return {"mix": mix, "refs": refs, "masks": masks}


def main_workload(args):
    return Workload("cfg2", args.batch, args.seconds, args.sources, args.size, args.shift, args.window)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(key):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        return json.load(open(path)).get(key)
    return None


# ----------------------------------------------------------------------------- CPU arm
_REF = {}


def _reference_ns():
    """The unmodified reference callables (oracle/_ref on the GPU box), or None."""
    if "ns" not in _REF:
        try:
            from oracle import reference_loader

            _REF["ns"] = reference_loader.load() if reference_loader.available() else None
        except Exception:
            _REF["ns"] = None
    return _REF["ns"]


def _cpu_worker(job):
    """The reference chain for one utterance: stft x (1 + C) -> |X|, angle, PSA labels ->
    mask * |X| -> e^{j angle} recombination -> istft x C -> pit_loss -> permute_si_sdr."""
    mix, refs, masks, size, shift, wname = job
    from oracle import signal_path as oracle

    ns = _reference_ns()
    win = window_fn(wname)
    if ns is None:
        res = oracle.separate_and_score(mix, refs, masks, size=size, shift=shift, window=win)
        if refs.shape[0] == 2:
            e = res["ests"][:, :mix.shape[0]].astype(np.float32)
            oracle.permute_si_sdr(refs[0], refs[1], e[0], e[1])
        return float(res["pit"]["loss"])
    n_src, n = refs.shape[0], mix.shape[0]
    spec = ns.stft(mix, time_dim=0, size=size, shift=shift, window=win)               # parallel_stft.py:146
    mag, phase = np.abs(spec), np.angle(spec)                                          # :262-263
    labels = []
    for c in range(n_src):
        s = ns.stft(refs[c], time_dim=0, size=size, shift=shift, window=win)
        labels.append(np.abs(s) * np.cos(phase - np.angle(s)))                         # :270-272
    labels = np.concatenate(labels, axis=1)
    cleaned = np.concatenate([masks[c] * mag for c in range(n_src)], axis=1)           # cell 29
    ests = []
    for c in range(n_src):
        sc = cleaned[:, c * mag.shape[1]:(c + 1) * mag.shape[1]] * np.exp(phase * 1j)  # cell 41 :1385-1388
        ests.append(ns.istft(sc, size=size, shift=shift, window=win)[:n])              # cell 39
    y_true = np.concatenate([labels, np.full((1, labels.shape[1]), float(mag.shape[0]))], axis=0)
    pit = oracle.pit_mse(y_true[None], cleaned[None], mag.shape[1])                    # cell 28 (TensorFlow) restated
    if n_src == 2:
        e = np.asarray(ests, dtype=np.float32)
        ns.permute_si_sdr(refs[0], refs[1], e[0], e[1])                                 # evaluate_metrics.py:28
    return float(pit["loss"])


def _pool(cores):
    import multiprocessing as mp

    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[var] = "1"
    return mp.get_context("fork").Pool(cores)


def cpu_throughput(wl, steps, warmup, budget_s):
    """The reference chain over all host cores.  Each step processes `sample` utterances of the
    workload (bounded so that the whole run stays within `budget_s`); also the 1-thread figure
    (how the reference actually runs: a Python loop over files)."""
    cores = os.cpu_count() or 1
    kind = "reference" if _reference_ns() is not None else "port"
    data = wl.make_set(seed=1)
    jobs = [(data["mix"][b], data["refs"][b], data["masks"][b], wl.size, wl.shift, wl.window)
            for b in range(wl.batch)]
    _cpu_worker(jobs[0])                                   # imports, window tables
    t0 = time.perf_counter()
    n1 = 0
    while n1 < min(4, wl.batch) and (n1 == 0 or time.perf_counter() - t0 < 3.0):
        _cpu_worker(jobs[n1])
        n1 += 1
    t_utt = (time.perf_counter() - t0) / n1
    per_step = budget_s / max(1, steps + warmup)
    sample = int(per_step / t_utt * cores * 0.7)
    sample = max(min(cores, wl.batch), min(wl.batch, sample))
    with _pool(cores) as pool:
        chunk = max(1, sample // (cores * 2))
        for _ in range(warmup):
            pool.map(_cpu_worker, jobs[:sample], chunksize=chunk)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker, jobs[:sample], chunksize=chunk)
        dt = time.perf_counter() - t0
    audio = steps * sample * wl.seconds
    return {"value": audio / dt, "cores": cores, "kind": kind, "sample_utts": sample, "steps": steps,
            "ms_per_step": 1e3 * dt / steps, "single_thread_value": wl.seconds / t_utt,
            "single_thread_ms_per_utt": 1e3 * t_utt, "single_thread_utts": n1}


def cpu_baseline_dict(r, wl):
    what = ("unmodified reference functions from oracle/_ref (stft, istft, si_sdr, permute_si_sdr; pit_loss is "
            "TensorFlow code and runs as its numpy restatement)" if r["kind"] == "reference"
            else "numpy port of the reference path (oracle/_ref absent)")
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
            "sample": "%d of the %d utterances of the %s batch per step x %d steps, multiprocessing.Pool(%d), "
                      "1 BLAS thread per worker; %s" % (r["sample_utts"], wl.batch, wl.name, r["steps"], r["cores"], what),
            "single_thread": {"value": r["single_thread_value"], "unit": UNIT, "cores": 1,
                              "ms_per_utterance": r["single_thread_ms_per_utt"],
                              "sample": "%d utterances in a plain Python loop, 1 thread -- how the reference "
                                        "itself runs (BASELINE.md 3a)" % r["single_thread_utts"]}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = main_workload(args)
    r = cpu_throughput(wl, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.config(),
        "cpu_baseline": cpu_baseline_dict(r, wl),
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
            time.sleep(0.003)

    def summary(self):
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        reasons = [name for bit, name in self.REASONS.items()
                   if self.reason_bits & bit and name not in ("gpu_idle",)]
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


class Ctx:
    """Per-process CUDA / torch.distributed plumbing."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def rendezvous(self):
        """device-side line-up right in front of a start event: the ranks leave the host barrier at slightly
        different times, and a collective inside the timed region would charge that skew to every rank"""
        if self.world > 1:
            self.dist.all_reduce(self.torch.zeros(1, device=self.dev))

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_replays(ctx, replay_once, est_calls=1, min_ms=None, yield_every=0):
    """Warm, estimate one replay, then time R replays with R chosen so that the region is >= min_ms (default
    MIN_TIMED_MS; same R on every rank).  Returns (total_ms max over ranks, R, clocks).  `yield_every`: a
    Python loop of eager launches holds the GIL, and the clock sampler thread would starve: hand the GIL over
    every so many calls (the GPU stays ahead of the host by its launch queue)."""
    torch = ctx.torch
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    start.record()
    for _ in range(est_calls):
        replay_once()
    stop.record()
    torch.cuda.synchronize()
    est = ctx.max_over_ranks(start.elapsed_time(stop) / est_calls)
    reps = max(1, int(math.ceil((min_ms or MIN_TIMED_MS) / max(est, 1e-3))))
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ctx.barrier()
    ctx.rendezvous()
    start.record()
    for i in range(reps):
        replay_once()
        if yield_every and i % yield_every == yield_every - 1:
            time.sleep(0)
    stop.record()
    torch.cuda.synchronize()
    ms_total = start.elapsed_time(stop)
    clocks = sampler.summary()
    ctx.barrier()
    return ctx.max_over_ranks(ms_total), reps, clocks


# ----------------------------------------------------------------------------- fused path (cfg2 / cfg2-Hann / cfg4)
def measure_fused(ctx, wl, steps, warmup, streams, reduce_mode="bucket", alone_launches=200, dev_sets=None):
    """Device-resident throughput of the fused path on workload `wl`.

    reduce_mode (N > 1): "bucket" = one all-reduce of the [steps, 4] sums per replay; "per_batch" = one NCCL
    all-reduce per step, captured in the graph on a dedicated communication stream."""
    import sepcore
    from sepcore import _lib

    torch, dist = ctx.torch, ctx.dist
    kw = wl.kw()
    if dev_sets is None:
        dev_sets = [{k: torch.from_numpy(v).to(ctx.dev) for k, v in wl.make_set(1000 * ctx.rank + i).items()}
                    for i in range(wl.n_sets)]
    torch.cuda.synchronize()
    per_batch = reduce_mode in ("per_batch", "push") and ctx.world > 1
    red = sepcore.distributed.all_reduce_sums if per_batch and reduce_mode == "per_batch" else None
    warm = sepcore.GraphedSeparator(dev_sets, max(warmup, 3), streams=streams, **kw)
    warm.replay()
    if ctx.world > 1:
        dist.all_reduce(warm.sums)           # warms NCCL up too
    # The timed region must be >= MIN_TIMED_MS whatever --steps is (a 20-step graph is 0.5 ms: mostly ramp and
    # tail of 50 us launches).  Estimate the step time on the warm-up graph, then time `reps` x --steps steps,
    # captured back to back in graphs of up to 1024 steps (no join between the --steps groups).
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    start.record()
    for _ in range(4):
        warm.replay()
    stop.record()
    torch.cuda.synchronize()
    est_step = ctx.max_over_ranks(start.elapsed_time(stop) / (4 * warm.steps))
    # (the short warm-up graph over-estimates the step time -- ramp and tail -- hence the margin)
    reps = max(1, int(math.ceil(1.7 * MIN_TIMED_MS / max(est_step * steps, 1e-3))))
    total_steps = reps * steps
    block = min(total_steps, 1024)
    n_blocks, tail = divmod(total_steps, block)
    peer = sepcore.distributed.PeerSums(slots=block) if per_batch and reduce_mode == "push" else None
    graph = sepcore.GraphedSeparator(dev_sets, block, streams=streams, reduce_each_step=red, push=peer, **kw)
    tail_graph = sepcore.GraphedSeparator(dev_sets, tail, streams=streams, reduce_each_step=red, push=peer, **kw) if tail else None
    graph.replay()
    torch.cuda.synchronize()
    if peer is not None:
        ctx.barrier()
        peer.arrived.zero_()
        ctx.barrier()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ctx.barrier()
    ctx.rendezvous()
    start.record()
    for _ in range(n_blocks):
        graph.replay()
        if ctx.world > 1 and not per_batch:
            dist.all_reduce(graph.sums)      # per-batch sums, one bucket per replay
    if tail_graph is not None:
        tail_graph.replay()
        if ctx.world > 1 and not per_batch:
            dist.all_reduce(tail_graph.sums)
    stop.record()
    torch.cuda.synchronize()
    ms_total = start.elapsed_time(stop)
    clocks = sampler.summary()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(ms_total)
    ms_step = ms_total / total_steps
    sums = graph.sums[0].cpu().numpy().tolist()
    push_check = None
    if peer is not None:
        # after the barrier every rank's pushes have landed: the inbox rows must add up to the NCCL all-reduce
        want = graph.sums.clone()
        dist.all_reduce(want)
        got = peer.reduced()[:block]
        push_check = {"arrived_min": int(peer.arrived[:block].min().item()), "arrived_expected": ctx.world * n_blocks,
                      "max_rel_diff_vs_nccl": float(((got - want).abs() / want.abs().clamp_min(1e-300)).max().item())}
        ctx.barrier()
        peer.close()

    # the dominant kernel alone: eager launches, one at a time, bracketed by events in the library
    n = wl.n
    ws = torch.zeros(sepcore.workspace_bytes(wl.batch, wl.sources, n, wl.size, wl.shift, kw["window"]),
                     dtype=torch.uint8, device=ctx.dev)
    outs = {"est": torch.empty((wl.batch, wl.sources, n), device=ctx.dev),
            "scores": torch.empty((wl.batch, sepcore.score_layout(wl.sources)["stride"]),
                                  dtype=torch.float64, device=ctx.dev),
            "sums": torch.empty(4, dtype=torch.float64, device=ctx.dev)}
    l0 = sepcore.launch_count()
    sepcore.separate_and_score(dev_sets[0]["mix"], dev_sets[0]["masks"], dev_sets[0]["refs"], out=outs,
                               workspace=ws, **kw)
    launches_per_step = sepcore.launch_count() - l0
    kernel = _lib.last_kernel()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for s in range(alone_launches):
        d = dev_sets[s % len(dev_sets)]
        sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], out=outs, workspace=ws, **kw)
    torch.cuda.synchronize()
    k_ms, bracketed = _lib.profile_collect()
    _lib.profile_enable(False)
    k_ms /= max(bracketed, 1)
    peak, peak_src = hbm_peak()
    ach = wl.bytes_per_launch / (ms_step * 1e-3) / 1e9
    ach_alone = wl.bytes_per_launch / (k_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": traffic_for("%d_%d_c%d_b%d" % (wl.size, wl.shift, wl.sources, wl.batch)),
        "peak_source": peak_src, "kernel": kernel, "bytes_per_launch": wl.bytes_per_launch,
        "launches_timed": total_steps,
        # achieved = algorithmic bytes per launch / (timed region / launches): the loop figure.  It is a
        # throughput inverse, NOT a launch duration: the graph keeps several launches in flight (one per stream)
        # and 2-3 of them are co-resident on an SM, which is how a stream of independent batches is meant to run.
        "ms_per_step_overlapped": ms_step,
        "kernel_ms_alone": k_ms,
        "alone": {"kernel_ms": k_ms, "achieved": ach_alone, "frac": ach_alone / peak, "launches_timed": bracketed,
                  "how": "the same kernel launched eagerly, one at a time, CUDA events around each launch on its "
                         "stream (includes ~4 us of launch / event overhead; ncu durations: profiles/)"},
    }
    return {"ms_total": ms_total, "replays": reps, "ms_per_step": ms_step, "clocks": clocks, "sums": sums,
            "launches_per_step": launches_per_step, "roofline": roofline, "graph_steps": block,
            "graph_replays": n_blocks + (1 if tail else 0), "push_check": push_check,
            "n_streams": graph.n_streams, "dev_sets": dev_sets,
            "value": ctx.world * wl.batch * wl.seconds / (ms_step * 1e-3)}


def fused_extra(ctx, name, wl, steps, streams):
    r = measure_fused(ctx, wl, steps, 3, streams, alone_launches=40)
    cfg = wl.config()
    return {"name": name, "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": ctx.world,
            "ms_per_step": r["ms_per_step"], "dtype": "f32", "scaling": "weak", "config": cfg,
            "timing": {"replays": r["replays"], "steps_timed": r["replays"] * steps, "graph_steps": r["graph_steps"],
                       "graph_replays": r["graph_replays"], "timed_ms": r["ms_total"], "streams": r["n_streams"]},
            "roofline": r["roofline"], "clocks": r["clocks"], "gpu_launches": r["launches_per_step"] * r["replays"] * steps,
            "check": {"pit_loss_sum": r["sums"][0], "si_sdr_sum": r["sums"][1], "n": r["sums"][3]}}


# ----------------------------------------------------------------------------- cfg3: scoring
def scoring_extra(ctx, utts=3000):
    """BASELINE config 3: SI-SDR/SDR over the synthetic 3000-utterance wsj0-2mix-shaped set, sharded by
    load over the ranks (strong scaling), one all-reduce of the dataset sums per step."""
    import sepcore
    from sepcore import _lib
    from sepcore.distributed import shard_by_load

    torch, dist = ctx.torch, ctx.dist
    rng = np.random.default_rng(4)
    all_lengths = (8000 * rng.uniform(2, 10, size=utts)).astype(np.int64)
    lengths = all_lengths if ctx.world == 1 else all_lengths[shard_by_load(all_lengths, ctx.world)[ctx.rank]]
    n_mine = len(lengths)
    padded = (lengths + 3) & ~3
    offs = np.zeros(2 * n_mine, dtype=np.int64)
    offs[1:] = np.cumsum(np.repeat(padded, 2))[:-1]
    total = int(np.repeat(padded, 2).sum())
    gen = torch.Generator(device=ctx.dev).manual_seed(5 + ctx.rank)
    refs = 0.1 * torch.randn(total, device=ctx.dev, generator=gen)
    snr = rng.uniform(-5, 20, size=n_mine)
    scale = torch.repeat_interleave(torch.from_numpy((10 ** (-snr / 20)).astype(np.float32)).to(ctx.dev),
                                    torch.from_numpy(2 * padded).to(ctx.dev))       # constant per utterance
    ests = refs + scale * (0.1 * torch.randn(total, device=ctx.dev, generator=gen))
    del scale
    est_offs = offs.copy()
    swap = np.arange(n_mine) % 2 == 1
    est_offs[0::2][swap], est_offs[1::2][swap] = offs[1::2][swap], offs[0::2][swap]
    torch.cuda.synchronize()
    res = sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
    kernel = _lib.last_kernel()
    torch.cuda.synchronize()
    from oracle import signal_path as oracle
    worst = 0.0
    for b in range(min(4, n_mine)):                      # parity spot check against the oracle
        n = int(lengths[b])
        r = [refs[offs[2 * b + c]:offs[2 * b + c] + n].cpu().numpy() for c in range(2)]
        e = [ests[est_offs[2 * b + c]:est_offs[2 * b + c] + n].cpu().numpy() for c in range(2)]
        v, p, _, _ = oracle.permute_si_sdr_detail(r[0], r[1], e[0], e[1])
        assert int(res["si_perm"][b].item()) == p, (b, p)
        worst = max(worst, abs(float(res["si_best"][b].item()) - float(v)))
    assert worst < 0.01, worst
    l0 = sepcore.launch_count()
    out = sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
    launches = sepcore.launch_count() - l0
    if ctx.world > 1:
        dist.all_reduce(out["sums"].clone())
    state = {"out": out}

    def once():
        state["out"] = sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
        if ctx.world > 1:
            dist.all_reduce(state["out"]["sums"])       # dataset sums over all shards (evaluate_metrics.py:53,90)

    ms_total, reps, clocks = timed_replays(ctx, once, est_calls=3)
    ms = ms_total / reps
    _lib.profile_enable(True)
    for _ in range(10):
        sepcore.score_flat_device(refs, ests, offs, est_offs, lengths, 2)
    torch.cuda.synchronize()
    k_ms, k_cnt = _lib.profile_collect()
    _lib.profile_enable(False)
    k_ms /= max(k_cnt, 1)
    audio_s = float(all_lengths.sum()) / SAMPLE_RATE
    bytes_alg = 16.0 * float(lengths.sum())
    peak, peak_src = hbm_peak()
    sums = state["out"]["sums"]
    return {
        "name": "cfg3_scoring", "metric": "audio-sec/sec SI-SDR+SDR scoring (evaluate_metrics)",
        "value": audio_s / (ms * 1e-3), "unit": UNIT, "n_gpus": ctx.world, "ms_per_step": ms, "scaling": "strong",
        "dtype": "f32 in, f64 accumulate",
        "config": {"workload": "cfg3: %d utterances, 2-10 s @ 8 kHz, 2 refs + 2 ests (%.2f GB > L2), sharded by load"
                               % (utts, 16.0 * float(all_lengths.sum()) / 1e9), "utterances_on_rank0": n_mine},
        "timing": {"replays": reps, "timed_ms": ms_total, "launches_per_step": launches},
        "roofline": {"bound": "hbm", "achieved": bytes_alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "kernel": kernel,
                     "kernel_ms": k_ms, "launches_timed": k_cnt, "bytes_per_launch": bytes_alg,
                     "traffic": traffic_for("cfg3_score_chunk"),
                     "whole_call": {"achieved": bytes_alg / (ms * 1e-3) / 1e9, "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak,
                                    "note": "adds finalize + sums kernels and the per-call metadata upload"}},
        "clocks": clocks, "gpu_launches": launches * reps,
        "check": {"oracle_utts": min(4, n_mine), "max_abs_db_err": worst,
                  "mean_si_sdr_db": float(sums[0].item() / sums[2].item())},
    }


def graphed_calls(ctx, once, calls):
    """`calls` consecutive invocations of `once` (rotating its buffer sets) captured in ONE CUDA graph: the timed loop of an
    eager Python call per launch is host-bound below ~150 us per call and starves the clock sampler of the GIL."""
    torch = ctx.torch
    side = torch.cuda.Stream(device=ctx.dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            once()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(calls):
            once()
    torch.cuda.synchronize()
    return graph.replay


# ----------------------------------------------------------------------------- cfg5: tcgen05 filterbank
def filterbank_extra(ctx, batch=64, sets=2):
    import sepcore
    from sepcore import _lib

    torch = ctx.torch
    n, C, taps, filters, stride = 32000, 2, 16, 256, 8
    K = (n - taps) // stride + 1
    gen = torch.Generator(device=ctx.dev).manual_seed(1 + ctx.rank)
    enc = 0.25 * torch.randn((taps, filters), device=ctx.dev, generator=gen)
    dec = 0.06 * torch.randn((filters, taps), device=ctx.dev, generator=gen)
    data = [(0.1 * torch.randn((batch, n), device=ctx.dev, generator=gen),
             torch.rand((batch, C, K, filters), device=ctx.dev, generator=gen)) for _ in range(sets)]
    for w, m in data:
        est = sepcore.filterbank_separate(w, enc, dec, m, stride=stride)
    kernel = _lib.last_kernel()
    torch.cuda.synchronize()
    from oracle import signal_path as oracle
    _, want = oracle.filterbank_separate(data[-1][0][0].cpu().numpy(), enc.cpu().numpy(), dec.cpu().numpy(),
                                         data[-1][1][0].cpu().numpy(), stride)
    err = float(np.max(np.abs(est[0].cpu().numpy() - want)) / np.max(np.abs(want)))
    assert err < 1e-4, err
    state = {"s": 0}

    def once():
        w, m = data[state["s"] % sets]
        state["s"] += 1
        sepcore.filterbank_separate(w, enc, dec, m, stride=stride)

    calls = 4 * sets
    ms_total, reps, clocks = timed_replays(ctx, graphed_calls(ctx, once, calls), est_calls=2)
    loop_ms = ms_total / (reps * calls)
    _lib.profile_enable(True)
    for _ in range(10):
        once()
    torch.cuda.synchronize()
    alone_ms, k_cnt = _lib.profile_collect()
    _lib.profile_enable(False)
    alone_ms /= max(k_cnt, 1)
    k_ms, k_cnt = loop_ms, reps * calls      # the kernel's average duration over the timed region (launches back to back)
    est_len = (K - 1) * stride + taps
    bytes_utt = 4 * n + 4 * C * K * filters + 4 * C * est_len
    flop_utt = 2 * K * taps * filters * (1 + C)
    peak, peak_src = hbm_peak()
    ach = bytes_utt * batch / (k_ms * 1e-3) / 1e9
    return {
        "name": "cfg5_filterbank", "metric": "audio-sec/sec conv filterbank encode->mask->decode (tcgen05)",
        "value": ctx.world * batch * 4.0 / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": ctx.world, "scaling": "weak",
        "ms_per_step": loop_ms, "dtype": "tf32x3 (fp32-accurate), f32 accumulate",
        "config": {"workload": "cfg5: %d x 4 s @ 8 kHz, N=256, L=16, stride 8, C=2, fp32 masks (%.0f MB per set, %d sets > L2)"
                               % (batch, bytes_utt * batch / 1e6, sets)},
        "timing": {"replays": reps, "calls_per_graph": calls, "timed_ms": ms_total},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": peak_src, "kernel": kernel, "kernel_ms": k_ms, "launches_timed": k_cnt,
                     "kernel_ms_alone": alone_ms,
                     "bytes_per_launch": bytes_utt * batch, "traffic": traffic_for("cfg5_filterbank_b64"),
                     "useful_tflops": flop_utt * batch / (k_ms * 1e-3) / 1e12,
                     "issued_tf32_tflops": 3 * flop_utt * batch / (k_ms * 1e-3) / 1e12},
        "clocks": clocks, "gpu_launches": reps * calls, "check": {"max_rel_err_vs_oracle": err},
    }


# ----------------------------------------------------------------------------- a14: Conv1D at the reference shape
def conv1d_extra(ctx, batch=64, sets=4):
    """Raw_with_Convlayer.ipynb:389: Conv1D(129, 2, sigmoid, 'same') on [B, K=800, 40] (4 s of 8 kHz audio in 40-sample
    rows).  Algorithmic bytes: read x 4*K*40 + write 4*K*129 per utterance (weights 41 KB, L2-resident)."""
    import sepcore
    from sepcore import _lib

    torch = ctx.torch
    K, c_in, filters = 800, 40, 129
    reps_b = 16                                            # utterances per call = batch * reps_b (a dataset shard)
    gen = torch.Generator(device=ctx.dev).manual_seed(3 + ctx.rank)
    w = 0.05 * torch.randn((2, c_in, filters), device=ctx.dev, generator=gen)
    b = 0.05 * torch.randn((filters,), device=ctx.dev, generator=gen)
    data = [0.1 * torch.randn((batch * reps_b, K, c_in), device=ctx.dev, generator=gen) for _ in range(sets)]
    out = sepcore.conv1d(data[0], w, b, padding="same", activation="sigmoid")
    kernel = _lib.last_kernel()
    torch.cuda.synchronize()
    from oracle import signal_path as oracle
    want = oracle.conv1d(data[0][0].cpu().numpy()[None], w.cpu().numpy(), b.cpu().numpy(), padding="same",
                         activation="sigmoid")
    err = float(np.max(np.abs(out[0].cpu().numpy() - want[0])))
    assert err < 1e-5, err
    state = {"s": 0}

    def once():
        sepcore.conv1d(data[state["s"] % sets], w, b, padding="same", activation="sigmoid")
        state["s"] += 1

    calls = 2 * sets
    ms_total, reps, clocks = timed_replays(ctx, graphed_calls(ctx, once, calls), est_calls=2)
    loop_ms = ms_total / (reps * calls)
    _lib.profile_enable(True)
    for _ in range(10):
        once()
    torch.cuda.synchronize()
    alone_ms, k_cnt = _lib.profile_collect()
    _lib.profile_enable(False)
    alone_ms /= max(k_cnt, 1)
    k_ms, k_cnt = loop_ms, reps * calls      # the kernel's average duration over the timed region (launches back to back)
    utts = batch * reps_b
    bytes_launch = utts * (4 * K * c_in + 4 * K * filters)
    peak, peak_src = hbm_peak()
    ach = bytes_launch / (k_ms * 1e-3) / 1e9
    return {
        "name": "a14_conv1d_reference_shape", "metric": "audio-sec/sec Conv1D(129, 2, sigmoid, 'same') encoder",
        "value": ctx.world * utts * 4.0 / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": ctx.world, "scaling": "weak",
        "ms_per_step": loop_ms, "dtype": "f32",
        "config": {"workload": "a14: x [%d, 800, 40] -> [%d, 800, 129], W [2, 40, 129], sigmoid (%.0f MB per set, %d sets > L2)"
                               % (utts, utts, bytes_launch / 1e6, sets)},
        "timing": {"replays": reps, "calls_per_graph": calls, "timed_ms": ms_total},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "peak_source": peak_src, "kernel": kernel, "kernel_ms": k_ms, "launches_timed": k_cnt,
                     "kernel_ms_alone": alone_ms,
                     "bytes_per_launch": bytes_launch, "traffic": traffic_for("a14_conv1d"),
                     "tflops": 2.0 * utts * K * 2 * c_in * filters / (k_ms * 1e-3) / 1e12},
        "clocks": clocks, "gpu_launches": reps * calls, "check": {"max_abs_err_vs_oracle": err},
    }


# ----------------------------------------------------------------------------- end to end
def measure_e2e(ctx, wl, steps):
    """The public host-pointer API on pinned host buffers; H2D + kernels + D2H every step inside the timed region."""
    import sepcore

    torch = ctx.torch
    kw = wl.kw()
    n = wl.n
    n_host = min(wl.n_sets, 6)
    host_sets = [wl.make_set(1000 * ctx.rank + i) for i in range(n_host)]
    pin = [{k: torch.from_numpy(v).pin_memory() for k, v in s.items()} for s in host_sets]
    stride = sepcore.score_layout(wl.sources)["stride"]
    pin_out = {"est": torch.empty((wl.batch, wl.sources, n)).pin_memory(),
               "scores": torch.empty((wl.batch, stride), dtype=torch.float64).pin_memory(),
               "sums": torch.empty(4, dtype=torch.float64).pin_memory()}
    np_out = {k: v.numpy() for k, v in pin_out.items()}
    np_in = [{k: v.numpy() for k, v in s.items()} for s in pin]
    e2e_steps = max(3, min(steps, 100))

    def run_pipe(pipe, feeds):
        for s in range(3):
            pipe.submit(*feeds[s % len(feeds)])
        pipe.drain()
        ctx.barrier()
        t0 = time.perf_counter()
        loss, tickets = 0.0, []
        for s in range(e2e_steps):
            tickets.append(pipe.submit(*feeds[s % len(feeds)]))             # H2D + kernels + D2H of step s
            if s >= 2:
                loss += float(pipe.result(tickets[s - 2])["sums"][0])       # step s-2's loss, read on the host
        for t in tickets[max(e2e_steps - 2, 0):]:
            loss += float(pipe.result(t)["sums"][0])
        torch.cuda.synchronize()
        return ctx.max_over_ranks(time.perf_counter() - t0), loss

    pipe = sepcore.HostPipeline(wl.batch, wl.sources, n, depth=3, **kw)
    dt, loss = run_pipe(pipe, [(d["mix"], d["masks"], d["refs"]) for d in pin])
    # the synchronous drop-in call on the same buffers, for reference
    sync_steps = max(3, min(steps, 30))
    for s in range(3):
        sepcore.separate_and_score(np_in[s % n_host]["mix"], np_in[s % n_host]["masks"], np_in[s % n_host]["refs"],
                                   out=np_out, **kw)
    t0 = time.perf_counter()
    for s in range(sync_steps):
        d = np_in[s % n_host]
        res = sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], out=np_out, **kw)
        _ = float(res["sums"][0])
    dt_sync = (time.perf_counter() - t0) / sync_steps
    # the same pipeline with the waveforms in their on-disk format (int16 PCM in, audiowrite's int16 out)
    to16 = lambda x: torch.from_numpy(np.round(np.clip(x, -1, 1) * 32767).astype(np.int16)).pin_memory()
    pin16 = [{"mix": to16(np_in[i]["mix"]), "refs": to16(np_in[i]["refs"]), "masks": pin[i]["masks"]}
             for i in range(n_host)]
    pipe16 = sepcore.HostPipeline(wl.batch, wl.sources, n, depth=3, pcm16=True, **kw)
    dt16, loss16 = run_pipe(pipe16, [(d["mix"], d["masks"], d["refs"]) for d in pin16])
    # ... and with the masks where the model leaves them: on the device (only waveforms cross PCIe)
    dev_masks = [pin[i]["masks"].to(ctx.dev) for i in range(n_host)]
    pipe_dm = sepcore.HostPipeline(wl.batch, wl.sources, n, depth=3, pcm16=True, device_masks=True, **kw)
    dt_dm, loss_dm = run_pipe(pipe_dm, [(pin16[i]["mix"], dev_masks[i], pin16[i]["refs"]) for i in range(n_host)])
    audio = ctx.world * e2e_steps * wl.batch * wl.seconds
    return {"value": audio / dt, "unit": UNIT,
            "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
            "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
            "api": "sepcore.HostPipeline.submit/result on pinned host buffers (3 slots; copy-in, "
                   "compute, copy-out streams) -> sep_fused_separate_ws_f32",
            "sync_call_ms_per_step": 1e3 * dt_sync,
            "sync_call": "sepcore.separate_and_score(numpy views of pinned memory), SEP_MEM_HOST",
            "loss_sum": loss,
            "pcm16": {"value": audio / dt16, "unit": UNIT,
                      "h2d_bytes_per_step": pipe16.h2d_bytes, "d2h_bytes_per_step": pipe16.d2h_bytes,
                      "ms_per_step": 1e3 * dt16 / e2e_steps, "loss_sum": loss16,
                      "api": "sepcore.HostPipeline(pcm16=True): int16 PCM mixture / references in (decoded on the "
                             "device like wavread), audiowrite's peak-normalised int16 estimates out"},
            "pcm16_device_masks": {"value": audio / dt_dm, "unit": UNIT,
                                   "h2d_bytes_per_step": pipe_dm.h2d_bytes, "d2h_bytes_per_step": pipe_dm.d2h_bytes,
                                   "ms_per_step": 1e3 * dt_dm / e2e_steps, "loss_sum": loss_dm,
                                   "api": "HostPipeline(pcm16=True, device_masks=True): the masks are the model's output "
                                          "and already live on the device (cell 29); only int16 waveforms cross PCIe"}}


# ----------------------------------------------------------------------------- GPU arm
def run_sepcore(args):
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    wl = main_workload(args)
    r = measure_fused(ctx, wl, args.steps, args.warmup, args.streams)
    extras = []
    if not args.no_extras:
        if ctx.world > 1:
            # the per-batch all-reduce the north_star names: one NCCL call per step inside the graph, on its own stream
            pb = measure_fused(ctx, wl, args.steps, 3, args.streams, reduce_mode="per_batch", alone_launches=8,
                               dev_sets=r["dev_sets"])
            extras.append({"name": "cfg2_per_batch_allreduce", "metric": METRIC, "value": pb["value"], "unit": UNIT,
                           "n_gpus": ctx.world, "ms_per_step": pb["ms_per_step"], "scaling": "weak", "dtype": "f32",
                           "config": dict(wl.config(), reduction="one NCCL all-reduce of the 32-byte [loss, SI-SDR, SDR, n] "
                                          "row PER STEP, captured in the graph on a dedicated communication stream"),
                           "timing": {"replays": pb["replays"], "graph_steps": pb["graph_steps"], "timed_ms": pb["ms_total"]},
                           "roofline": pb["roofline"], "clocks": pb["clocks"],
                           "vs_bucketed": pb["ms_per_step"] / r["ms_per_step"]})
            # ... and with no collective call at all: the kernel's epilogue pushes the sums into every rank's inbox
            ps = measure_fused(ctx, wl, args.steps, 3, args.streams, reduce_mode="push", alone_launches=8,
                               dev_sets=r["dev_sets"])
            extras.append({"name": "cfg2_per_batch_push", "metric": METRIC, "value": ps["value"], "unit": UNIT,
                           "n_gpus": ctx.world, "ms_per_step": ps["ms_per_step"], "scaling": "weak", "dtype": "f32",
                           "config": dict(wl.config(), reduction="PER STEP, no collective call: the fused kernel stores its "
                                          "[loss, SI-SDR, SDR, n] row into every rank's inbox over NVLink (peer-mapped "
                                          "memory, sep_fused_separate_push_f32); readers add the rows"),
                           "timing": {"replays": ps["replays"], "graph_steps": ps["graph_steps"], "timed_ms": ps["ms_total"]},
                           "roofline": ps["roofline"], "clocks": ps["clocks"], "check": ps["push_check"],
                           "vs_bucketed": ps["ms_per_step"] / r["ms_per_step"]})
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(ctx, wl, args.steps)
    r.pop("dev_sets")
    torch.cuda.empty_cache()
    if not args.no_extras:
        ex_steps = 240
        for fn in (lambda: fused_extra(ctx, "cfg2_hann_256_64", Workload("cfg2", wl.batch, wl.seconds, 2, 256, 64, "hann"),
                                       ex_steps, args.streams),
                   lambda: fused_extra(ctx, "cfg4_3spk_hann_512_128", Workload("cfg4", 64, 8.0, 3, 512, 128, "hann"),
                                       60, args.streams),
                   lambda: scoring_extra(ctx),
                   lambda: filterbank_extra(ctx),
                   lambda: conv1d_extra(ctx)):
            try:
                extras.append(fn())
            except Exception as exc:                      # an extra must never take the headline down with it
                extras.append({"error": "%s: %s" % (type(exc).__name__, exc)})
            torch.cuda.empty_cache()

    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_dict(cpu_throughput(wl, steps=3, warmup=1, budget_s=20.0), wl)

    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": wl.config(),
            "timing": {"replays": r["replays"], "steps_timed": r["replays"] * args.steps, "graph_steps": r["graph_steps"],
                       "graph_replays": r["graph_replays"], "timed_ms": r["ms_total"], "streams": r["n_streams"],
                       "launch": "%d x --steps = %d steps timed, captured in CUDA graphs of %d steps (independent steps "
                                 "round-robin on %d streams inside a graph)"
                                 % (r["replays"], r["replays"] * args.steps, r["graph_steps"], r["n_streams"]),
                       "parallelism": "utterance-sharded x%d, all-reduce of per-batch sums bucketed per replay "
                                      "(per-batch variant: extra.cfg2_per_batch_allreduce)" % ctx.world},
            "roofline": r["roofline"],
            "e2e": e2e, "gpu_launches": int(r["launches_per_step"] * args.steps * r["replays"] * ctx.world),
            "clocks": r["clocks"],
            "check": {"pit_loss_sum": r["sums"][0], "si_sdr_sum": r["sums"][1], "n": r["sums"][3]},
            "extra": extras,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        elif ctx.world > 1:
            line["cpu_baseline_note"] = "measured on rank 0 at N=1 only (and by --impl reference)"
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_sepcore(args)


if __name__ == "__main__":
    main()
