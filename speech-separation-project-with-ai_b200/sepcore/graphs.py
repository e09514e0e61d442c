"""CUDA-graph replay of the fused hot path.

One fused step at BASELINE's 64 x 4 s batch is ~10-20 us of GPU time -- less
than a Python -> ctypes -> cudaLaunchKernel round trip -- so steady-state
callers capture a sequence of steps once and replay it.  Capture uses torch's
stream-capture plumbing around plain C-ABI calls; the library allocates nothing
during capture because the workspace is provided by the caller.
"""
from __future__ import annotations

from . import fused


class GraphedSeparator:
    """Captures `separate_and_score` over a rotating list of static buffer sets.

    buffer_sets: list of dicts with CUDA tensors 'mix' [B,N], 'masks' [B,C,T,F],
    optional 'refs' [B,C,N], 'frame_lengths', 'valid_samples'.  Each set gets its
    own output buffers ('est', 'scores', 'sums') -- replaying never allocates.
    steps: how many fused steps one replay runs; step s uses set s % len(sets).
    streams: independent steps are captured round-robin on this many streams (fork /
    join inside the graph), so the serial tail of one step (last tile -> utterance
    finalisation -> batch sums) and its last partial wave of CTAs overlap the next
    step's main phase.  Concurrent steps must never share outputs or a workspace (the
    in-kernel finalisation counters live there), so the number of streams actually used
    is the largest divisor of len(buffer_sets) that is <= `streams`: step s runs on lane
    s % n_streams with set s % len(buffer_sets), and because n_streams divides the number
    of sets every set is only ever touched by one lane (stream order serialises its reuse).
    reduce_each_step: None, or a callable `f(sums_row)` that all-reduces one step's
    [loss, SI-SDR, SDR, n] row in place (e.g. `sepcore.distributed.all_reduce_sums`).  It is
    captured into the graph on ONE dedicated communication stream, in step order -- the same
    total order on every rank, as NCCL requires -- waiting only on the step that produced the
    row, so the compute lanes never wait for the collective (per-batch all-reduce).
    push: None, or a `sepcore.distributed.PeerSums` with >= `steps` slots: step s also pushes its sums into slot
    s of every rank's inbox from inside the fused kernel (no collective call at all).
    """

    def __init__(self, buffer_sets, steps, size=256, shift=128, window=None, want_est=True, streams=1,
                 reduce_each_step=None, push=None):
        import torch

        self.sets = buffer_sets
        self.steps = int(steps)
        self.kw = dict(size=size, shift=shift, window=window, want_est=want_est)
        self.push = push
        first = buffer_sets[0]
        batch, n = (int(v) for v in first["mix"].shape)
        n_src = int(first["masks"].shape[1])
        dev = first["mix"].device
        nbytes = fused.workspace_bytes(batch, n_src, n, size, shift, window)
        scored = first.get("refs") is not None
        # per-STEP sums rows: the batch sums of every step survive the replay, so the
        # caller can all-reduce them in one bucket (one NCCL call per replay)
        self.sums = torch.zeros((self.steps, 4), dtype=torch.float64, device=dev) if scored else None
        self.outs = []
        for _ in buffer_sets:
            o = {"workspace": torch.zeros(nbytes, dtype=torch.uint8, device=dev)}   # zero-filled once
            if want_est:
                o["est"] = torch.empty((batch, n_src, n), dtype=torch.float32, device=dev)
            if first.get("refs") is not None:
                stride = fused.score_layout(n_src)["stride"]
                o["scores"] = torch.empty((batch, stride), dtype=torch.float64, device=dev)
            self.outs.append(o)
        # warm up outside capture (plan creation, function attributes, lazy module load)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(len(buffer_sets)):
                self._step(i, 0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        want = max(1, min(int(streams), len(buffer_sets), max(self.steps, 1)))
        self.n_streams = max(d for d in range(1, want + 1) if len(buffer_sets) % d == 0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            main = torch.cuda.current_stream(dev)
            if self.n_streams == 1 and reduce_each_step is None:
                for s in range(self.steps):
                    self._step(s % len(buffer_sets), s)
            else:
                lanes = [torch.cuda.Stream(device=dev) for _ in range(self.n_streams)]
                comm = torch.cuda.Stream(device=dev) if reduce_each_step is not None else None
                for lane in lanes + ([comm] if comm is not None else []):   # fork
                    lane.wait_stream(main)
                for s in range(self.steps):
                    lane = lanes[s % self.n_streams]
                    with torch.cuda.stream(lane):
                        self._step(s % len(buffer_sets), s)
                    if comm is not None and self.sums is not None:
                        done = torch.cuda.Event()
                        done.record(lane)
                        with torch.cuda.stream(comm):
                            comm.wait_event(done)
                            reduce_each_step(self.sums[s])
                for lane in lanes + ([comm] if comm is not None else []):   # join
                    main.wait_stream(lane)

    def _step(self, i, s):
        b, o = self.sets[i], dict(self.outs[i])
        if self.sums is not None:
            o["sums"] = self.sums[s]
        return fused.separate_and_score(
            b["mix"], b["masks"], b.get("refs"), b.get("frame_lengths"), b.get("valid_samples"),
            out=o, workspace=o["workspace"], push=None if self.push is None else self.push.target(s), **self.kw)

    def replay(self):
        self.graph.replay()

    def results(self, i):
        """Parsed outputs of buffer set i after a replay."""
        o = self.outs[i]
        res = {"est": o.get("est")}
        if "scores" in o:
            n_src = int(self.sets[i]["masks"].shape[1])
            res.update(fused.parse_scores(o["scores"], n_src))
        return res
