"""Repeats one fused cfg2 call many times and checks that every result is bit-identical to the first:
the in-kernel finalisation (completion counters, last strip of an utterance finalises it) must be deterministic."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import torch

import bench
import sepcore

ap = argparse.ArgumentParser()
ap.add_argument("--calls", type=int, default=3000)
ap.add_argument("--sources", type=int, default=2)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--shift", type=int, default=128)
a = ap.parse_args()
wl = bench.Workload("stress", 64, 4.0, a.sources, a.size, a.shift, "blackman")
d = {k: torch.from_numpy(v).cuda() for k, v in wl.make_set(seed=3).items()}
kw = wl.kw()
ref = sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], **kw)
torch.cuda.synchronize()
r0, e0, s0 = ref["scores"].clone(), ref["est"].clone(), ref["sums"].clone()
bad = 0
for i in range(a.calls):
    res = sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], **kw)
    if not (torch.equal(res["scores"], r0) and torch.equal(res["sums"], s0) and torch.equal(res["est"], e0)):
        bad += 1
torch.cuda.synchronize()
print("%d repeated calls (size %d, shift %d, %d sources): %d differ from the first" % (a.calls, a.size, a.shift, a.sources, bad))
sys.exit(1 if bad else 0)
