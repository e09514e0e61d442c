"""Repeated-call check of the tcgen05 Conv1D kernels: every result bit-identical to the first one and within 1e-5 of the
oracle on sampled rows (hand-offs between the staging / MMA / epilogue / store roles are mbarriers only: a missed
dependency shows up as a flaky tile).  usage: python tools/stress_conv1d.py [calls]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import numpy as np
import torch

import sepcore
from oracle import signal_path as oracle

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for batch, rows, filters, act in [(64, 800, 129, "sigmoid"), (37, 1001, 129, "relu"), (16, 800, 17, "sigmoid"), (9, 999, 64, "sigmoid")]:
    gen = torch.Generator(device="cuda").manual_seed(rows + filters)
    x = 0.3 * torch.randn((batch, rows, 40), device="cuda", generator=gen)
    w = 0.2 * torch.randn((2, 40, filters), device="cuda", generator=gen)
    b = 0.1 * torch.randn((filters,), device="cuda", generator=gen)
    first = sepcore.conv1d(x, w, b, padding="same", activation=act).clone()
    want = oracle.conv1d(x[:2].cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), padding="same", activation=act)
    err = float(np.max(np.abs(first[:2].cpu().numpy() - want)))
    n_bad = 0
    for _ in range(calls):
        out = sepcore.conv1d(x, w, b, padding="same", activation=act)
        if not torch.equal(out, first):
            n_bad += 1
    print("batch %d rows %d filters %d %s: max err vs oracle %.2e, %d of %d calls differ from the first" % (batch, rows, filters, act, err, n_bad, calls))
    bad += n_bad + (err > 2e-5)
print("bad", bad)
sys.exit(1 if bad else 0)
