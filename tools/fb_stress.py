import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/speech-separation-project-with-ai_b200"]
import numpy as np
import sepcore
from oracle import signal_path as oracle
def rel_err(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
bad = 0
for trial in range(6):
    for n, n_src, want_code in [(2000, 2, True), (2000, 2, False), (32000, 2, True), (32000, 2, False), (1040, 1, True), (1032, 3, True), (2048, 2, True)]:
        rng = np.random.default_rng(n + n_src)
        batch, taps, filters, stride = 2, 16, 256, 8
        wave = (0.1 * rng.standard_normal((batch, n))).astype(np.float32)
        enc = (0.25 * rng.standard_normal((taps, filters))).astype(np.float32)
        dec = (0.06 * rng.standard_normal((filters, taps))).astype(np.float32)
        frames = (n - taps) // stride + 1
        masks = rng.random((batch, n_src, frames, filters)).astype(np.float32)
        out = sepcore.filterbank_separate(wave, enc, dec, masks, stride=stride, want_code=want_code)
        est = out[0] if want_code else out
        for b in range(batch):
            wc, we = oracle.filterbank_separate(wave[b], enc, dec, masks[b], stride)
            e = rel_err(est[b], we)
            if e > 1e-4:
                bad += 1
                d = np.abs(est[b] - we).max(axis=0)
                idx = np.nonzero(d > 1e-4 * np.abs(we).max())[0]
                print("trial", trial, (n, n_src, want_code), "b", b, "err %.3g" % e, "bad samples", len(idx), "first", idx[:4], "last", idx[-4:], "frames", idx[0] // 8, idx[-1] // 8)
print("bad", bad)
