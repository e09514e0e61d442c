import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/speech-separation-project-with-ai_b200"]
import numpy as np, torch
import sepcore
from oracle import signal_path as oracle
def rel_err(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for mode in ("host", "device"):
    bad = tot = 0
    for trial in range(4):
        for n, n_src in [(2000, 2), (32000, 2), (1032, 3)]:
            rng = np.random.default_rng(n + n_src)
            batch, taps, filters, stride = 2, 16, 256, 8
            wave = (0.1 * rng.standard_normal((batch, n))).astype(np.float32)
            enc = (0.25 * rng.standard_normal((taps, filters))).astype(np.float32)
            dec = (0.06 * rng.standard_normal((filters, taps))).astype(np.float32)
            frames = (n - taps) // stride + 1
            masks = rng.random((batch, n_src, frames, filters)).astype(np.float32)
            args = (wave, enc, dec, masks)
            if mode == "device": args = tuple(torch.from_numpy(x).cuda() for x in args)
            est, code = sepcore.filterbank_separate(*args, stride=stride, want_code=True)
            if mode == "device": est = est.cpu().numpy()
            for b in range(batch):
                wc, we = oracle.filterbank_separate(wave[b], enc, dec, masks[b], stride)
                tot += 1
                bad += rel_err(est[b], we) > 1e-4
    print(mode, "nodump" if os.environ.get("SEPCORE_FB_NODUMP") else "dump", "bad", bad, "of", tot)
