import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/speech-separation-project-with-ai_b200"]
import numpy as np, torch
import sepcore
n, n_src = 2000, 2
rng = np.random.default_rng(5)
batch, taps, filters, stride = 2, 16, 256, 8
wave = (0.1 * rng.standard_normal((batch, n))).astype(np.float32)
enc = (0.25 * rng.standard_normal((taps, filters))).astype(np.float32)
dec = (0.06 * rng.standard_normal((filters, taps))).astype(np.float32)
frames = (n - taps) // stride + 1
masks = rng.random((batch, n_src, frames, filters)).astype(np.float32)
args = tuple(torch.from_numpy(x).cuda() for x in (wave, enc, dec, masks))
good = sepcore.filterbank_separate(*args, stride=stride).cpu().numpy()
for rep in range(2):
    est, code = sepcore.filterbank_separate(*args, stride=stride, want_code=True)
    est = est.cpu().numpy()
    diff = np.abs(est - good)
    scale = np.abs(good).max()
    for b in range(batch):
        for c in range(n_src):
            bad = np.nonzero(diff[b, c] > 1e-5 * scale)[0]
            fr = np.unique(bad // 8)
            print("rep", rep, "b", b, "c", c, "bad samples", len(bad), "bad hop blocks", len(fr), "of", frames + 1,
                  "first", fr[:12], "maxdiff/scale %.3g" % (diff[b, c].max() / scale))
    # per-frame decoded contribution check: recompute est for c=0 from the dumped code with numpy
    cd = code.cpu().numpy()
    for b in range(1):
        for c in range(n_src):
            y = (cd[b] * masks[b, c]).astype(np.float64) @ dec.astype(np.float64)   # [K, 16]
            ref = np.zeros(est.shape[-1])
            for k in range(frames): ref[k * 8:k * 8 + 16] += y[k]
            # leave-one-chunk-out: which chunk's absence explains the error at the worst hop block?
            d = est[b, c] - ref
            worst = int(np.argmax(np.abs(d)) // 8)
            k = worst
            errs = []
            for j in range(16):
                yj = (cd[b, k, 16 * j:16 * j + 16] * masks[b, c, k, 16 * j:16 * j + 16]).astype(np.float64) @ dec[16 * j:16 * j + 16].astype(np.float64)
                errs.append(float(np.abs(d[k * 8:k * 8 + 8] + yj[:8]).max()))
            print("   b", b, "c", c, "worst block", worst, "row in tile", worst % 127, "resid if chunk j added back:", np.round(np.array(errs) / scale, 4))
