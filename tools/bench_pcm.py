"""Sample-format kernels against the HBM roofline: audiowrite's
float32 -> int16 conversion with peak normalisation (10 B per sample: two reads + one int16 write)
and int16 -> float32 decode (6 B per sample), on a 3000-utterance wsj0-shaped set (144 M samples)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import numpy as np, torch
import sepcore
from sepcore import _lib

rows, n = 3000, 48000
dev = torch.device("cuda", 0)
x = 0.3 * torch.randn((rows, n), device=dev)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {}
for name, fn, nbytes in (("audiowrite_int16(normalize=True)", lambda: sepcore.audiowrite_int16(x, True), 10.0 * rows * n),
                         ("audiowrite_int16(normalize=False)", lambda: sepcore.audiowrite_int16(x, False), 6.0 * rows * n)):
    pcm, _ = fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10): fn()
    torch.cuda.synchronize()
    ms, cnt = _lib.profile_collect()
    _lib.profile_enable(False)
    ms /= cnt
    out[name] = {"ms": ms, "GB/s": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
_lib.profile_enable(True)
for _ in range(10): f = sepcore.pcm16_to_float32(pcm)
torch.cuda.synchronize()
ms, cnt = _lib.profile_collect(); _lib.profile_enable(False); ms /= cnt
out["pcm16_to_float32"] = {"ms": ms, "GB/s": 6.0 * rows * n / ms / 1e6, "frac": 6.0 * rows * n / ms / 1e6 / peak}
# parity spot check
from oracle import signal_path as oracle
want, _ = oracle.audiowrite_int16(x[5].cpu().numpy(), True)
assert np.array_equal(sepcore.audiowrite_int16(x, True)[0][5].cpu().numpy(), want)
print(json.dumps({"metric": "sample-format kernels vs HBM roofline", "samples": rows * n, "peak_GB/s": peak, "kernels": out}))
