"""BSS Eval v4 (sep_bss_eval_f32) on a wsj0-2mix-shaped set: utterances per second and the share of each kernel.
usage: python tools/bench_bss.py [--utts 300]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sepcore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=300)
args = ap.parse_args()
rng = np.random.default_rng(4)
lens = (8000 * rng.uniform(2, 10, size=args.utts)).astype(np.int64)
refs, ests = [], []
for n in lens:
    r = (0.1 * rng.standard_normal((2, n))).astype(np.float32)
    r[1, 1:] += 0.9 * r[1, :-1]                      # one coloured source
    e = (np.array([[0.9, 0.2], [0.1, 0.8]], np.float32) @ r + 0.03 * rng.standard_normal((2, n)).astype(np.float32))
    refs.append(r)
    ests.append(e.astype(np.float32))
sepcore.bss_eval_batch(refs[:8], ests[:8], 2)        # warm up
torch.cuda.synchronize()
t0 = time.perf_counter()
res = sepcore.bss_eval_batch(refs, ests, 2)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
audio = float(lens.sum()) / 8000.0
from oracle import bss_eval as B  # noqa: E402
t1 = time.perf_counter()
want = B.bss_eval(refs[0], ests[0])
t_cpu = time.perf_counter() - t1
print(json.dumps({"metric": "BSS Eval v4 (512 taps, 2 sources), whole host-pointer call", "utterances": args.utts,
                  "audio_s": audio, "seconds": dt, "utt_per_s": args.utts / dt, "audio_s_per_s": audio / dt,
                  "cpu_restatement_s_per_utt": t_cpu,
                  "check_sir_db": [float(res["sir"][0][0, 0]), float(want[5][2, 0, 0])],
                  "mean_value_db": float(np.mean(res["value"]))}))
