"""Runs a few eager fused steps (no CUDA graph) -- the command profiled under ncu.
usage: python tools/run_steps.py [--steps 6] [--size 256 --shift 128 --sources 2 --batch 64]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]

import numpy as np
import torch

import bench
import sepcore

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--sources", type=int, default=2)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--shift", type=int, default=128)
ap.add_argument("--window", default="blackman")
args = ap.parse_args()
wl = bench.Workload("run_steps", args.batch, args.seconds, args.sources, args.size, args.shift, args.window)
sets = [{k: torch.from_numpy(v).cuda() for k, v in wl.make_set(seed=i).items()} for i in range(3)]
kw = wl.kw()
for s in range(args.steps):
    d = sets[s % 3]
    res = sepcore.separate_and_score(d["mix"], d["masks"], d["refs"], **kw)
torch.cuda.synchronize()
print("ran", args.steps, "steps; pit loss sum", float(res["sums"][0]))
