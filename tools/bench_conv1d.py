"""a14 (Raw_with_Convlayer.ipynb:389) alone: prints bench.py's conv1d extra line.  usage: python tools/bench_conv1d.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "speech-separation-project-with-ai_b200")]
import bench  # noqa: E402

ctx = bench.Ctx()
line = bench.conv1d_extra(ctx)
print(json.dumps({k: line[k] for k in ("name", "value", "ms_per_step", "roofline", "clocks", "check")}))
