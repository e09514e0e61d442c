"""Recipe for oracle/_ref/: the UNMODIFIED reference files the hot path lives in, copied byte for byte
from /root/reference so that they travel to the GPU box (oracle/_ref/ is git-ignored, NOT
gpurun-ignored: it never enters the history, like a built .so).

The reference is pure Python (no compiled code), so "building" it is a copy:
    parallel_stft.py, metrics/evaluate_metrics.py, uPIT_baseline.ipynb (cells 38-40: istft,
    _biorthogonal_window_loopy, audiowrite are defined only there).
oracle/reference_loader.py imports them with stub modules for tensorflow / librosa / soundfile /
museval.  `bench.py --impl reference` times them when this directory exists
(cpu_baseline.kind = "reference"), else the numpy restatement (kind = "port").

TEST / BASELINE INFRASTRUCTURE -- never imported by the product.
usage: python oracle/build_ref.py   (also run by __graft_entry__.build() when /root/reference exists)"""
import hashlib
import json
import os
import shutil

SRC = os.environ.get("SEP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("parallel_stft.py", os.path.join("metrics", "evaluate_metrics.py"), "uPIT_baseline.ipynb")


def build():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        os.chmod(dst, 0o644)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref built" if build() else "reference tree not found at %s" % SRC)
