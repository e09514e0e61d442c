"""Generates tests/golden/audiowrite_golden.npz by running the UNMODIFIED reference cell
(uPIT_baseline.ipynb cell 40, `audiowrite`, wav writer replaced by a capture) in the authoring
container, where /root/reference exists.  Usage: python -m oracle.make_golden_audio"""
import os

import numpy as np

from oracle import reference_loader


def main():
    ref = reference_loader.load()
    rng = np.random.default_rng(40)
    out = {}
    cases = {
        "quiet": (0.2 * rng.standard_normal(4001)).astype(np.float32),
        "loud": (0.7 * rng.standard_normal(3000)).astype(np.float32),        # clips without normalisation
        "edge": np.array([1.0, -1.0, 0.99996948, 1.0000305, -1.0000305, 0.0, 3.0517578e-05, -3.0517578e-05,
                          0.5, -0.5, 2.0, -2.0], dtype=np.float32),
    }
    for name, x in cases.items():
        out[name + "_x"] = x
        for norm in (False, True):
            pcm, clipped = ref.audiowrite_int16(x, norm)
            out["%s_%d_pcm" % (name, norm)] = pcm
            out["%s_%d_clipped" % (name, norm)] = np.int64(clipped)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "audiowrite_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("clipped")})


if __name__ == "__main__":
    main()
