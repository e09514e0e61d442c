"""Generates tests/golden/record_golden.npz (authoring container only: needs /root/reference).

The independent pure-Python encoder of tests/test_records.py must first reproduce the reference's committed
.tfrecords files byte for byte; it then encodes a few frames of the same arrays (frames 100..104 of two STFT
records, samples 1000..1039 of a raw-waveform record) in the files' own map order -- a small fixture that travels
to the GPU box.  TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "speech-separation-project-with-ai_b200")]

from oracle import tfrecord_reader as tr  # noqa: E402
import test_records as T  # noqa: E402

CASES = [("tr_tfrecord", "447o0302_0.62948_441c0212_-0.62948.tfrecords"),
         ("cv_one_source_tfrecord", "447o0302_1.3388_22ho010i_-1.3388_s2.tfrecords"),
         ("tt_raw_tfrecord", "447o0302_0.62948_441c0212_-0.62948.tfrecords")]


def main():
    out = {}
    for i, (d, f) in enumerate(CASES):
        path = os.path.join(T.REF, d, f)
        data = open(path, "rb").read()
        rec = next(tr.records(path))
        ex = tr.read_sequence_example(path)
        order = T._key_order(rec)
        inputs, labels = np.stack(ex["inputs"]), np.stack(ex["labels"])
        length, name = float(ex["length"][0][0]), ex["name"][0][0]
        if i < 2:
            assert T.py_encode(inputs, labels, length, name, order) == data, "encoder does not reproduce " + path
        sl = slice(100, 105) if i < 2 else slice(1000, 1040)
        a, b = inputs[sl], labels[sl]
        out.update({"inputs_%d" % i: a, "labels_%d" % i: b, "length_%d" % i: np.float32(length),
                    "name_%d" % i: name.decode(), "order_%d" % i: np.array(order),
                    "record_%d" % i: np.frombuffer(T.py_encode(a, b, length, name, order), dtype=np.uint8)})
    out["count"] = len(CASES)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "record_golden.npz"), **out)


if __name__ == "__main__":
    main()
