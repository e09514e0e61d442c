"""Dependency-free reader for the reference's committed TFRecord golden vectors.

TEST INFRASTRUCTURE.  The reference's ``mycode/tfrecords/**`` files hold
``tf.train.SequenceExample`` protos written by uPIT_baseline.ipynb cell 10
(:439-553).  TensorFlow is not installed, so the two layers are decoded by hand:

* TFRecord framing: u64 length, u32 masked-crc(length), payload, u32 crc(payload)
* protobuf wire format of SequenceExample:
    SequenceExample { Features context = 1; FeatureLists feature_lists = 2; }
    FeatureLists    { map<string, FeatureList> feature_list = 1; }
    FeatureList     { repeated Feature feature = 1; }
    Feature         { oneof { BytesList = 1; FloatList = 2; Int64List = 3; } }
    FloatList       { repeated float value = 1 [packed]; }
"""
from __future__ import annotations

import struct

import numpy as np


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 2:
            n, pos = _varint(buf, pos)
            yield num, buf[pos:pos + n]
            pos += n
        elif wire == 0:
            v, pos = _varint(buf, pos)
            yield num, v
        elif wire == 5:
            yield num, buf[pos:pos + 4]
            pos += 4
        elif wire == 1:
            yield num, buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wire)


def _feature(buf):
    for num, body in _fields(buf):
        if num == 2:      # FloatList
            vals = []
            for n2, packed in _fields(body):
                if n2 == 1:
                    vals.append(np.frombuffer(bytes(packed), dtype="<f4"))
            return np.concatenate(vals) if vals else np.zeros(0, "<f4")
        if num == 1:      # BytesList
            return [bytes(b) for n2, b in _fields(body) if n2 == 1]
        if num == 3:      # Int64List
            out = []
            for n2, packed in _fields(body):
                if n2 == 1:
                    if isinstance(packed, int):
                        out.append(packed)
                    else:
                        p = 0
                        while p < len(packed):
                            v, p = _varint(packed, p)
                            out.append(v)
            return np.array(out, dtype=np.int64)
    return None


def records(path):
    with open(path, "rb") as fh:
        data = fh.read()
    pos = 0
    while pos < len(data):
        (n,) = struct.unpack_from("<Q", data, pos)
        pos += 12
        yield memoryview(data)[pos:pos + n]
        pos += n + 4


def read_sequence_example(path):
    """Returns {feature_list_name: [per-step value, ...]} of the first record."""
    rec = next(records(path))
    out = {}
    for num, body in _fields(rec):
        if num != 2:
            continue
        for n1, entry in _fields(body):        # map entries
            if n1 != 1:
                continue
            key, flist = None, None
            for n2, val in _fields(entry):
                if n2 == 1:
                    key = bytes(val).decode()
                elif n2 == 2:
                    flist = [_feature(f) for n3, f in _fields(val) if n3 == 1]
            out[key] = flist
    return out


def read_mixed(path):
    """A ``*_tfrecord`` file -> dict(inputs [T,2F], labels [T,2F], length, name)."""
    ex = read_sequence_example(path)
    return {
        "inputs": np.stack(ex["inputs"]),
        "labels": np.stack(ex["labels"]),
        "length": float(ex["length"][0][0]),
        "name": ex["name"][0][0].decode(),
    }
