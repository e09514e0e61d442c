"""The feature / label records the reference writes next to the signal path.

Reference: `make_sequence_example(inputs, labels, length, name)` -- parallel_stft_single.py:238-254
(parallel_stft.py:217-229) -- and `with tf.io.TFRecordWriter(path) as writer:
writer.write(ex.SerializeToString())` -- :287-309.  TensorFlow is not part of this stack: the
SequenceExample wire format and the TFRecord framing are produced by libsepcore (sep_record_encode, host
code), byte for byte what TensorFlow writes (checked against the reference's committed .tfrecords files).
The names below mirror the reference's so that its writer loop ports line by line.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

KEYS = ("inputs", "labels", "length", "name")


class SequenceExample:
    """What `make_sequence_example` returns: holds the arrays; `SerializeToString()` gives the protobuf
    payload (without the TFRecord framing), `record()` the framed record."""

    def __init__(self, inputs, labels, length, name, key_order=None):
        self.inputs = np.ascontiguousarray(np.asarray(inputs), dtype=np.float32)
        self.labels = np.ascontiguousarray(np.asarray(labels), dtype=np.float32)
        if self.inputs.ndim != 2 or self.labels.ndim != 2 or self.inputs.shape[0] != self.labels.shape[0]:
            raise ValueError("inputs and labels must be [T, W_in] and [T, W_lab] with the same T")
        self.length = float(length)
        self.name = name.encode("utf-8") if isinstance(name, str) else bytes(name)
        if key_order is not None:
            key_order = [KEYS.index(k) if isinstance(k, str) else int(k) for k in key_order]
        self.key_order = key_order

    def record(self):
        lib = _lib.load()
        frames, w_in = self.inputs.shape
        w_lab = self.labels.shape[1]
        size = C.c_int64()
        _lib.check(lib.sep_record_size(frames, w_in, w_lab, len(self.name), C.byref(size)), "sep_record_size")
        out = np.empty(size.value, dtype=np.uint8)
        written = C.c_int64()
        order = None if self.key_order is None else (C.c_int32 * 4)(*self.key_order)
        _lib.check(lib.sep_record_encode(self.inputs.ctypes.data, self.labels.ctypes.data, frames, w_in, w_lab,
                                         C.c_float(self.length), self.name, len(self.name), order,
                                         out.ctypes.data, size.value, C.byref(written)), "sep_record_encode")
        return out[:written.value].tobytes()

    def SerializeToString(self):
        return self.record()[12:-4]


def make_sequence_example(inputs, labels, length, name, genders=False, key_order=None):
    """Reference signature (parallel_stft_single.py:238): inputs [T, 2F] (|X| ++ angle X), labels [T, C F],
    length = frames of the unpadded utterance, name = utterance id."""
    return SequenceExample(inputs, labels, length, name, key_order)


class TFRecordWriter:
    """`tf.io.TFRecordWriter` for this format: `write()` takes what `SerializeToString()` returned (or a
    SequenceExample) and appends the framed record."""

    def __init__(self, path):
        self._fh = open(path, "wb")

    def write(self, payload):
        if isinstance(payload, SequenceExample):
            self._fh.write(payload.record())
            return
        import struct

        lib = _lib.load()
        # frame a raw payload: reuse the encoder's CRC through a zero-feature record is not possible, so the
        # framing of foreign payloads is done here with the same masked CRC-32C (C helper below)
        crc = lambda b: int(lib.sep_record_masked_crc(b, len(b)))
        head = struct.pack("<Q", len(payload))
        self._fh.write(head + struct.pack("<I", crc(head)) + payload + struct.pack("<I", crc(payload)))

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def gen_feats_record(mix_wav, s1_wav, s2_wav, max_len, part_name, size=256, shift=128, window=None):
    """The 'mixed' branch of the reference's gen_feats (parallel_stft_single.py:257-311) without the file I/O:
    zero-pad the three waveforms to max_len, STFT features on the GPU (inputs = |X| ++ angle X, labels = PSA
    labels of both sources), length = frames of the UNPADDED mixture; returns the SequenceExample."""
    from .plan import get_plan
    from .signal_path import stft_features

    pad = lambda x: np.pad(np.asarray(x, dtype=np.float32), (0, max_len - len(x)), "constant", constant_values=(0))
    mix = pad(mix_wav)[None]
    refs = np.stack([pad(s1_wav), pad(s2_wav)])[None]
    feats, labels = stft_features(mix, refs, size=size, shift=shift, window=window)
    frames = get_plan(size, shift, window, True).frames(len(mix_wav))
    return make_sequence_example(feats[0], labels[0], frames, part_name)
