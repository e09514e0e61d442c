"""Generate the committed golden vectors under tests/golden/ (authoring container only).

TEST INFRASTRUCTURE.  Run as ``python -m oracle.make_golden`` from the repo root
with the reference tree at /root/reference.  Three files are written:

* ``wsj0_fixture.npz``     -- the reference's committed 8 kHz int16 wsj0-2mix
  toy set (``mycode/wsj0_2mix/use_this/tt/{mix,s1,s2}``, 4 utterances) and the 8
  estimated outputs under ``test_wav/`` -- data fixtures, no code.
* ``tfrecord_golden.npz``  -- a strided subset of the reference's OWN STFT
  outputs (``mycode/tfrecords/tr_tfrecord`` and ``tr_one_source_tfrecord``:
  |X|, angle X, PSA labels, per-speaker |S|, valid lengths) plus full-array
  float64 checksums, decoded with ``oracle/tfrecord_reader.py``.
* ``reference_run.npz``    -- outputs of the UNMODIFIED reference functions
  (``oracle/reference_loader.py``) on small seeded inputs: stft at several
  parameterisations, istft, biorthogonal windows, segment_axis, frame counts,
  si_sdr / permute_si_sdr on the committed ref/est pairs.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.io.wavfile as wavfile
import scipy.signal.windows as windows

from . import reference_loader, tfrecord_reader

ROOT = reference_loader.REFERENCE_ROOT
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

UTTS = [
    "447o0302_0.62948_441c0212_-0.62948",
    "447o0302_1.3388_22ho010i_-1.3388",
    "447o0302_2.1067_422o030k_-2.1067",
    "447o0303_0.14144_441c0212_-0.14144",
]
FRAME_STRIDE = 9          # keep every 9th frame of the 626-frame golden records


def _wav(path):
    sr, data = wavfile.read(path)
    assert sr == 8000 and data.dtype == np.int16 and data.ndim == 1
    return data


def make_fixture():
    arrays = {"names": np.array(UTTS)}
    for i, name in enumerate(UTTS):
        for kind in ("mix", "s1", "s2"):
            arrays[f"tt_{kind}_{i}"] = _wav(f"{ROOT}/mycode/wsj0_2mix/use_this/tt/{kind}/{name}.wav")
        for kind in ("s1", "s2"):
            arrays[f"est_{kind}_{i}"] = _wav(f"{ROOT}/test_wav/{name}_{kind}.wav")
    np.savez_compressed(os.path.join(OUT, "wsj0_fixture.npz"), **arrays)


def make_tfrecord_golden():
    arrays = {"names": np.array(UTTS), "frame_stride": np.array(FRAME_STRIDE)}
    for i, name in enumerate(UTTS):
        rec = tfrecord_reader.read_mixed(f"{ROOT}/mycode/tfrecords/tr_tfrecord/{name}.tfrecords")
        assert rec["name"] == name and rec["inputs"].shape == (626, 258)
        arrays[f"inputs_{i}"] = rec["inputs"][::FRAME_STRIDE].astype(np.float32)
        arrays[f"labels_{i}"] = rec["labels"][::FRAME_STRIDE].astype(np.float32)
        arrays[f"length_{i}"] = np.array(rec["length"])
        mag = rec["inputs"][:, :129].astype(np.float64)
        lab = rec["labels"].astype(np.float64)
        arrays[f"checksum_{i}"] = np.array([mag.sum(), (mag * mag).sum(), lab.sum(), (lab * lab).sum()])
        for kind in ("s1", "s2"):
            ex = tfrecord_reader.read_sequence_example(
                f"{ROOT}/mycode/tfrecords/tr_one_source_tfrecord/{name}_{kind}.tfrecords")
            arrays[f"{kind}_mag_{i}"] = np.stack(ex["inputs"])[::FRAME_STRIDE].astype(np.float32)
            arrays[f"{kind}_len_{i}"] = np.array(float(ex["length"][0][0]))
    np.savez_compressed(os.path.join(OUT, "tfrecord_golden.npz"), **arrays)


def make_reference_run():
    ref = reference_loader.load()
    rng = np.random.default_rng(20261018)
    arrays = {}

    # --- frame geometry and framing ---
    geo = []
    for n, size, shift in [(32256, 256, 128), (80256, 256, 128), (32384, 256, 64),
                           (64768, 512, 128), (100, 256, 128), (256, 256, 128), (1000, 64, 16)]:
        t = int(ref._samples_to_stft_frames(n, size, shift))
        geo.append([n, size, shift, t, int(ref._stft_frames_to_samples(t, size, shift))])
    arrays["geometry"] = np.array(geo)
    arrays["segment_doc"] = ref.segment_axis(np.arange(10), 4, 2)
    ragged = np.arange(46).reshape(2, 23)
    for end in ("cut", "pad", "wrap"):
        arrays[f"segment_{end}"] = ref.segment_axis(ragged, 5, 2, axis=1, end=end, endvalue=-1)

    # --- STFT at the parameterisations used by the reference and by BASELINE ---
    wave = (0.1 * rng.standard_normal(3000)).astype(np.float32)
    arrays["wave"] = wave
    cfgs = {
        "blackman_256_128": dict(size=256, shift=128, window=windows.blackman),
        "blackman_256_64": dict(size=256, shift=64, window=windows.blackman),
        "hann_256_64": dict(size=256, shift=64, window=windows.hann),
        "hann_512_128": dict(size=512, shift=128, window=windows.hann),
        "hamming_128_32": dict(size=128, shift=32, window=windows.hamming),
        "blackman_1024_256": dict(size=1024, shift=256, window=windows.blackman),
    }
    for key, kw in cfgs.items():
        spec = ref.stft(wave, time_dim=0, **kw)
        arrays[f"stft_{key}"] = spec
        arrays[f"istft_of_stft_{key}"] = ref.istft(spec, **kw)
        arrays[f"synth_{key}"] = ref._biorthogonal_window_loopy(kw["window"](kw["size"]), kw["shift"])
        t, f = spec.shape
        rand_spec = rng.standard_normal((min(t, 12), f)) + 1j * rng.standard_normal((min(t, 12), f))
        arrays[f"randspec_{key}"] = rand_spec
        arrays[f"istft_rand_{key}"] = ref.istft(rand_spec, **kw)
    arrays["stft_nofade"] = ref.stft(wave, time_dim=0, size=256, shift=128, fading=False)
    arrays["stft_winlen200"] = ref.stft(wave, time_dim=0, size=256, shift=128, window_length=200)
    arrays["istft_winlen200"] = ref.istft(arrays["stft_winlen200"], size=256, shift=128, window_length=200)
    batch = (0.1 * rng.standard_normal((3, 2000))).astype(np.float32)
    arrays["wave_batch"] = batch
    arrays["stft_batch_default_dim"] = ref.stft(batch, size=256, shift=128)
    arrays["stft_batch_time0"] = ref.stft(np.ascontiguousarray(batch.T), time_dim=0, size=256, shift=128)

    # --- SI-SDR on the committed reference / estimate pairs ---
    rows = []
    for name in UTTS:
        r1 = _wav(f"{ROOT}/mycode/wsj0_2mix/use_this/tt/s1/{name}.wav").astype(np.float32) / 32768
        r2 = _wav(f"{ROOT}/mycode/wsj0_2mix/use_this/tt/s2/{name}.wav").astype(np.float32) / 32768
        e1 = _wav(f"{ROOT}/test_wav/{name}_s1.wav").astype(np.float32) / 32768
        e2 = _wav(f"{ROOT}/test_wav/{name}_s2.wav").astype(np.float32) / 32768
        n = min(r1.size, e1.size)
        r1, r2, e1, e2 = r1[:n], r2[:n], e1[:n], e2[:n]
        rows.append([ref.si_sdr(r1, e1), ref.si_sdr(r2, e2), ref.si_sdr(r1, e2), ref.si_sdr(r2, e1),
                     ref.permute_si_sdr(r1, r2, e1, e2), n])
    arrays["si_sdr_table"] = np.array(rows, dtype=np.float64)
    arrays["si_sdr_mean"] = np.array(np.mean(np.array([np.float32(r[4]) for r in rows], dtype=np.float32)))
    np.savez_compressed(os.path.join(OUT, "reference_run.npz"), **arrays)


def main():
    os.makedirs(OUT, exist_ok=True)
    make_fixture()
    make_tfrecord_golden()
    make_reference_run()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
