"""Drop-in for the reference's `metrics/evaluate_metrics.py` (same module path,
names and signatures); the arithmetic runs in libsepcore on the GPU.

wav decoding is I/O and out of scope of the CUDA path: `wavread` uses scipy
(soundfile is not installed here); for the 16-bit PCM the reference works with,
sf.read(dtype='float32') and int16 / 32768 are the same values.
"""
import os

import numpy as np

from sepcore.scoring import (permute_si_sdr, pow_norm, pow_np_norm, score_batch, si_sdr,  # noqa: F401
                             truncate_to_min_len)


def wavread(path):
    """evaluate_metrics.py:7-12."""
    from scipy.io import wavfile

    sample_rate, data = wavfile.read(path)
    if data.dtype == np.int16:
        wav = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        wav = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        wav = data.astype(np.float32)
    return wav, sample_rate


def _load_pairs(wav_dir, test_dir):
    """The directory walk and truncate-to-min-length of evaluate_metrics.py:37-48."""
    refs, ests = [], []
    for name in os.listdir(wav_dir + 'tt/mix'):
        ref_s1, _ = wavread(wav_dir + 'tt/s1/' + name)
        ref_s2, _ = wavread(wav_dir + 'tt/s2/' + name)
        est_s1, _ = wavread(test_dir + name[:-4] + '_s1.wav')
        est_s2, _ = wavread(test_dir + name[:-4] + '_s2.wav')
        ref_s1, ref_s2, est_s1, est_s2 = truncate_to_min_len(ref_s1, ref_s2, est_s1, est_s2)
        refs.append([ref_s1, ref_s2])
        ests.append([est_s1, est_s2])
    return refs, ests


def eval_si_sdr(wav_dir, test_dir):
    """Mean over files of permute_si_sdr (evaluate_metrics.py:36-55): one batched
    GPU launch instead of a Python loop; float32 mean like the reference."""
    refs, ests = _load_pairs(wav_dir, test_dir)
    res = score_batch(refs, ests, 2)
    return np.mean(np.array([np.float32(v) for v in res["si_best"]]))


def eval_sdr(wav_dir, test_dir):
    """Mean over files of the mean SDR of the assignment museval selects (evaluate_metrics.py:57-92).
    The reference calls museval.metrics.bss_eval(reference, estimated, window=np.inf, hop=np.inf,
    compute_permutation=True): BSS Eval v4 image criteria with 512-tap time-invariant distortion filters,
    the permutation chosen by mean SIR, NaN fallback `np.mean(np.nan_to_num(sdr))` (:83-86).  Here the
    whole directory is one batched GPU call (sepcore.bss_eval_batch -> sep_bss_eval_f32).

    museval is third party and not installable in this stack: the CUDA path is checked against a float64
    restatement of museval's published algorithm (the test-side CPU checker) -- parity with museval itself is
    UNPINNED."""
    from sepcore.bss import bss_eval_batch

    refs, ests = _load_pairs(wav_dir, test_dir)
    res = bss_eval_batch(refs, ests, 2)
    return np.mean(np.array(res["value"]))
