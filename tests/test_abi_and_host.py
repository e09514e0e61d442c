"""CPU tests of the boundary and the host logic: the C-ABI library loads and
exports every symbol include/sepcore.h declares (no compute calls without a
GPU), the product fails loudly without one, and the host-side mirror of the
reference's index logic agrees with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import PKG, ROOT

HEADER = os.path.join(ROOT, "include", "sepcore.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sep_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sepcore import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libsepcore.so does not export %s" % name
    # the ctypes prototype table mirrors the header one to one
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.sep_version() >= 100
    assert lib.sep_score_stride(2) == 3 * 4 + 2 + 6 and lib.sep_score_stride(3) == 3 * 9 + 6 + 6
    assert lib.sep_score_stride(0) < 0 and lib.sep_score_stride(5) < 0


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    for cite in ("parallel_stft.py:146-196", "parallel_stft.py:37-123", "uPIT_baseline.ipynb:1269-1307",
                 "uPIT_baseline.ipynb:1023-1059", "metrics/evaluate_metrics.py:14-92",
                 "Raw_with_Convlayer.ipynb:389"):
        assert cite in text, cite


def test_argument_errors_do_not_need_a_gpu():
    from sepcore import _lib
    lib = _lib.load()
    handle = ctypes.c_void_p()
    taps = np.ones(5000)
    rc = lib.sep_plan_create(ctypes.byref(handle), 5000, 100,
                             taps.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1)
    assert rc == _lib.ERR_UNSUPPORTED and b"size=5000 unsupported" in lib.sep_last_error()
    rc = lib.sep_plan_create(ctypes.byref(handle), 256, 0,
                             np.ones(256).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1)
    assert rc == _lib.ERR_INVALID
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED)
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_INVALID)


def test_no_cpu_fallback():
    """Without a CUDA device the product raises; it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sepcore
    with pytest.raises(sepcore.SepcoreError):
        sepcore.stft(np.zeros(1000, np.float32), time_dim=0, size=256, shift=128)
    with pytest.raises(sepcore.SepcoreError):
        sepcore.si_sdr(np.ones(10, np.float32), np.ones(10, np.float32))
    # nothing in the product imports the oracle
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, os.path.join(dirpath, f)


def test_missing_library_is_an_import_error(monkeypatch):
    from sepcore import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsepcore.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_host_index_logic_matches_oracle(reference_run):
    import sepcore
    from oracle import signal_path as sp
    for n, size, shift, frames, samples in reference_run["geometry"]:
        assert sepcore._samples_to_stft_frames(n, size, shift) == frames
        assert sepcore._stft_frames_to_samples(frames, size, shift) == samples
    assert np.array_equal(sepcore.segment_axis(np.arange(10), 4, 2), reference_run["segment_doc"])
    ragged = np.arange(46).reshape(2, 23)
    for end in ("cut", "pad", "wrap"):
        got = sepcore.segment_axis(ragged, 5, 2, axis=1, end=end, endvalue=-1)
        assert np.array_equal(got, reference_run[f"segment_{end}"])
    view = sepcore.segment_axis(np.arange(12.0), 4, 2)
    assert view.base is not None                               # a view, like the reference
    x = np.arange(24.0).reshape(2, 3, 4)
    assert np.array_equal(sepcore.segment_axis(x, 2, 1, axis=2), sp.segment_axis(x, 2, 1, axis=2))
    with pytest.raises(ValueError):
        sepcore.segment_axis(np.arange(10), 4, 5)
    assert sepcore.segment_raw(np.arange(85.0), 40).shape == (3, 40)


def test_drop_in_module_names():
    """The reference's module paths and names resolve to the CUDA-backed functions."""
    import parallel_stft
    import upit
    from metrics import evaluate_metrics
    for name in ("segment_axis", "_samples_to_stft_frames", "_stft_frames_to_samples", "stft"):
        assert callable(getattr(parallel_stft, name))
    for name in ("istft", "_biorthogonal_window_loopy", "pit_with_outputsize"):
        assert callable(getattr(upit, name))
    for name in ("wavread", "pow_np_norm", "pow_norm", "si_sdr", "permute_si_sdr", "eval_si_sdr", "eval_sdr"):
        assert callable(getattr(evaluate_metrics, name))
    assert upit.pit_with_outputsize(129).__name__ == "pit_loss"     # custom_objects key, :1374
    import inspect
    sig = inspect.signature(parallel_stft.stft)
    assert list(sig.parameters) == ["time_signal", "time_dim", "size", "shift", "window", "fading",
                                    "window_length"]
    assert sig.parameters["size"].default == 1024 and sig.parameters["shift"].default == 256
    sig = inspect.signature(upit.istft)
    assert list(sig.parameters) == ["stft_signal", "size", "shift", "window", "fading", "window_length"]


def test_score_layout_matches_library():
    import sepcore
    from sepcore import _lib
    for c in (1, 2, 3, 4):
        assert sepcore.score_layout(c)["stride"] == _lib.load().sep_score_stride(c)


def test_warp_fft_index_algebra_emulation():
    """tools/emulate_wfft.py restates the lane / register / shared-memory index algebra of csrc/fft256w.cuh
    (8 x 4 x 8 whole-warp FFT, both exchanges, pair split and Hermitian merge) in numpy: it must reproduce
    numpy.fft -- the layout rules the kernel is written against."""
    import importlib.util
    import numpy as np

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("emulate_wfft", os.path.join(root, "tools", "emulate_wfft.py"))
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    X = em.from_lanes(em.wfft256(em.to_lanes(x)))
    assert np.abs(X - np.fft.fft(x)).max() < 1e-10
    assert np.abs(em.from_lanes(em.wfft256(em.to_lanes(X), inv=True)) / 256 - x).max() < 1e-12
    a, b = rng.standard_normal(256), rng.standard_normal(256)
    XA, XB = em.split_planar(em.wfft256(em.to_lanes(0.5 * (a + 1j * b))))
    FA, FB = np.fft.rfft(a), np.fft.rfft(b)
    for q in range(32):
        for r in range(5):
            k = q + 32 * r
            if k <= 128 and (r < 4 or q == 0):
                assert abs(XA[q, r] - FA[k]) < 1e-10 and abs(XB[q, r] - FB[k]) < 1e-10
    y = em.from_lanes(em.wfft256(em.merge_pair(XA + 1j * XB, np.conj(XA) + 1j * np.conj(XB)), inv=True)) / 256
    assert np.abs(y.real - a).max() < 1e-12 and np.abs(y.imag - b).max() < 1e-12
