"""Load the UNMODIFIED reference functions from /root/reference (authoring container only).

TEST INFRASTRUCTURE -- used by ``oracle/make_golden.py`` to generate the
committed golden vectors and by the CPU tests (when the tree is present) to pin
the numpy restatement in ``oracle/signal_path.py``.  The GPU box has no
/root/reference: nothing in the ``-m gpu`` tests or ``smoke()`` calls this module;
``bench.py --impl reference`` / ``cpu_baseline`` use it there through the copies in
``oracle/_ref/`` (``oracle/build_ref.py``) to TIME the unmodified reference on the host cores.

The reference modules import tensorflow / librosa / soundfile / museval at the
top and use ``np.int``, ``scipy.signal.blackman`` and ``scipy.zeros``, all gone
from current numpy/scipy.  Empty stub modules and three aliases make them
import and run unmodified.  ``istft`` and
``_biorthogonal_window_loopy`` exist only in ``uPIT_baseline.ipynb`` cells
38-39; their source is exec'd as-is.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the authoring container; on the GPU box the byte-for-byte copies that
    oracle/build_ref.py put under oracle/_ref/ (git-ignored, travels with the snapshot)."""
    for root in (os.environ.get("SEP_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if root and os.path.isfile(os.path.join(root, "parallel_stft.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _default_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "parallel_stft.py"))


_cache = {}


def load():
    """Returns a namespace object with the reference's own callables:
    stft, segment_axis, _samples_to_stft_frames, _stft_frames_to_samples,
    istft, _biorthogonal_window_loopy, si_sdr, permute_si_sdr, pow_norm,
    pow_np_norm."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    import scipy
    import scipy.signal

    for name in ("tensorflow", "librosa", "librosa.display", "soundfile", "museval"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["librosa"].display = sys.modules["librosa.display"]
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(scipy.signal, "blackman"):
        scipy.signal.blackman = scipy.signal.windows.blackman
    if not hasattr(scipy, "zeros"):
        scipy.zeros = np.zeros

    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "metrics")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the product ships drop-in modules with the same names; make sure the
    # reference's own files are the ones imported here
    saved = {k: sys.modules.pop(k, None) for k in ("parallel_stft", "evaluate_metrics")}
    try:
        ps = importlib.import_module("parallel_stft")
        em = importlib.import_module("evaluate_metrics")
        assert os.path.dirname(ps.__file__) == REFERENCE_ROOT, ps.__file__
    finally:
        for k in ("parallel_stft", "evaluate_metrics"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
        for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "metrics")):
            if p in sys.path:
                sys.path.remove(p)

    with open(os.path.join(REFERENCE_ROOT, "uPIT_baseline.ipynb")) as fh:
        nb = json.load(fh)
    scope = {"np": np, "scipy": scipy, "signal": scipy.signal, "irfft": np.fft.irfft}
    for cell in (38, 39):
        exec("".join(nb["cells"][cell]["source"]), scope)
    # cell 40: audiowrite -- run unmodified, with the wav writer replaced by a capture
    import threading
    captured = {}

    def _capture(path, samplerate, data):
        captured["data"] = np.array(data, copy=True)

    if not hasattr(np, "float"):
        np.float = float
    scope40 = {"np": np, "threading": threading, "wav_write": _capture}
    exec("".join(nb["cells"][40]["source"]), scope40)

    def audiowrite_int16(data, normalize=False):
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            clipped = scope40["audiowrite"](data, "unused.wav", 16000, normalize, False)
        return captured["data"], int(clipped)

    ns = types.SimpleNamespace(
        stft=ps.stft,
        segment_axis=ps.segment_axis,
        _samples_to_stft_frames=ps._samples_to_stft_frames,
        _stft_frames_to_samples=ps._stft_frames_to_samples,
        istft=scope["istft"],
        _biorthogonal_window_loopy=scope["_biorthogonal_window_loopy"],
        si_sdr=em.si_sdr,
        permute_si_sdr=em.permute_si_sdr,
        pow_norm=em.pow_norm,
        pow_np_norm=em.pow_np_norm,
        audiowrite_int16=audiowrite_int16,
    )
    _cache["ns"] = ns
    return ns
