"""STFT plans: the host evaluates the reference's window callable once, the
library caches windows, twiddles and the biorthogonal synthesis window
(uPIT_baseline.ipynb:1234-1259, cell 38) on the device."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib


def default_window():
    """scipy's symmetric Blackman: the reference's `signal.blackman` default
    (parallel_stft.py:147), which modern scipy only exposes under signal.windows."""
    from scipy.signal import windows

    return windows.blackman


def evaluate_window(window, size, window_length=None):
    """parallel_stft.py:183-187: window(size), or window(window_length) zero padded."""
    if window is None:
        window = default_window()
    if callable(window):
        if window_length is None:
            taps = np.asarray(window(size), dtype=np.float64)
        else:
            taps = np.asarray(window(window_length), dtype=np.float64)
            taps = np.pad(taps, (0, size - window_length), mode="constant")
    else:
        taps = np.asarray(window, dtype=np.float64)
    if taps.shape != (size,):
        raise ValueError("window must evaluate to `size`=%d taps, got %r" % (size, taps.shape))
    return np.ascontiguousarray(taps)


class Plan:
    """Immutable handle on a sep_plan bound to one CUDA device."""

    def __init__(self, size, shift, taps, fading=True):
        lib = _lib.load()
        self.size, self.shift, self.fading = int(size), int(shift), bool(fading)
        self.bins = self.size // 2 + 1
        self.window = taps
        handle = C.c_void_p()
        _lib.check(lib.sep_plan_create(C.byref(handle), self.size, self.shift,
                                       taps.ctypes.data_as(C.POINTER(C.c_double)), int(self.fading)),
                   "sep_plan_create")
        self._handle = handle
        self._lib = lib

    @property
    def handle(self):
        return self._handle

    def frames(self, n_samples):
        out = C.c_int()
        _lib.check(self._lib.sep_plan_frames(self._handle, int(n_samples), C.byref(out)))
        return out.value

    def istft_samples(self, frames):
        out = C.c_int64()
        _lib.check(self._lib.sep_plan_istft_samples(self._handle, int(frames), C.byref(out)))
        return out.value

    def synthesis_window(self):
        out = np.empty(self.size, dtype=np.float64)
        _lib.check(self._lib.sep_plan_synthesis_window(
            self._handle, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.sep_plan_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


_cache = {}
_cache_lock = threading.Lock()


def _device_key():
    try:
        import sys

        torch = sys.modules.get("torch")
        if torch is not None and torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def get_plan(size, shift, window=None, fading=True, window_length=None):
    """Cached plan for (size, shift, window taps, fading) on the current device."""
    taps = evaluate_window(window, int(size), window_length)
    key = (int(size), int(shift), bool(fading), taps.tobytes(), _device_key())
    with _cache_lock:
        plan = _cache.get(key)
        if plan is None:
            plan = Plan(size, shift, taps, fading)
            _cache[key] = plan
        return plan
