"""ctypes binding of libsepcore.so (the C ABI declared in include/sepcore.h).

There is no CPU fallback: if the shared library is missing or a call fails the
error is raised, never papered over.
"""
from __future__ import annotations

import ctypes as C
import os

MEM_HOST, MEM_DEVICE = 0, 1
ACT = {None: 0, "linear": 0, "sigmoid": 1, "relu": 2}
PAD = {"valid": 0, "same": 1}
MAX_SOURCES = 4

ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = -1, -2, -3, -4


class SepcoreError(RuntimeError):
    """A libsepcore call failed (CUDA error, out of memory)."""


class SepcoreUnsupported(NotImplementedError):
    """Legal in the reference, not built in the CUDA library (e.g. non power-of-two size)."""


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "SEPCORE_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "libsepcore.so"))

_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
_i32p, _i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
_vp, _int, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes); mirrors include/sepcore.h one to one
PROTOTYPES = {
    "sep_version": (_int, []),
    "sep_last_error": (C.c_char_p, []),
    "sep_launch_count": (_i64, []),
    "sep_last_kernel": (C.c_char_p, []),
    "sep_record_masked_crc": (C.c_uint32, [C.c_char_p, _i64]),
    "sep_record_size": (_int, [_int, _int, _int, _int, C.POINTER(_i64)]),
    "sep_record_encode": (_int, [_vp, _vp, _int, _int, _int, C.c_float, C.c_char_p, _int, _i32p, _vp, _i64,
                                 C.POINTER(_i64)]),
    "sep_bss_eval_row_width": (_int, [_int]),
    "sep_bss_eval_f32": (_int, [_vp, _vp, _i64p, _i64p, _i64p, _int, _int, _i64, _i64, _int, _vp, _int, _vp]),
    "sep_profile_enable": (_int, [_int]),
    "sep_profile_collect": (_int, [_f64p, C.POINTER(_int)]),
    "sep_plan_create": (_int, [C.POINTER(_vp), _int, _int, _f64p, _int]),
    "sep_plan_destroy": (_int, [_vp]),
    "sep_plan_frames": (_int, [_vp, _i64, C.POINTER(_int)]),
    "sep_plan_istft_samples": (_int, [_vp, _int, C.POINTER(_i64)]),
    "sep_plan_synthesis_window": (_int, [_vp, _f64p]),
    "sep_score_stride": (_int, [_int]),
    "sep_segment_axis_f32": (_int, [_vp, _int, _i64, _int, _int, _vp, _int, _vp]),
    "sep_stft_f32": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _int, _vp]),
    "sep_stft_features_f32": (_int, [_vp, _vp, _vp, _int, _int, _i64, _vp, _vp, _int, _vp]),
    "sep_istft_f32": (_int, [_vp, _vp, _int, _int, _vp, _int, _vp]),
    "sep_recombine_istft_f32": (_int, [_vp, _vp, _vp, _int, _int, _int, _vp, _int, _vp]),
    "sep_fused_separate_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _i64, _vp, _vp,
                                      _vp, _int, _vp]),
    "sep_fused_workspace_bytes": (_int, [_vp, _int, _int, _i64, C.POINTER(_i64)]),
    "sep_fused_separate_ws_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _i64, _vp, _vp,
                                         _vp, _vp, _i64, _int, _vp]),
    "sep_fused_separate_push_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _i64, _vp, _vp,
                                           _vp, _vp, _i64, _vp, _int, _int, _int, _int, _vp]),
    "sep_peer_alloc": (_int, [_i64, C.POINTER(_vp), C.c_char_p]),
    "sep_peer_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "sep_peer_close": (_int, [_vp]),
    "sep_peer_free": (_int, [_vp]),
    "sep_pit_mse_f32": (_int, [_vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "sep_score_batch_f32": (_int, [_vp, _vp, _i64p, _i64p, _i64p, _int, _int, _i64, _i64, _vp, _vp,
                                   _int, _vp]),
    "sep_dot_f32": (_int, [_vp, _vp, _i64, _vp, _int, _vp]),
    "sep_filterbank_separate_f32": (_int, [_vp, _vp, _vp, _vp, _int, _int, _i64, _int, _int, _int, _vp,
                                           _vp, _int, _vp]),
    "sep_conv1d_f32": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _vp,
                              _int, _vp]),
    "sep_pcm16_to_f32": (_int, [_vp, _i64, _vp, _int, _vp]),
    "sep_audiowrite_i16_f32": (_int, [_vp, _int, _i64, _int, _vp, _vp, _int, _vp]),
    "sep_audiowrite_i16_f64": (_int, [_vp, _int, _i64, _int, _vp, _vp, _int, _vp]),
}

_lib = None


def load():
    """Loads libsepcore.so once; raises ImportError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libsepcore.so not found at %s -- build it with `make -C %s` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. sepcore has no CPU fallback."
            % (LIB_PATH, os.path.dirname(LIB_PATH)))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == 0:
        return
    msg = load().sep_last_error().decode(errors="replace")
    text = "%s: %s (code %d)" % (what, msg, rc) if what else "%s (code %d)" % (msg, rc)
    if rc == ERR_UNSUPPORTED:
        raise SepcoreUnsupported(text)
    if rc == ERR_INVALID:
        raise ValueError(text)
    if rc == ERR_NOMEM:
        raise MemoryError(text)
    raise SepcoreError(text)


def launch_count():
    return int(load().sep_launch_count())


def last_kernel():
    """Name of the dominant kernel launched by the last library call on this thread."""
    return load().sep_last_kernel().decode(errors="replace")


def profile_enable(on=True):
    check(load().sep_profile_enable(1 if on else 0), "sep_profile_enable")


def profile_collect():
    """(total kernel ms, bracketed launches) since the last call; resets the log."""
    total, count = C.c_double(), C.c_int()
    check(load().sep_profile_collect(C.byref(total), C.byref(count)), "sep_profile_collect")
    return total.value, count.value
