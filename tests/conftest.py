import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speech-separation-project-with-ai_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def wsj0():
    """The reference's committed wsj0-2mix toy set + test_wav estimates, as float32 / 32768."""
    z = _load("wsj0_fixture.npz")
    names = [str(n) for n in z["names"]]
    utts = []
    for i, name in enumerate(names):
        f = lambda key: z[key].astype(np.float32) / 32768.0
        utts.append({
            "name": name,
            "mix": f(f"tt_mix_{i}"), "s1": f(f"tt_s1_{i}"), "s2": f(f"tt_s2_{i}"),
            "est_s1": f(f"est_s1_{i}"), "est_s2": f(f"est_s2_{i}"),
        })
    return utts


@pytest.fixture(scope="session")
def tfrecord_golden():
    return _load("tfrecord_golden.npz")


@pytest.fixture(scope="session")
def reference_run():
    return _load("reference_run.npz")


def rel_err(a, b):
    """max |a - b| / max |b|: the 1e-4 'relative' criterion of BASELINE.json is
    relative to the array's scale (per-element relative error is meaningless
    for near-zero bins)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return float(d / (n if n > 0 else 1.0))
