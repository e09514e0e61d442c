"""CPU oracle for the separation signal path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy, the algorithms of the reference
(jsjs4013/Speech-Separation-Project-with-AI) that lie on the hot path named in
BASELINE.json: STFT framing/window/rFFT, magnitude/phase/PSA labels, mask
application, phase recombination, biorthogonal synthesis window, iSTFT
overlap-add, utterance-level PIT-MSE, SI-SDR / image-SDR scoring and the
Conv1D "learned filterbank" front end.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker or as the timed CPU
baseline.  The product (``sepcore``) never imports this package and has no CPU
fallback: it fails loudly when ``libsepcore.so`` is missing.

Parity status:
  * pinned   : framing, STFT, |X|, angle(X), PSA labels  -> committed TFRecord
               golden vectors of the reference (tests/golden/tfrecord_golden.npz)
               and outputs of the unmodified reference functions run in the
               authoring container (tests/golden/reference_run.npz).
  * pinned   : SI-SDR + 2-speaker permutation            -> reference functions
               run on the committed wsj0-2mix / test_wav pairs.
  * pinned   : iSTFT / biorthogonal window               -> executed notebook
               cells 38-39 (reference_run.npz).
  * restated : PIT-MSE, Conv1D filterbank (TensorFlow is not installable here;
               no reference test pins them) -> numpy restatement cross-checked
               against an independent torch CPU formulation.
  * UNPINNED : SDR via museval.metrics.bss_eval (third party, not vendored, no
               version pin, not installed) -> "parity unpinned".
"""
from . import signal_path  # noqa: F401
