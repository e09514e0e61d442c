"""Drop-in for the signal functions of the reference's `parallel_stft.py`
(and the byte-identical copies in `parallel_stft_single.py:39-198`): same names,
same signatures, CUDA underneath.  `from parallel_stft import stft` keeps working.

Out of scope (file / TFRecord I/O, dataset walking -- SURVEY.md section 2):
audioread, make_sequence_example, gen_feats, main.
"""
from sepcore.signal_path import (_samples_to_stft_frames, _stft_frames_to_samples,  # noqa: F401
                                 segment_axis, stft)
