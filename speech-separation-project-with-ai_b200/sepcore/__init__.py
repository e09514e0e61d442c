"""sepcore -- B200-native (sm_100a) separation signal path.

Host side of libsepcore.so (C ABI in include/sepcore.h): the reference's Python
signatures (jsjs4013/Speech-Separation-Project-with-AI) over hand-written CUDA
kernels.  No CPU fallback: importing the compute modules loads the shared
library and raises ImportError if it has not been built.
"""
from . import _lib
from ._lib import SepcoreError, SepcoreUnsupported, launch_count
from .plan import Plan, get_plan
from .signal_path import (_biorthogonal_window_loopy, _samples_to_stft_frames,
                          _stft_frames_to_samples, istft, recombine_istft, segment_axis, stft,
                          stft_features)
from .losses import pit_mse, pit_with_outputsize
from .scoring import (permute_si_sdr, pow_norm, pow_np_norm, score_batch, score_flat_device, si_sdr,
                      truncate_to_min_len)
from .filterbank import conv1d, filterbank_separate, segment_raw
from .fused import parse_scores, score_layout, separate_and_score, workspace_bytes
from .bss import bss_eval, bss_eval_batch
from .records import SequenceExample, TFRecordWriter, gen_feats_record, make_sequence_example
from .audio_io import audiowrite, audiowrite_int16, pcm16_to_float32
from .tf_metrics import SiSdr, custom_sisdr_loss, sisdr_values
from .graphs import GraphedSeparator
from .pipeline import HostPipeline
from . import distributed

__all__ = [
    "SepcoreError", "SepcoreUnsupported", "launch_count", "Plan", "get_plan",
    "segment_axis", "_samples_to_stft_frames", "_stft_frames_to_samples", "stft", "stft_features",
    "_biorthogonal_window_loopy", "istft", "recombine_istft",
    "pit_mse", "pit_with_outputsize",
    "pow_np_norm", "pow_norm", "si_sdr", "permute_si_sdr", "score_batch", "score_flat_device",
    "truncate_to_min_len", "conv1d", "segment_raw", "filterbank_separate",
    "separate_and_score", "score_layout", "parse_scores", "workspace_bytes", "GraphedSeparator", "HostPipeline",
    "audiowrite", "audiowrite_int16", "pcm16_to_float32", "SiSdr", "custom_sisdr_loss", "sisdr_values",
    "distributed", "bss_eval", "bss_eval_batch",
    "SequenceExample", "TFRecordWriter", "gen_feats_record", "make_sequence_example",
]
