"""Drop-in for the signal / loss functions defined in `uPIT_baseline.ipynb`:
istft + _biorthogonal_window_loopy (cells 38-39, :1234-1307), pit_with_outputsize
(cell 28, :1023-1059), plus the batched mask/phase recombination of cell 41."""
from sepcore.losses import pit_with_outputsize  # noqa: F401
from sepcore.signal_path import (_biorthogonal_window_loopy, istft, recombine_istft,  # noqa: F401
                                 segment_axis, stft)
