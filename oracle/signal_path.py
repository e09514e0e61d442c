"""Numpy restatement of the reference's separation signal path (CPU oracle).

TEST INFRASTRUCTURE -- not product code (see ``oracle/__init__.py``).

Every function cites the reference lines it follows.  ``.py`` citations are
``file:line`` under the reference tree; notebook citations are raw JSON line
numbers of the ``.ipynb`` plus the 0-based cell index.

Arithmetic follows the reference's dtypes: the STFT/iSTFT chain runs in float64
(float32 audio times a float64 window), SI-SDR runs in the dtype of its inputs
(float32 for wav data), PIT-MSE is evaluated in float64 by default (TensorFlow
evaluated it in float32 with an unspecified summation order; float64 is the
tighter checker) and can be asked for float32.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

# ----------------------------------------------------------------------------
# a1  frame geometry                      parallel_stft.py:125-144
# ----------------------------------------------------------------------------


def samples_to_stft_frames(samples, size, shift):
    """Number of frames covering ``samples`` (parallel_stft.py:125-134).

    ceil((samples - size + shift) / shift), evaluated in floating point exactly
    as the reference does, returned as a Python int (the reference's ``np.int``
    no longer exists in numpy >= 1.24).
    """
    return int(np.ceil((float(samples) - size + shift) / shift))


def stft_frames_to_samples(frames, size, shift):
    """Samples spanned by ``frames`` frames (parallel_stft.py:136-144)."""
    return frames * shift + size - shift


# ----------------------------------------------------------------------------
# a2  framing                             parallel_stft.py:37-123
# ----------------------------------------------------------------------------


def segment_axis(a, length, overlap=0, axis=None, end="cut", endvalue=0):
    """Chop ``a`` along ``axis`` into frames of ``length`` overlapping by ``overlap``.

    Follows parallel_stft.py:37-123.  The reference returns a strided *view*
    where it can; values are what matter for parity, so this restatement
    gathers with an index matrix (always a copy).  ``end`` in {'cut', 'pad',
    'wrap'} treats a ragged tail as the reference does (:72-99).
    """
    a = np.asarray(a)
    if axis is None:
        a = a.reshape(-1)
        axis = 0
    if axis < 0:
        axis += a.ndim
    if overlap >= length:
        raise ValueError("frames cannot overlap by more than 100%")
    if overlap < 0 or length <= 0:
        raise ValueError("overlap must be nonnegative and length must be positive")
    hop = length - overlap
    n_in = a.shape[axis]

    if n_in < length or (n_in - length) % hop:
        if n_in > length:
            whole = (n_in - length) // hop
            rounddown = length + whole * hop
            roundup = rounddown + hop
        else:
            roundup, rounddown = length, 0
        moved = np.moveaxis(a, axis, -1)
        if end == "cut":
            moved = moved[..., :rounddown]
        elif end in ("pad", "wrap"):
            grown = np.empty(moved.shape[:-1] + (roundup,), dtype=a.dtype)
            grown[..., :n_in] = moved
            if end == "pad":
                grown[..., n_in:] = endvalue
            else:
                grown[..., n_in:] = moved[..., : roundup - n_in]
            moved = grown
        # any other `end` leaves the ragged array and trips the checks below,
        # like the reference's trailing asserts (:101-105)
        a = np.moveaxis(moved, -1, axis)

    n_in = a.shape[axis]
    if n_in == 0:
        raise ValueError(
            "Not enough data points to segment array in 'cut' mode; try 'pad' or 'wrap'"
        )
    assert n_in >= length
    assert (n_in - length) % hop == 0
    n_frames = 1 + (n_in - length) // hop
    index = hop * np.arange(n_frames)[:, None] + np.arange(length)[None, :]
    return np.take(a, index, axis=axis)


# ----------------------------------------------------------------------------
# a3  STFT                                parallel_stft.py:146-196
# ----------------------------------------------------------------------------


def make_window(window, size, window_length=None):
    """Evaluate the window callable like parallel_stft.py:183-187.

    ``window`` may also be an array of ``size`` taps (already evaluated).
    """
    if callable(window):
        if window_length is None:
            w = np.asarray(window(size), dtype=np.float64)
        else:
            w = np.asarray(window(window_length), dtype=np.float64)
            w = np.concatenate([w, np.zeros(size - window_length)])
    else:
        w = np.asarray(window, dtype=np.float64)
    if w.shape != (size,):
        raise ValueError("window must have `size` taps")
    return w


def default_window():
    """The reference default, scipy's symmetric Blackman (parallel_stft.py:147)."""
    from scipy.signal import windows

    return windows.blackman


def stft(time_signal, time_dim=None, size=1024, shift=256, window=None,
         fading=True, window_length=None):
    """STFT of an N-d real signal along ``time_dim`` (parallel_stft.py:146-196).

    fade-pad (size-shift zeros both sides, :169-173) -> tail-pad to whole frames
    (:175-180) -> window (:182-187) -> frames * window -> rfft (:189-196).
    Output: complex128, frames at ``time_dim``, bins at ``time_dim + 1``.
    """
    x = np.asarray(time_signal)
    if window is None:
        window = default_window()
    if time_dim is None:
        time_dim = int(np.argmax(x.shape))
    lead = size - shift
    if fading:
        widths = [(0, 0)] * x.ndim
        widths[time_dim] = (lead, lead)
        x = np.pad(x, widths, mode="constant")
    n_padded = x.shape[time_dim]
    frames = samples_to_stft_frames(n_padded, size, shift)
    total = stft_frames_to_samples(frames, size, shift)
    widths = [(0, 0)] * x.ndim
    widths[time_dim] = (0, total - n_padded)
    x = np.pad(x, widths, mode="constant")

    w = make_window(window, size, window_length)
    framed = segment_axis(x, size, size - shift, axis=time_dim)
    shape = [1] * framed.ndim
    shape[time_dim + 1] = size
    return np.fft.rfft(framed * w.reshape(shape), axis=time_dim + 1)


# ----------------------------------------------------------------------------
# a4  magnitude / phase / PSA labels      parallel_stft.py:262-272
# ----------------------------------------------------------------------------


def mag_phase(spec):
    """|X| and angle(X) (parallel_stft.py:262-263)."""
    return np.abs(spec), np.angle(spec)


def features(mix_spec):
    """Network input row = [|X| , angle X] (parallel_stft.py:271)."""
    m, p = mag_phase(mix_spec)
    return np.concatenate((m, p), axis=-1)


def psa_labels(mix_spec, source_specs):
    """Phase-sensitive labels |S_c| cos(angle X - angle S_c), concatenated over c.

    parallel_stft.py:272 (2 speakers); uPIT_baseline.ipynb:488-489 (cell 10).
    """
    mix_angle = np.angle(mix_spec)
    cols = [np.abs(s) * np.cos(mix_angle - np.angle(s)) for s in source_specs]
    return np.concatenate(cols, axis=-1)


# ----------------------------------------------------------------------------
# a5/a6  mask application and phase recombination
#        uPIT_baseline.ipynb:1087-1088 (cell 29), :1385-1388 (cell 41)
# ----------------------------------------------------------------------------


def apply_masks(masks, mix_mag):
    """cleaned_c = mask_c * |X| concatenated on the last axis (cell 29 :1087-1090).

    ``masks`` is [..., C, T, F]; returns [..., T, C*F].
    """
    masks = np.asarray(masks)
    n_src = masks.shape[-3]
    return np.concatenate([masks[..., c, :, :] * mix_mag for c in range(n_src)], axis=-1)


def recombine(cleaned, mix_phase):
    """spec_c = cleaned_c * exp(j * angle X) (cell 41 :1385-1388)."""
    return cleaned * np.exp(mix_phase * 1j)


# ----------------------------------------------------------------------------
# a7  biorthogonal synthesis window       uPIT_baseline.ipynb:1234-1259 (cell 38)
# ----------------------------------------------------------------------------


def biorthogonal_window(analysis_window, shift):
    """Synthesis window w[n] / q[n mod shift] / size (cell 38).

    q[m] sums w[m + k*shift]^2 for k = 0..size/shift over the taps that satisfy
    ``index + 1 < size`` (:1253) -- so the last tap (index size-1) never enters
    a sum, a quirk of the reference that is kept.
    """
    w = np.asarray(analysis_window, dtype=np.float64)
    size = len(w)
    if size % shift != 0:
        raise AssertionError("fft size must be a multiple of shift")
    hops = size // shift
    q = np.zeros(shift)
    for m in range(shift):
        for k in range(hops + 1):
            idx = m + k * shift
            if idx + 1 < size:
                q[m] += w[idx] ** 2
    return w / np.tile(q, hops) / size


# ----------------------------------------------------------------------------
# a8  iSTFT                               uPIT_baseline.ipynb:1269-1307 (cell 39)
# ----------------------------------------------------------------------------


def istft(stft_signal, size=1024, shift=256, window=None, fading=True,
          window_length=None):
    """Inverse STFT by per-frame irfft, synthesis window, overlap-add (cell 39).

    The per-frame Python loop (:1300-1301) is kept on purpose: it is how the
    reference spends its time, and this function doubles as the timed CPU
    baseline.
    """
    spec = np.asarray(stft_signal)
    assert spec.shape[1] == size // 2 + 1
    if window is None:
        window = default_window()
    w = make_window(window, size, window_length)
    synth = biorthogonal_window(w, shift) * size          # :1294-1297
    n_frames = spec.shape[0]
    out = np.zeros(n_frames * shift + size - shift)
    for j in range(n_frames):                              # :1300
        start = j * shift
        out[start:start + size] += synth * np.real(np.fft.irfft(spec[j]))
    if fading:                                             # :1304-1305
        lead = size - shift
        out = out[lead:len(out) - lead]
    return out


# ----------------------------------------------------------------------------
# a9  utterance-level PIT-MSE             uPIT_baseline.ipynb:1023-1059 (cell 28)
#                                         copy: Raw_with_Convlayer.ipynb:338-374
# ----------------------------------------------------------------------------


def permutations(n_src):
    """Lexicographic permutations; perm[c] = estimate assigned to label c."""
    return list(itertools.permutations(range(n_src)))


def pit_mse(y_true, y_pred, output_size, dtype=np.float64):
    """PIT-MSE with length masking (cell 28 :1024-1057).

    y_true [B, T+1, C*F]: labels plus one extra time row holding the valid
    length; y_pred [B, T, C*F].  Predictions are masked by t < int(length)
    (:1031-1046), labels are not.  D[b,i,j] = sum_{t,f} (pred_i*m - label_j)^2.
    cost_perm = sum_c D[perm[c], c] / length.  C = 2 gives the reference's
    cost1 (perm (0,1)) and cost2 (perm (1,0)); idx = cost1 > cost2 (:1054), so a
    tie keeps perm 0.  For C > 2 (not in the reference) the first minimum in
    lexicographic order is selected (strict <), which reduces to the same rule.

    Returns dict(pair=D [B,C,C], costs [B,P], idx [B] int, loss scalar = sum_b
    min cost (:1055, a SUM over the batch, not a mean)).
    """
    yt = np.asarray(y_true, dtype=dtype)
    yp = np.asarray(y_pred, dtype=dtype)
    n_batch, rows, width = yt.shape
    n_time = rows - 1
    feat = int(output_size)
    n_src = width // feat
    assert width == n_src * feat and yp.shape == (n_batch, n_time, width)
    lengths = yt[:, n_time, 0]
    valid = (np.arange(n_time)[None, :] < lengths.astype(np.int64)[:, None]).astype(dtype)
    labels = yt[:, :n_time, :].reshape(n_batch, n_time, n_src, feat)
    preds = yp.reshape(n_batch, n_time, n_src, feat) * valid[:, :, None, None]
    pair = np.empty((n_batch, n_src, n_src), dtype=dtype)
    for i in range(n_src):
        for j in range(n_src):
            diff = preds[:, :, i, :] - labels[:, :, j, :]
            # reduce over time first, then over bins, like :1049-1052
            pair[:, i, j] = np.sum(np.sum(diff * diff, axis=1), axis=1)
    perms = permutations(n_src)
    costs = np.empty((n_batch, len(perms)), dtype=dtype)
    for p, perm in enumerate(perms):
        acc = pair[:, perm[0], 0].copy()
        for c in range(1, n_src):
            acc = acc + pair[:, perm[c], c]
        costs[:, p] = acc / lengths
    idx = np.zeros(n_batch, dtype=np.int64)
    best = costs[:, 0].copy()
    for p in range(1, len(perms)):
        better = costs[:, p] < best
        idx[better] = p
        best[better] = costs[better, p]
    return {"pair": pair, "costs": costs, "idx": idx, "loss": dtype(np.sum(best)),
            "lengths": lengths}


def pit_with_outputsize(output_size):
    """Closure factory with the reference's names (cell 28 :1023, :1059)."""

    def pit_loss(y_true, y_pred):
        return pit_mse(y_true, y_pred, output_size)["loss"]

    return pit_loss


def pit_mse_grad(y_true, y_pred, output_size, dtype=np.float64):
    """d loss / d y_pred of :func:`pit_mse` (not in the reference; TF autodiff
    would produce it): 2*m*(m*p_i - l_{c: perm[c]=i}) / length."""
    res = pit_mse(y_true, y_pred, output_size, dtype)
    yt = np.asarray(y_true, dtype=dtype)
    yp = np.asarray(y_pred, dtype=dtype)
    n_batch, rows, width = yt.shape
    n_time, feat = rows - 1, int(output_size)
    n_src = width // feat
    perms = permutations(n_src)
    lengths = res["lengths"]
    valid = (np.arange(n_time)[None, :] < lengths.astype(np.int64)[:, None]).astype(dtype)
    grad = np.zeros_like(yp)
    for b in range(n_batch):
        perm = perms[int(res["idx"][b])]
        for c in range(n_src):
            i = perm[c]
            p = yp[b, :, i * feat:(i + 1) * feat] * valid[b][:, None]
            lab = yt[b, :n_time, c * feat:(c + 1) * feat]
            grad[b, :, i * feat:(i + 1) * feat] = 2.0 * valid[b][:, None] * (p - lab) / lengths[b]
    return grad


# ----------------------------------------------------------------------------
# a10-a12  SI-SDR scoring                 metrics/evaluate_metrics.py:14-55
# ----------------------------------------------------------------------------


def pow_np_norm(signal):
    """Squared 2-norm (metrics/evaluate_metrics.py:14-17)."""
    return np.square(np.linalg.norm(signal, ord=2))


def pow_norm(s1, s2):
    """Inner product (metrics/evaluate_metrics.py:19-20)."""
    return np.sum(s1 * s2)


def si_sdr(original, estimated):
    """Scale-invariant SDR in dB (metrics/evaluate_metrics.py:22-26), computed in
    the dtype of the inputs (float32 for wav data), no epsilon."""
    target = pow_norm(estimated, original) * original / pow_np_norm(original)
    noise = estimated - target
    return 10 * np.log10(pow_np_norm(target) / pow_np_norm(noise))


def permute_si_sdr(ref1, ref2, est1, est2):
    """Best of the two speaker assignments, halved (evaluate_metrics.py:28-34).
    ``sdr1 > sdr2`` picks perm 0; a tie or NaN falls to perm 1."""
    return permute_si_sdr_detail(ref1, ref2, est1, est2)[0]


def permute_si_sdr_detail(ref1, ref2, est1, est2):
    """As :func:`permute_si_sdr`, also returning (perm, sum_perm0, sum_perm1)."""
    straight = si_sdr(ref1, est1) + si_sdr(ref2, est2)
    crossed = si_sdr(ref1, est2) + si_sdr(ref2, est1)
    if straight > crossed:
        return straight * 0.5, 0, straight, crossed
    return crossed * 0.5, 1, straight, crossed


def truncate_to_min_len(ref_s1, ref_s2, est_s1, est_s2):
    """Cut all four signals to min(len(ref_s1), len(est_s1)) (evaluate_metrics.py:46-48)."""
    n = min(np.size(ref_s1), np.size(est_s1))
    return ref_s1[:n], ref_s2[:n], est_s1[:n], est_s2[:n]


def eval_si_sdr_arrays(utterances):
    """Dataset mean of permute_si_sdr over (ref_s1, ref_s2, est_s1, est_s2) tuples,
    i.e. evaluate_metrics.py:36-55 with the wav reading factored out."""
    values = []
    for quad in utterances:
        values.append(permute_si_sdr(*truncate_to_min_len(*quad)))
    return np.mean(np.array(values)), values


# ----------------------------------------------------------------------------
# a13  SDR                                metrics/evaluate_metrics.py:57-92
#      PARITY UNPINNED: the reference delegates to museval.metrics.bss_eval
#      (third party, un-vendored, no version pin, not installed here).
# ----------------------------------------------------------------------------


def image_sdr(reference, estimated):
    """10 log10(|s|^2 / |s_hat - s|^2) in float64.

    With ``bsseval_sources_version=False`` (museval's default, which
    evaluate_metrics.py:79-81 uses) bss_eval's SDR has the distortion
    e_spat + e_interf + e_artif = s_hat - s, so the SDR value does not depend
    on the 512-tap projection filters.  Restated from the published BSS Eval v4
    definition; not checkable against museval in this environment.
    """
    s = np.asarray(reference, dtype=np.float64)
    e = np.asarray(estimated, dtype=np.float64)
    return 10 * np.log10(np.sum(s * s) / np.sum((e - s) ** 2))


def permute_sdr_detail(ref1, ref2, est1, est2):
    """Mean image-SDR over the two sources for the better assignment.

    museval picks the permutation by mean SIR (needs the filter projections);
    this restatement picks by mean SDR and says so -- parity unpinned.
    NaN handling follows evaluate_metrics.py:83-86.
    """
    def _mean(pair):
        pair = np.array(pair)
        m = np.mean(pair)
        if np.isnan(m):
            m = np.mean(np.nan_to_num(pair))
        return m

    straight = _mean([image_sdr(ref1, est1), image_sdr(ref2, est2)])
    crossed = _mean([image_sdr(ref1, est2), image_sdr(ref2, est1)])
    if crossed > straight:
        return crossed, 1, straight, crossed
    return straight, 0, straight, crossed


def eval_sdr_arrays(utterances):
    """Dataset mean of :func:`permute_sdr_detail` (evaluate_metrics.py:57-92)."""
    values = [permute_sdr_detail(*truncate_to_min_len(*quad))[0] for quad in utterances]
    return np.mean(np.array(values)), values


# ----------------------------------------------------------------------------
# a14  Conv1D "learned filterbank"        Raw_with_Convlayer.ipynb:83-100, :389
# ----------------------------------------------------------------------------


def segment_raw(wave, seg_len=40):
    """Non-overlapping ``seg_len``-sample rows, zero padded (cell 2 :83-96)."""
    wave = np.asarray(wave)
    k = int(np.ceil(len(wave) / seg_len))
    padded = np.concatenate([wave, np.zeros(k * seg_len - len(wave), dtype=wave.dtype)])
    return padded.reshape(k, seg_len)


def _activation(name):
    if name in (None, "linear"):
        return lambda v: v
    if name == "sigmoid":
        return lambda v: 1.0 / (1.0 + np.exp(-v))
    if name == "relu":
        return lambda v: np.maximum(v, 0.0)
    raise ValueError("unknown activation %r" % (name,))


def conv1d(x, kernel, bias=None, stride=1, padding="same", activation=None,
           dtype=np.float64):
    """Keras ``Conv1D`` forward on [B, K, C_in] with kernel [k, C_in, N].

    Raw_with_Convlayer.ipynb:389 uses filters=129, kernel_size=2, sigmoid,
    padding='same' on [B, K, 40].  TensorFlow 'same' padding: out = ceil(K/s),
    total = max((out-1)*s + k - K, 0), left = total // 2 (so k=2, s=1 pads only
    on the right, by one row).
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(kernel, dtype=dtype)
    n_batch, n_rows, c_in = x.shape
    taps, c_in_w, n_filt = w.shape
    assert c_in == c_in_w
    if padding == "same":
        n_out = -(-n_rows // stride)
        total = max((n_out - 1) * stride + taps - n_rows, 0)
        left = total // 2
        right = total - left
    elif padding == "valid":
        n_out = (n_rows - taps) // stride + 1
        left = right = 0
    else:
        raise ValueError(padding)
    xp = np.pad(x, ((0, 0), (left, right), (0, 0)))
    out = np.zeros((n_batch, n_out, n_filt), dtype=dtype)
    for j in range(taps):
        rows = xp[:, j:j + (n_out - 1) * stride + 1:stride, :]
        out += rows @ w[j]
    if bias is not None:
        out += np.asarray(bias, dtype=dtype)
    return _activation(activation)(out)


def filterbank_separate(wave, enc, dec, masks, stride, dtype=np.float64):
    """BASELINE config 5 (a generalisation; the reference has no decoder).

    frames[k] = wave[k*stride : k*stride+L];  code = relu(frames @ enc) [K, N];
    est_c = overlap_add((code * mask_c) @ dec, stride)  with enc [L, N],
    dec [N, L], masks [C, K, N].  K = (len - L)//stride + 1.
    """
    wave = np.asarray(wave, dtype=dtype)
    enc = np.asarray(enc, dtype=dtype)
    dec = np.asarray(dec, dtype=dtype)
    masks = np.asarray(masks, dtype=dtype)
    seg = enc.shape[0]
    k_frames = (len(wave) - seg) // stride + 1
    frames = segment_axis(wave[: (k_frames - 1) * stride + seg], seg, seg - stride)
    code = np.maximum(frames @ enc, 0.0)
    outs = []
    for c in range(masks.shape[0]):
        rows = (code * masks[c]) @ dec
        y = np.zeros((k_frames - 1) * stride + seg, dtype=dtype)
        for k in range(k_frames):
            y[k * stride:k * stride + seg] += rows[k]
        outs.append(y)
    return code, np.stack(outs)


# ----------------------------------------------------------------------------
# The whole hot path for one utterance, chained the way the reference chains
# it.  Used as the checker of the fused CUDA kernel
# and as the timed CPU baseline.
# ----------------------------------------------------------------------------


def separate_and_score(mix, refs, masks, size=256, shift=128, window=None,
                       length=None):
    """mix [N] f32, refs [C, N] f32, masks [C, T, F] -> dict.

    stft(mix), stft(ref_c) -> |X|, angle X, PSA labels (3.1) -> cleaned_c =
    mask_c |X| (cell 29) -> spec_c = cleaned_c e^{j angle X} (cell 41) ->
    est_c = istft(spec_c)[:N] -> PIT-MSE on (labels+length row, cleaned) (cell
    28) -> SI-SDR for every (ref, est) pairing (evaluate_metrics.py:22-34).
    """
    if window is None:
        window = default_window()
    mix = np.asarray(mix)
    refs = np.asarray(refs)
    n_src = refs.shape[0]
    spec = stft(mix, time_dim=0, size=size, shift=shift, window=window)
    mag, phase = mag_phase(spec)
    src_specs = [stft(refs[c], time_dim=0, size=size, shift=shift, window=window)
                 for c in range(n_src)]
    labels = psa_labels(spec, src_specs)
    n_frames, n_bins = mag.shape
    cleaned = apply_masks(masks, mag)
    ests = []
    for c in range(n_src):
        spec_c = recombine(cleaned[:, c * n_bins:(c + 1) * n_bins], phase)
        ests.append(istft(spec_c, size=size, shift=shift, window=window))
    ests = np.stack(ests)
    if length is None:
        length = n_frames
    y_true = np.concatenate([labels, np.full((1, labels.shape[1]), float(length))], axis=0)
    pit = pit_mse(y_true[None], cleaned[None], n_bins)
    n = mix.shape[0]
    est32 = ests[:, :n].astype(np.float32)
    ref32 = refs.astype(np.float32)
    si = np.array([[si_sdr(ref32[j], est32[i]) for j in range(n_src)]
                   for i in range(n_src)])
    return {"spec": spec, "labels": labels, "cleaned": cleaned, "ests": ests,
            "pit": pit, "si_sdr_pair": si}


# ----------------------------------------------------------------- 8f rank 3: sample formats
def pcm16_to_float32(pcm):
    """int16 PCM -> float32 as sf.read(dtype='float32') / librosa.load give it for 16-bit
    wavs (metrics/evaluate_metrics.py:7-12, parallel_stft.py:213)."""
    return (np.asarray(pcm, dtype=np.int16).astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def audiowrite_int16(data, normalize=False):
    """The conversion inside audiowrite (uPIT_baseline.ipynb:1331-1346, cell 40), without the
    file write: returns (int16 samples, number of clipped samples).  `np.float` in the
    reference is the removed alias of float (only reached for integer input)."""
    data = np.array(data, copy=True)
    int16_max = np.iinfo(np.int16).max
    int16_min = np.iinfo(np.int16).min
    if normalize:
        if not data.dtype.kind == 'f':
            data = data.astype(float)
        data /= np.max(np.abs(data))
    if data.dtype.kind == 'f':
        data *= int16_max
    sample_to_clip = int(np.sum(data > int16_max))
    data = np.clip(data, int16_min, int16_max)
    return data.astype(np.int16), sample_to_clip


# ----------------------------------------------------------------- 8f rank 4: TF-style SI-SDR
def sisdr_values(y_true, y_pred):
    """SiSdr.update_state's `values` (vq-vae_for_1d_data.ipynb cell 13 :398-421), float32 like TF:
    y_true [B, L + 1, 1] (last row = length), y_pred [B, L', 1]; cut to the shorter of L, L'."""
    y_true = np.asarray(y_true, dtype=np.float32)
    y_pred = np.asarray(y_pred, dtype=np.float32)
    labels = y_true[:, :-1, :]
    label_size, pred_size = labels.shape[1], y_pred.shape[1]
    if label_size < pred_size:
        y_pred = y_pred[:, :label_size, :]
    elif label_size > pred_size:
        labels = labels[:, :pred_size, :]
    dot = np.matmul(np.transpose(y_pred, (0, 2, 1)), labels)                    # [B, 1, 1]
    target = dot * labels / np.square(np.linalg.norm(labels, axis=1))[..., None]
    noise = y_pred - target
    values = 10 * np.log10(np.square(np.linalg.norm(target, axis=1)) / np.square(np.linalg.norm(noise, axis=1)))
    return values.reshape(-1)


def custom_sisdr_loss(y_true, y_pred):
    """cell 14 :457-469.  (The loss does not cut the sequences: equal lengths expected.)"""
    return -float(np.mean(sisdr_values(y_true, y_pred)))
