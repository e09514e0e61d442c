"""CPU restatement of ``museval.metrics.bss_eval`` (BSS Eval v4) for the reference's call site
``metrics/evaluate_metrics.py:79-81``:

    museval.metrics.bss_eval(reference_stack, estimated_stack, window=np.inf, hop=np.inf,
                             compute_permutation=True)          # filters_len=512, framewise_filters=False,
                                                                # bsseval_sources_version=False (defaults)

TEST INFRASTRUCTURE.  **PARITY UNPINNED**: museval is a third-party dependency of the reference, it is
not vendored under /root/reference, the reference has no requirements file (era: museval 0.3.x / 0.4.0,
Python 3.7) and it is not installed here, and no reference test or committed output pins its result
(SURVEY.md 8c).  This module restates the published algorithm -- E. Vincent et al., "First stereo audio
source separation evaluation campaign" (2007) / BSS Eval v4 as implemented in museval.metrics and
mir_eval.separation.bss_eval_images -- function by function, in float64 like museval (its `_zeropad`
copies the float32 wav data into ``np.zeros`` arrays):

  _compute_reference_correlations   G[j, i] = Toeplitz matrices of the cross-correlations of the zero-padded
                                    references, lags 0..filters_len-1, via FFT of length 2^ceil(log2(n + L - 1))
  _compute_projection_filters       D = cross-correlations references x estimate; C = solve(G + eps I, D)
  _project                          sum_j fftconvolve(C[j], reference_j)
  _bss_decomp_mtifilt               s_true, e_spat, e_interf, e_artif
  _bss_crit (images version)        SDR, ISR, SIR, SAR with _safe_db
  bss_eval                          one window (window = hop = inf), silent source -> NaN, permutation = argmax of
                                    the mean SIR over the candidate permutations (itertools order)

The CUDA path (csrc/bss.cu) does NOT follow this recipe literally: it never forms the projections in the time
domain but evaluates their energies from the correlation tables (d^T G^-1 d identities), so agreement
between the two is a real check of both.
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.fft
from scipy.linalg import toeplitz
from scipy.signal import fftconvolve

FILTERS_LEN = 512


def _zeropad(sig, n_zeros, axis=0):
    """museval._zeropad: `n_zeros` zeros appended along `axis`; the result is float64."""
    sig = np.moveaxis(np.asarray(sig), axis, 0)
    out = np.zeros((sig.shape[0] + n_zeros,) + sig.shape[1:])
    out[:sig.shape[0], ...] = sig
    return np.moveaxis(out, 0, axis)


def _n_fft(nsampl, filters_len):
    return int(2 ** np.ceil(np.log2(nsampl + filters_len - 1.0)))


def compute_reference_correlations(reference_sources, filters_len=FILTERS_LEN):
    """reference_sources [nsrc, nsampl, nchan] -> G [nsrc, nsrc, nchan, nchan, L, L], sf [nsrc, nchan, n_fft]."""
    nsrc, nsampl, nchan = reference_sources.shape
    refs = np.moveaxis(reference_sources, 1, 2)                      # nsrc x nchan x nsampl
    refs = _zeropad(refs, filters_len - 1, axis=2)
    n_fft = _n_fft(nsampl, filters_len)
    sf = scipy.fft.fft(refs, n=n_fft, axis=2)
    G = np.zeros((nsrc, nsrc, nchan, nchan, filters_len, filters_len))
    items = list(itertools.product(range(nsrc), range(nchan)))
    for (i, c1), (j, c2) in itertools.combinations_with_replacement(items, 2):
        ssf = np.real(scipy.fft.ifft(sf[j, c2] * np.conj(sf[i, c1])))
        ss = toeplitz(np.hstack((ssf[0], ssf[-1:-filters_len:-1])), r=ssf[:filters_len])
        G[j, i, c2, c1] = ss
        G[i, j, c1, c2] = ss.T
    return G, sf


def _reshape_G(G):
    """nsrc x nsrc x nchan x nchan x L x L  ->  (nsrc nchan L) x (nsrc nchan L)."""
    G = np.moveaxis(G, (1, 3), (3, 4))
    nsrc, nchan, flen = G.shape[0:3]
    return np.reshape(G, (nsrc * nchan * flen, nsrc * nchan * flen))


def compute_projection_filters(G, sf, estimated_source):
    """Least-squares projection of one estimate [nsampl, nchan] on the delayed references (delays 0..L-1).
    G [nsrc, nsrc, nchan, nchan, L, L] -> C [nsrc, nchan, L, nchan]; G [nchan, nchan, L, L] (one reference
    source, museval's G[jtrue, jtrue]) -> C [nchan, L, nchan]."""
    eps = np.finfo(float).eps
    nsampl, nchan = estimated_source.shape
    single = G.ndim == 4
    if single:                                                       # a single reference source
        G = G[None, None, ...]
        sf = sf[None, ...]
    nsrc = G.shape[0]
    filters_len = G.shape[-1]
    est = _zeropad(estimated_source.T, filters_len - 1, axis=1)
    n_fft = _n_fft(nsampl, filters_len)
    sef = scipy.fft.fft(est, n=n_fft)
    D = np.zeros((nsrc, nchan, filters_len, nchan))
    for j, cj, c in itertools.product(range(nsrc), range(nchan), range(nchan)):
        ssef = np.real(scipy.fft.ifft(sf[j, cj] * np.conj(sef[c])))
        D[j, cj, :, c] = np.hstack((ssef[0], ssef[-1:-filters_len:-1]))
    D = D.reshape(nsrc * nchan * filters_len, nchan)
    Gm = _reshape_G(G)
    try:
        C = np.linalg.solve(Gm + eps * np.eye(Gm.shape[0]), D)
    except np.linalg.LinAlgError:
        C = np.linalg.lstsq(Gm, D, rcond=None)[0]
    C = C.reshape(nsrc, nchan, filters_len, nchan)
    return C[0] if single else C


def _project(reference_sources, C):
    """Filter the references with the projection filters C and sum: [nsampl + L - 1, nchan]."""
    if reference_sources.ndim == 2:
        reference_sources = reference_sources[None, ...]
        C = C[None, ...]
    nsrc, nsampl, nchan = reference_sources.shape
    filters_len = C.shape[-2]
    refs = _zeropad(reference_sources, filters_len - 1, axis=1)
    sproj = np.zeros((nchan, nsampl + filters_len - 1))
    for j, cj, c in itertools.product(range(nsrc), range(nchan), range(nchan)):
        sproj[c] += fftconvolve(C[j, cj, :, c], refs[j, :, cj])[:nsampl + filters_len - 1]
    return sproj.T


def bss_decomp_mtifilt(reference_sources, estimated_source, j, C, Cj):
    filters_len = Cj.shape[-2]
    s_true = _zeropad(reference_sources[j], filters_len - 1, axis=0)
    est = _zeropad(estimated_source, filters_len - 1, axis=0)
    e_spat = _project(reference_sources[j], Cj) - s_true
    e_interf = _project(reference_sources, C) - s_true - e_spat
    e_artif = -s_true - e_spat - e_interf + est
    return s_true, e_spat, e_interf, e_artif


def _safe_db(num, den):
    if den == 0:
        return np.inf
    return 10 * np.log10(num / den)


def bss_crit(s_true, e_spat, e_interf, e_artif):
    """Images version (bsseval_sources_version=False): SDR, ISR, SIR, SAR."""
    energy_s_true = np.sum(s_true ** 2)
    sdr = _safe_db(energy_s_true, np.sum((e_spat + e_interf + e_artif) ** 2))
    isr = _safe_db(energy_s_true, np.sum(e_spat ** 2))
    sir = _safe_db(np.sum((s_true + e_spat) ** 2), np.sum(e_interf ** 2))
    sar = _safe_db(np.sum((s_true + e_spat + e_interf) ** 2), np.sum(e_artif ** 2))
    return sdr, isr, sir, sar


def _any_source_silent(sources):
    """True when a source is identically zero (museval: NaN metrics for the window)."""
    return bool(np.any(np.all(np.sum(sources, axis=tuple(range(2, sources.ndim))) == 0, axis=1)))


def bss_eval(reference_sources, estimated_sources, compute_permutation=True, filters_len=FILTERS_LEN):
    """One-window BSS Eval v4 (window = hop = inf, time-invariant filters).

    reference_sources / estimated_sources: [nsrc, nsampl] or [nsrc, nsampl, nchan].
    Returns (sdr, isr, sir, sar, perm): metric arrays [nsrc, 1] for the selected assignment
    (row jtrue = reference jtrue against estimate perm[jtrue]), perm [nsrc] -- plus, as a sixth
    element, the full table s_r [4, nsrc (jtrue), nsrc (jest)] for the tests."""
    ref = np.atleast_3d(np.asarray(reference_sources))
    est = np.atleast_3d(np.asarray(estimated_sources))
    if ref.shape != est.shape:
        raise ValueError("reference and estimated sources must have the same shape")
    nsrc = est.shape[0]
    if compute_permutation:
        cand = np.array(list(itertools.permutations(range(nsrc))))
    else:
        cand = np.arange(nsrc)[None, :]
    s_r = np.full((4, nsrc, nsrc), np.nan)
    if not _any_source_silent(ref) and not _any_source_silent(est):
        G, sf = compute_reference_correlations(ref, filters_len)
        for jest in range(nsrc):
            C = compute_projection_filters(G, sf, est[jest])
            for jtrue in range(nsrc):
                if not np.any(cand[:, jtrue] == jest):
                    continue
                Cj = compute_projection_filters(G[jtrue, jtrue], sf[jtrue], est[jest])
                s_r[:, jtrue, jest] = bss_crit(*bss_decomp_mtifilt(ref, est[jest], jtrue, C, Cj))
    dum = np.arange(nsrc)
    mean_sir = np.array([np.mean(s_r[2, dum, perm]) for perm in cand])
    popt = cand[np.argmax(mean_sir)]
    sel = s_r[:, dum, popt]
    return sel[0][:, None], sel[1][:, None], sel[2][:, None], sel[3][:, None], popt, s_r


def eval_sdr_one(ref_s1, ref_s2, est_s1, est_s2):
    """evaluate_metrics.py:66-88 for one file: mean SDR of the SIR-selected assignment, with the NaN fallback."""
    reference = np.stack((np.reshape(ref_s1, (-1, 1)), np.reshape(ref_s2, (-1, 1))), axis=0)
    estimated = np.stack((np.reshape(est_s1, (-1, 1)), np.reshape(est_s2, (-1, 1))), axis=0)
    sdr, isr, sir, sar, perm, table = bss_eval(reference, estimated)
    sdr_back = sdr
    value = np.mean(sdr_back)
    if np.isnan(value):
        value = np.mean(np.nan_to_num(sdr_back))
    return float(value), perm, table


def eval_sdr_arrays(utterances):
    """Dataset mean of eval_sdr_one over (ref_s1, ref_s2, est_s1, est_s2) tuples (evaluate_metrics.py:57-92,
    wav reading factored out; truncate-to-min-length :70-72)."""
    values = []
    for ref_s1, ref_s2, est_s1, est_s2 in utterances:
        n = min(np.size(ref_s1), np.size(est_s1))
        values.append(eval_sdr_one(ref_s1[:n], ref_s2[:n], est_s1[:n], est_s2[:n])[0])
    return float(np.mean(np.array(values))), values
