// score.cuh -- per-utterance finalisation shared by the fused path and the
// scoring kernel: PIT permutation search (uPIT_baseline.ipynb:1049-1055, cell
// 28), SI-SDR with permutation (metrics/evaluate_metrics.py:22-34) and image
// SDR (evaluate_metrics.py:57-92 via museval -- parity unpinned).
#pragma once

#include "common.cuh"

namespace sep {

// 10 log10(ratio).  The ratio is formed in float64 (that is where cancellation
// would hurt); the logarithm itself only needs float32: a relative rounding of
// 6e-8 in the argument moves the result by 2.6e-7 dB, and log10f is good to an ulp,
// while a float64 log10 costs hundreds of instructions on the serial tail of the
// kernel.  Ratios beyond the float range (> 380 dB) saturate to +-inf like the
// reference's float32 arithmetic does much earlier.
__device__ __forceinline__ double db10(double ratio) {
  return 10.0 * static_cast<double>(log10f(static_cast<float>(ratio)));
}

// Gram statistics of one utterance, float64:
//   g[i][j] = <est_i, ref_j>,  ee[i] = |est_i|^2,  er[j] = |ref_j|^2
// Writes si_pair[C*C], si_best, si_perm, sdr_pair[C*C], sdr_best, sdr_perm
// (2*C*C + 4 doubles) to `out`.
template <int C>
__device__ void finalize_scores(const double *g, const double *ee, const double *er, double *out) {
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  double si[C * C], sd[C * C];
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) {
      const double dot = g[i * C + j];
      // target = <e,o> o / |o|^2 ; noise = e - target  (evaluate_metrics.py:23-24)
      const double tgt = dot * dot / er[j];
      double noise = ee[i] - tgt;
      if (noise < 0.0) noise = 0.0;
      si[i * C + j] = db10(tgt / noise);
      double dist = ee[i] - 2.0 * dot + er[j];
      if (dist < 0.0) dist = 0.0;
      sd[i * C + j] = db10(er[j] / dist);
    }
  // SI-SDR: `if sdr1 > sdr2` keeps the earlier permutation only on a strict
  // win; a tie or NaN moves on to the later one (evaluate_metrics.py:31-34).
  int perm[SEP_MAX_SOURCES];
  int si_perm = 0, sd_perm = 0;
  double si_best = 0.0, sd_best = 0.0;
  for (int p = 0; p < P; ++p) {
    nth_permutation(C, p, perm);
    double a = 0.0, m = 0.0, m0 = 0.0;
    for (int c = 0; c < C; ++c) {
      a += si[perm[c] * C + c];
      const double v = sd[perm[c] * C + c];
      m += v;
      // np.nan_to_num: nan -> 0, +-inf -> +-DBL_MAX  (evaluate_metrics.py:85-86)
      m0 += isnan(v) ? 0.0 : (isinf(v) ? copysign(1.7976931348623157e308, v) : v);
    }
    m /= C;
    if (isnan(m)) m = m0 / C;
    if (p == 0 || !(si_best > a)) { si_best = a; si_perm = p; }
    if (p == 0 || m > sd_best) { sd_best = m; sd_perm = p; }
  }
  for (int i = 0; i < C * C; ++i) out[i] = si[i];
  out[C * C] = si_best / C;
  out[C * C + 1] = si_perm;
  for (int i = 0; i < C * C; ++i) out[C * C + 2 + i] = sd[i];
  out[2 * C * C + 2] = sd_best;
  out[2 * C * C + 3] = sd_perm;
}

// PIT-MSE finalisation: pair[i][j] = sum (m * pred_i - label_j)^2, length = valid frames.
// Writes pair[C*C], costs[P], perm, loss (C*C + P + 2 doubles) to `out`.
template <int C>
__device__ void finalize_pit(const double *pair, double length, double *out) {
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  int perm[SEP_MAX_SOURCES];
  int best = 0;
  double best_cost = 0.0;
  for (int i = 0; i < C * C; ++i) out[i] = pair[i];
  for (int p = 0; p < P; ++p) {
    nth_permutation(C, p, perm);
    double acc = 0.0;
    for (int c = 0; c < C; ++c) acc += pair[perm[c] * C + c];   // cost1 / cost2, :1049-1052
    acc /= length;
    out[C * C + p] = acc;
    if (p == 0 || acc < best_cost) { best_cost = acc; best = p; }  // idx = cost1 > cost2, :1054
  }
  out[C * C + P] = best;
  out[C * C + P + 1] = best_cost;
}

}  // namespace sep
