// fused.cuh -- argument block shared by the generic and the register-resident fused kernels.
#pragma once

#include "common.cuh"
#include "score.cuh"

namespace sep {

struct FusedArgs {
  const float *mix;       // [B, n]
  const float *masks;     // [B, C, T, F]
  const float *refs;      // [B, C, n] or null
  const float *lengths;   // [B] or null
  const int32_t *valid;   // [B] or null
  float *est;             // [B, C, n] or null
  double *partials;       // [B * tiles, NV] or null (no refs)
  int64_t n;
  int T, size, shift, pad, tb, tiles;
  int batch, lookahead;   // lookahead: tiles ahead of this CTA to pull into L2 (0 = off)
  const float *win_half, *syn;
  const float2 *tw_half, *tw_full, *tw16;
};

template <int C>
struct FusedVals {
  static constexpr int PIT = C * C;                 // pit pair sums
  static constexpr int NV = 2 * C * C + 2 * C;      // + gram, |e|^2, |r|^2
};

// One warp per utterance: sums the tile partials in tile order, then the
// permutation searches.  scores row layout: see sepcore.h.
template <int C>
__global__ void fused_finalize_kernel(const double *__restrict__ partials, int tiles,
                                      const float *__restrict__ lengths, int T,
                                      double *__restrict__ scores, int stride) {
  constexpr int NV = FusedVals<C>::NV;
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  const int b = blockIdx.x, lane = threadIdx.x;
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.0;
  for (int t = lane; t < tiles; t += 32) {
    const double *src = partials + (static_cast<int64_t>(b) * tiles + t) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += src[i];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
    double *row = scores + static_cast<int64_t>(b) * stride;
    const double len = lengths ? static_cast<double>(lengths[b]) : static_cast<double>(T);
    finalize_pit<C>(v, len, row);
    finalize_scores<C>(v + C * C, v + 2 * C * C, v + 2 * C * C + C, row + C * C + P + 2);
  }
}

// sums[4] = {sum pit_loss, sum si_best, sum sdr_best, batch}; one warp, fixed order.
static __global__ void fused_sums_kernel(const double *__restrict__ scores, int batch, int stride,
                                  int off_pit, int off_si, int off_sdr, double *__restrict__ sums) {
  const int lane = threadIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < batch; b += 32) {
    const double *row = scores + static_cast<int64_t>(b) * stride;
    s0 += row[off_pit];
    s1 += row[off_si];
    s2 += row[off_sdr];
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { sums[0] = s0; sums[1] = s1; sums[2] = s2; sums[3] = batch; }
}

template <int C>
static int launch_fused_finalize(const FusedArgs &a, int batch, double *d_scores, double *d_sums,
                                 cudaStream_t stream) {
  const int stride = sep_score_stride(C);
  fused_finalize_kernel<C><<<batch, 32, 0, stream>>>(a.partials, a.tiles, a.lengths, a.T, d_scores,
                                                     stride);
  SEP_LAUNCHED();
  if (d_sums) {
    const int P = factorial(C);
    const int off_pit = C * C + P + 1, off_si = C * C + P + 2 + C * C;
    const int off_sdr = off_si + 2 + C * C;
    fused_sums_kernel<<<1, 32, 0, stream>>>(d_scores, batch, stride, off_pit, off_si, off_sdr, d_sums);
    SEP_LAUNCHED();
  }
  return SEP_OK;
}

// Register-resident 256/512-point specialisation (fused_fast.cu).  Sets *handled when it
// launched; otherwise the generic kernel runs.
int fused_fast_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                   double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);

}  // namespace sep
