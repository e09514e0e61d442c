// fused.cuh -- argument block shared by the generic and the register-resident fused kernels.
#pragma once

#include "common.cuh"

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
  const float *win_half, *syn;
  const float2 *tw_half, *tw_full;
};

template <int C>
struct FusedVals {
  static constexpr int PIT = C * C;                 // pit pair sums
  static constexpr int NV = 2 * C * C + 2 * C;      // + gram, |e|^2, |r|^2
};

// Register-resident 256/512-point specialisation (fused_fast.cu).  Sets *handled when it
// launched; otherwise the generic kernel runs.
int fused_fast_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                   double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);

}  // namespace sep
