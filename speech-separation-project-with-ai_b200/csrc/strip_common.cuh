// strip_common.cuh -- helpers shared by the warp-autonomous strip kernels
// (fused_strip.cu: size 256; fused_strip512.cu: size 512).
#pragma once

#include <algorithm>

#include "common.cuh"
#include "fused.cuh"

namespace sep {

constexpr int kWT = 16 * 18 + 8;                 // transposed window table (size 256), floats

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp-level "last strip of an utterance finalises it" (single-launch mode); same
// arithmetic and order as fused_finalize_kernel / fused_sums_kernel.
template <int C>
__device__ __forceinline__ void finalize_by_warp(const FusedArgs &a, int b, int lane) {
  constexpr int NV = FusedVals<C>::NV;
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  constexpr int STRIDE = 3 * C * C + P + 6;
  int last = 0;
  if (lane == 0) {
    __threadfence();                                        // publish this strip's partial row
    last = atomicAdd(a.counters + b, 1) == a.tiles - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.0;
  for (int t = lane; t < a.tiles; t += 32) {
    const double *src = a.partials + (static_cast<int64_t>(b) * a.tiles + t) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldcg(src + i);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  int last_utt = 0;
  if (lane == 0) {
    double *row = a.scores + static_cast<int64_t>(b) * STRIDE;
    const double len = a.lengths ? static_cast<double>(a.lengths[b]) : static_cast<double>(a.T);
    finalize_pit<C>(v, len, row);
    finalize_scores<C>(v + C * C, v + 2 * C * C, v + 2 * C * C + C, row + C * C + P + 2);
    a.counters[b] = 0;
    __threadfence();                                        // publish the score row
    last_utt = atomicAdd(a.counters + a.batch, 1) == a.batch - 1;
  }
  last_utt = __shfl_sync(0xffffffffu, last_utt, 0);
  if (!last_utt) return;
  __threadfence();
  if (a.sums) {
    constexpr int off_pit = C * C + P + 1, off_si = C * C + P + 2 + C * C, off_sdr = off_si + 2 + C * C;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int u = lane; u < a.batch; u += 32) {
      const double *row = a.scores + static_cast<int64_t>(u) * STRIDE;
      s0 += __ldcg(row + off_pit);
      s1 += __ldcg(row + off_si);
      s2 += __ldcg(row + off_sdr);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { a.sums[0] = s0; a.sums[1] = s1; a.sums[2] = s2; a.sums[3] = a.batch; }
    push_sums(a, lane, s0, s1, s2, a.batch);
  }
  if (lane == 0) a.counters[a.batch] = 0;
}

// Strips per utterance: minimise the iterations of the busiest warp (rounds of strips x
// longest strip); ties go to fewer strips (fewer recomputed halo frames).
// fpi = frames per warp iteration (4: two half-warps x a frame pair; 2: two half-warps x one frame).
static inline void pick_strips(int T, int H, int fpi, int batch, int warps_total, int *strips, int *iters) {
  int best_s = 1, best_i = (T + H + fpi - 1) / fpi;
  int64_t best_cost = -1;
  const int smax = std::max(1, std::min(T / fpi, 256));
  for (int s = 1; s <= smax; ++s) {
    const int it = (T + H * s + fpi - 1) / fpi;
    const int nmin = (H + fpi) / fpi;                 // a strip must own at least one hop block
    if (it < s * nmin) break;
    const int64_t nmax = (it + s - 1) / s;
    const int64_t rounds = (static_cast<int64_t>(batch) * s + warps_total - 1) / warps_total;
    const int64_t cost = rounds * nmax;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; best_i = it; }
  }
  *strips = best_s;
  *iters = best_i;
}

}  // namespace sep
