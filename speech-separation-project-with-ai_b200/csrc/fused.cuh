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
  double *scores;         // [B, stride]   (in-kernel finalisation, fast path)
  double *sums;           // [4] or null
  int *counters;          // [B + 1] zero on entry, left zero: tiles done per utterance, utterances done
  // per-batch reduction without a collective call (sep_fused_separate_push_f32): the CTA that adds the batch sums also
  // PUSHES them into every rank's inbox over NVLink -- peers[p] = rank p's inbox, [slots][world][4] float64 followed by
  // [slots] uint64 arrival counters; this rank's row of slot `push_slot`
  double *const *peers;
  int world, rank, push_slot, push_slots;
  int64_t n;
  int T, size, shift, pad, tb, tiles;
  int batch, lookahead;   // lookahead: tiles ahead of this CTA to pull into L2 (0 = off)
  int strip_iters, vec_ok;  // strip kernel: iterations per utterance; rows 16-byte aligned
  const float *win_half, *syn;
  const float *win_t, *syn_t;   // [16][18] transposed tables (size 256)
  const float2 *tw_half, *tw_full, *tw16;
  const float2 *win2_t, *syn2_t, *tw512_t;   // size 512 strip kernel tables, [16][18]
};

#ifdef __CUDACC__
// Called by one whole warp that holds the batch sums on every lane.  Lane p < world stores the four float64 values into
// rank p's inbox (peer-mapped memory: plain stores that travel over NVLink / NVSwitch), makes them visible system-wide
// and bumps that inbox's arrival counter.  Nothing ever waits here: a reader checks arrived[slot] == world after its own
// synchronisation (stream / host), then adds the `world` rows -- an all-gather by one-sided puts plus a local reduction.
__device__ __forceinline__ void push_sums(const FusedArgs &a, int lane, double s0, double s1, double s2, double n) {
  if (a.peers == nullptr || lane >= a.world) return;
  double *inbox = a.peers[lane];
  double *row = inbox + (static_cast<int64_t>(a.push_slot) * a.world + a.rank) * 4;
  row[0] = s0; row[1] = s1; row[2] = s2; row[3] = n;
  __threadfence_system();
  unsigned long long *arrived = reinterpret_cast<unsigned long long *>(inbox + static_cast<int64_t>(a.push_slots) * a.world * 4);
  atomicAdd_system(arrived + a.push_slot, 1ull);
}
#endif

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

// Last-tile-done finalisation inside the producing kernel (one launch per step).
// Called by every thread of a CTA after its partial row has been written by thread 0.
// The CTA that completes an utterance reduces that utterance's partials in tile order
// (same arithmetic as fused_finalize_kernel, so results are bit-identical); the CTA
// that completes the last utterance adds the batch sums in utterance order.  The
// counters are left at zero for the next launch.
template <int C>
__device__ __forceinline__ void finalize_in_kernel(const FusedArgs &a, int b, int *s_flag) {
  constexpr int NV = FusedVals<C>::NV;
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  constexpr int STRIDE = 3 * C * C + P + 6;
  if (threadIdx.x == 0) {
    __threadfence();                                        // publish this tile's partials
    *s_flag = atomicAdd(a.counters + b, 1) == a.tiles - 1;
  }
  __syncthreads();
  if (!*s_flag || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
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

// sums[4] = {sum pit_loss, sum si_best, sum sdr_best, batch}; one warp, fixed order.
static __global__ void fused_sums_kernel(const double *__restrict__ scores, int batch, int stride,
                                  int off_pit, int off_si, int off_sdr, double *__restrict__ sums, const FusedArgs a) {
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
  push_sums(a, lane, s0, s1, s2, batch);
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
    fused_sums_kernel<<<1, 32, 0, stream>>>(d_scores, batch, stride, off_pit, off_si, off_sdr, d_sums, a);
    SEP_LAUNCHED();
  }
  return SEP_OK;
}

// Register-resident 256/512-point specialisation (fused_fast.cu).  Sets *handled when it
// launched; otherwise the generic kernel runs.
int fused_wstrip_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                     double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);
int fused_strip_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                    double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);
int fused_wstrip512_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                        double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);
int fused_strip512_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                       double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);
int fused_fast_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                   double *d_sums, Scratch &s, cudaStream_t stream, bool *handled);

}  // namespace sep
