// score.cu -- SI-SDR / SDR scoring of a ragged batch of utterances.
//
// Reference: pow_np_norm / pow_norm / si_sdr / permute_si_sdr / eval_si_sdr /
// eval_sdr, metrics/evaluate_metrics.py:14-92.  The reference makes five
// passes over every signal pair; here every sample of the C references and C
// estimates is read exactly once: each CTA streams one chunk of one utterance
// and accumulates the Gram statistics <e_i, r_j>, |e_i|^2, |r_j|^2 in float64
// (products of float32 are exact in float64, so the one-pass identity
// |noise|^2 = |e|^2 - <e,r>^2/|r|^2 does not cancel catastrophically).
// Chunk partials are reduced per utterance in chunk order (deterministic).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "score.cuh"

namespace sep {

constexpr int kScoreChunk = 8192;   // samples per CTA
constexpr int kScoreThreads = 256;

template <int C, bool VEC>
__device__ __forceinline__ void score_accumulate(const float *const *r, const float *const *e,
                                                 int count, double *gram, double *ee, double *er) {
  constexpr int V = VEC ? 4 : 1;
  for (int i = threadIdx.x * V; i < count; i += kScoreThreads * V) {
    float rv[C][V], ev[C][V];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (VEC) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(r[c] + i));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(e[c] + i));
        rv[c][0] = a.x; rv[c][V > 1 ? 1 : 0] = a.y; rv[c][V > 2 ? 2 : 0] = a.z; rv[c][V > 3 ? 3 : 0] = a.w;
        ev[c][0] = b.x; ev[c][V > 1 ? 1 : 0] = b.y; ev[c][V > 2 ? 2 : 0] = b.z; ev[c][V > 3 ? 3 : 0] = b.w;
      } else {
        rv[c][0] = __ldg(r[c] + i);
        ev[c][0] = __ldg(e[c] + i);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (i + v >= count) break;
      double rd[C], ed[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { rd[c] = rv[c][v]; ed[c] = ev[c][v]; }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        er[c] = fma(rd[c], rd[c], er[c]);
        ee[c] = fma(ed[c], ed[c], ee[c]);
#pragma unroll
        for (int j = 0; j < C; ++j) gram[c * C + j] = fma(ed[c], rd[j], gram[c * C + j]);
      }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kScoreThreads)
score_chunk_kernel(const float *__restrict__ refs, const float *__restrict__ ests,
                   const int64_t *__restrict__ ref_off, const int64_t *__restrict__ est_off,
                   const int64_t *__restrict__ lengths, const int32_t *__restrict__ chunk_start,
                   int batch, double *__restrict__ partials) {
  constexpr int NV = C * C + 2 * C;
  __shared__ double red[NV * (kScoreThreads / 32)];
  __shared__ int s_b;
  const int item = blockIdx.x;
  if (threadIdx.x == 0) {   // utterance owning this chunk: last b with chunk_start[b] <= item
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_start[mid] <= item) lo = mid; else hi = mid - 1;
    }
    s_b = lo;
  }
  __syncthreads();
  const int b = s_b;
  const int64_t begin = static_cast<int64_t>(item - chunk_start[b]) * kScoreChunk;
  const int count = static_cast<int>(min(static_cast<int64_t>(kScoreChunk), lengths[b] - begin));
  const float *r[C], *e[C];
  bool aligned = true;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    r[c] = refs + ref_off[b * C + c] + begin;
    e[c] = ests + est_off[b * C + c] + begin;
    aligned = aligned && ((reinterpret_cast<uintptr_t>(r[c]) | reinterpret_cast<uintptr_t>(e[c])) & 15) == 0;
  }
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.0;
  if (aligned) {
    const int body = count & ~3;
    score_accumulate<C, true>(r, e, body, v, v + C * C, v + C * C + C);
    if (body < count) {
      const float *rt[C], *et[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { rt[c] = r[c] + body; et[c] = e[c] + body; }
      score_accumulate<C, false>(rt, et, count - body, v, v + C * C, v + C * C + C);
    }
  } else {
    score_accumulate<C, false>(r, e, count, v, v + C * C, v + C * C + C);
  }
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
    double *dst = partials + static_cast<int64_t>(item) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = v[i];
  }
}

template <int C>
__global__ void score_finalize_kernel(const double *__restrict__ partials,
                                      const int32_t *__restrict__ chunk_start,
                                      double *__restrict__ scores) {
  constexpr int NV = C * C + 2 * C;
  const int b = blockIdx.x, lane = threadIdx.x;
  double v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.0;
  for (int q = chunk_start[b] + lane; q < chunk_start[b + 1]; q += 32) {
    const double *src = partials + static_cast<int64_t>(q) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += src[i];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0)
    finalize_scores<C>(v, v + C * C, v + C * C + C, scores + static_cast<int64_t>(b) * (2 * C * C + 4));
}

// sums[3] = {sum si_best, sum sdr_best, batch}, one CTA, fixed order.
__global__ void score_sums_kernel(const double *__restrict__ scores, int batch, int C,
                                  double *__restrict__ sums) {
  __shared__ double red[2 * 8];
  const int stride = 2 * C * C + 4;
  double v[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    v[0] += scores[static_cast<int64_t>(b) * stride + C * C];
    v[1] += scores[static_cast<int64_t>(b) * stride + 2 * C * C + 2];
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) { sums[0] = v[0]; sums[1] = v[1]; sums[2] = batch; }
}


// <a, b> in float64: grid-stride partials, then one CTA adds them in block order.
__global__ void __launch_bounds__(256)
dot_partial_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n,
                   double *__restrict__ partials) {
  __shared__ double red[8];
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    v[0] = fma(static_cast<double>(__ldg(a + i)), static_cast<double>(__ldg(b + i)), v[0]);
  block_sum<1>(v, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
}

__global__ void dot_final_kernel(const double *__restrict__ partials, int count,
                                 double *__restrict__ out) {
  __shared__ double red[8];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < count; i += blockDim.x) v[0] += partials[i];
  block_sum<1>(v, red);
  if (threadIdx.x == 0) *out = v[0];
}

template <int C>
static int run_score(const float *d_refs, const float *d_ests, const int64_t *d_roff,
                     const int64_t *d_eoff, const int64_t *d_len, const int32_t *d_start, int batch,
                     int items, double *d_scores, double *d_sums, Scratch &s, cudaStream_t stream) {
  constexpr int NV = C * C + 2 * C;
  double *partials;
  int rc;
  if ((rc = s.alloc(&partials, static_cast<size_t>(items) * NV))) return rc;
  if (items > 0) {
    profile_begin(stream, "score_chunk_kernel<C=%d>", C);
    score_chunk_kernel<C><<<items, kScoreThreads, 0, stream>>>(d_refs, d_ests, d_roff, d_eoff, d_len,
                                                               d_start, batch, partials);
    profile_end(stream);
    SEP_LAUNCHED();
  }
  score_finalize_kernel<C><<<batch, 32, 0, stream>>>(partials, d_start, d_scores);
  SEP_LAUNCHED();
  if (d_sums) {
    score_sums_kernel<<<1, 256, 0, stream>>>(d_scores, batch, C, d_sums);
    SEP_LAUNCHED();
  }
  return SEP_OK;
}

// Pinned staging for per-call metadata: a small per-thread ring; a slot is reused only after the copy that
// read it has completed (event), so entry points can stay asynchronous.
struct MetaSlot { void *host = nullptr; size_t cap = 0; cudaEvent_t copied = nullptr; };
static int meta_slot(size_t bytes, MetaSlot **out) {
  static thread_local MetaSlot ring[4];
  static thread_local unsigned next = 0;
  MetaSlot &m = ring[next++ % 4];
  if (m.copied == nullptr) SEP_CUDA(cudaEventCreateWithFlags(&m.copied, cudaEventDisableTiming));
  else SEP_CUDA(cudaEventSynchronize(m.copied));
  if (m.cap < bytes) {
    if (m.host) cudaFreeHost(m.host);
    m.host = nullptr;
    m.cap = 0;
    if (cudaMallocHost(&m.host, bytes + bytes / 2) != cudaSuccess) {
      set_error("sep_score_batch_f32: cannot allocate %zu bytes of pinned staging memory", bytes);
      return SEP_ERR_NOMEM;
    }
    m.cap = bytes + bytes / 2;
  }
  *out = &m;
  return SEP_OK;
}

}  // namespace sep

using namespace sep;

// ref_offsets / est_offsets / lengths are HOST arrays in both memory modes
// (small metadata the host computes when it applies the truncate-to-min rule).
extern "C" int sep_score_batch_f32(const float *refs, const float *ests, const int64_t *ref_offsets,
                                   const int64_t *est_offsets, const int64_t *lengths, int batch,
                                   int n_src, int64_t total_ref, int64_t total_est, double *scores,
                                   double *sums, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(refs && ests && ref_offsets && est_offsets && lengths && scores,
              "sep_score_batch_f32: null argument");
  SEP_REQUIRE(batch >= 1, "sep_score_batch_f32: empty batch");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES, "sep_score_batch_f32: n_src=%d out of range", n_src);
  int rc = check_mem(mem);
  if (rc) return rc;
  const int C = n_src;
  // All metadata travels in ONE pinned staging buffer (ring of 4, reused once its copy has completed), so
  // the call never blocks on the stream: [ref_off B*C | est_off B*C | lengths B] int64, [chunk_start B+1] int32.
  const size_t n_off = static_cast<size_t>(batch) * C;
  const size_t meta_bytes = (2 * n_off + batch) * sizeof(int64_t) + (static_cast<size_t>(batch) + 1) * sizeof(int32_t);
  MetaSlot *slot = nullptr;
  if ((rc = meta_slot(meta_bytes, &slot))) return rc;
  int64_t *h_roff = static_cast<int64_t *>(slot->host), *h_eoff = h_roff + n_off, *h_len = h_eoff + n_off;
  int32_t *start = reinterpret_cast<int32_t *>(h_len + batch);
  start[0] = 0;
  for (int b = 0; b < batch; ++b) {
    SEP_REQUIRE(lengths[b] >= 0, "sep_score_batch_f32: negative length at %d", b);
    for (int c = 0; c < C; ++c) {
      SEP_REQUIRE(ref_offsets[b * C + c] >= 0 && ref_offsets[b * C + c] + lengths[b] <= total_ref,
                  "sep_score_batch_f32: reference %d/%d out of bounds", b, c);
      SEP_REQUIRE(est_offsets[b * C + c] >= 0 && est_offsets[b * C + c] + lengths[b] <= total_est,
                  "sep_score_batch_f32: estimate %d/%d out of bounds", b, c);
      h_roff[b * C + c] = ref_offsets[b * C + c];
      h_eoff[b * C + c] = est_offsets[b * C + c];
    }
    h_len[b] = lengths[b];
    const int64_t chunks = (lengths[b] + kScoreChunk - 1) / kScoreChunk;
    SEP_REQUIRE(start[b] + chunks < INT32_MAX, "sep_score_batch_f32: batch too large");
    start[b + 1] = start[b] + static_cast<int32_t>(chunks);
  }
  Scratch s(stream);
  const float *d_refs, *d_ests;
  if ((rc = stage_in(s, refs, static_cast<size_t>(total_ref), mem, &d_refs))) return rc;
  if ((rc = stage_in(s, ests, static_cast<size_t>(total_est), mem, &d_ests))) return rc;
  unsigned char *d_meta;
  if ((rc = s.alloc(&d_meta, meta_bytes))) return rc;
  SEP_CUDA(cudaMemcpyAsync(d_meta, slot->host, meta_bytes, cudaMemcpyHostToDevice, stream));
  SEP_CUDA(cudaEventRecord(slot->copied, stream));
  const int64_t *d_roff = reinterpret_cast<const int64_t *>(d_meta), *d_eoff = d_roff + n_off, *d_len = d_eoff + n_off;
  const int32_t *d_start = reinterpret_cast<const int32_t *>(d_len + batch);
  double *d_scores, *d_sums;
  const size_t n_scores = static_cast<size_t>(batch) * (2 * C * C + 4);
  if ((rc = stage_out(s, scores, n_scores, mem, &d_scores))) return rc;
  if ((rc = stage_out(s, sums, static_cast<size_t>(3), mem, &d_sums))) return rc;
  const int items = start[batch];
  switch (C) {
    case 1: rc = run_score<1>(d_refs, d_ests, d_roff, d_eoff, d_len, d_start, batch, items, d_scores, d_sums, s, stream); break;
    case 2: rc = run_score<2>(d_refs, d_ests, d_roff, d_eoff, d_len, d_start, batch, items, d_scores, d_sums, s, stream); break;
    case 3: rc = run_score<3>(d_refs, d_ests, d_roff, d_eoff, d_len, d_start, batch, items, d_scores, d_sums, s, stream); break;
    default: rc = run_score<4>(d_refs, d_ests, d_roff, d_eoff, d_len, d_start, batch, items, d_scores, d_sums, s, stream); break;
  }
  if (rc) return rc;
  if ((rc = copy_back(s, scores, d_scores, n_scores, mem))) return rc;
  if ((rc = copy_back(s, sums, d_sums, static_cast<size_t>(3), mem))) return rc;
  return finish(s, mem);      // device mode: asynchronous (the metadata sits in the pinned ring)
}

// pow_norm / pow_np_norm (metrics/evaluate_metrics.py:14-20): sum(a * b) in float64.
extern "C" int sep_dot_f32(const float *a, const float *b, int64_t n, double *out, int mem,
                           void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(a && b && out && n >= 0, "sep_dot_f32: bad argument");
  int rc = check_mem(mem);
  if (rc) return rc;
  Scratch s(stream);
  const float *d_a, *d_b;
  double *d_out, *partials;
  if ((rc = stage_in(s, a, static_cast<size_t>(n), mem, &d_a))) return rc;
  if (b == a) d_b = d_a;
  else if ((rc = stage_in(s, b, static_cast<size_t>(n), mem, &d_b))) return rc;
  if ((rc = stage_out(s, out, static_cast<size_t>(1), mem, &d_out))) return rc;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 2047) / 2048, 1184)));
  if ((rc = s.alloc(&partials, static_cast<size_t>(blocks)))) return rc;
  dot_partial_kernel<<<blocks, 256, 0, stream>>>(d_a, d_b, n, partials);
  SEP_LAUNCHED();
  dot_final_kernel<<<1, 256, 0, stream>>>(partials, blocks, d_out);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, out, d_out, static_cast<size_t>(1), mem))) return rc;
  return finish(s, mem);
}
