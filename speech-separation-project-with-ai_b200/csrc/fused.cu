// fused.cu -- the fused hot path: wave -> STFT -> mask -> iSTFT -> overlap-add,
// with PSA labels + PIT-MSE partials and SI-SDR/SDR Gram partials as epilogues.
// Spectra never touch HBM.
//
// Reference chain being fused:
//   stft                parallel_stft.py:146-196      (mixture and each source)
//   |X|, PSA labels     parallel_stft.py:262-272
//   mask * |X|          uPIT_baseline.ipynb:1087-1088 (cell 29)
//   * exp(j angle X)    uPIT_baseline.ipynb:1385-1388 (cell 41)   == mask * X
//   istft               uPIT_baseline.ipynb:1269-1307 (cell 39)
//   pit_loss            uPIT_baseline.ipynb:1023-1059 (cell 28)
//   si_sdr / permute    metrics/evaluate_metrics.py:14-34
//
// Work split: a CTA owns TB consecutive frames of one utterance ("owned"
// frames, counted once in the PIT sums) and the TB output hop-blocks with the
// same indices.  It also transforms the size/shift - 1 frames to the left of
// its tile (halo, recomputed, never exchanged) so that overlap-add completes
// inside shared memory without atomics.  Per-tile partial sums go to a
// float64 workspace; a second tiny kernel reduces them per utterance in tile
// order (deterministic) and runs the permutation searches.
//
// This file holds the generic path (any power-of-two size, 1..4 sources); the
// register-resident 256/512-point specialisation lives in fused_fast.cu.
#include <initializer_list>

#include "common.cuh"
#include "fft.cuh"
#include "fused.cuh"
#include "score.cuh"

namespace sep {

template <int C>
__global__ void __launch_bounds__(256) fused_generic_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int size = a.size, shift = a.shift, M = size >> 1, F = M + 1, R = size / shift;
  const int T = a.T, tb = a.tb;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool score = a.refs != nullptr;

  const int own_lo = tile * tb, own_hi = min(own_lo + tb, T);       // owned frames
  const int j0 = max(own_lo, R - 1), j1 = own_hi;                   // output hop-blocks
  const int t_lo = max(own_lo - R + 1, 0), t_hi = own_hi;           // frames transformed
  const int nframes = t_hi - t_lo, slots = tb + R - 1;
  const int tile_cap = ((slots - 1) * shift + size + 3) & ~3;
  const int tile_len = (nframes - 1) * shift + size;
  const int64_t s0 = static_cast<int64_t>(t_lo) * shift - a.pad;    // original index of tile[0]

  // shared layout: wave tiles [1+C][tile_cap] | frame buffer [C][slots][size] | per-warp scratch | red
  float *tiles = reinterpret_cast<float *>(smem_raw);
  float *fb = tiles + static_cast<size_t>(tile_cap) * (1 + C);
  float2 *bufs = reinterpret_cast<float2 *>(fb + static_cast<size_t>(C) * slots * size);
  const int fpad = (F + 1) & ~1;                       // floats, keeps float2 alignment
  const int per_warp_f2 = 2 * M + F + C * fpad;        // A, B, X | labels + masks as floats
  float2 *A = bufs + static_cast<size_t>(warp) * per_warp_f2, *B = A + M, *XB = B + M;
  float *LB = reinterpret_cast<float *>(XB + F);       // [C][fpad] labels
  float *MB = LB + C * fpad;                           // [C][fpad] masks
  double *red = reinterpret_cast<double *>(bufs + static_cast<size_t>(nwarps) * per_warp_f2);

  // ---- stage the contiguous waveform spans (fade / tail padding = zeros) ----
  const int nsig = score ? 1 + C : 1;
  for (int sgn = 0; sgn < nsig; ++sgn) {
    const float *row = sgn == 0 ? a.mix + static_cast<int64_t>(b) * a.n
                                : a.refs + (static_cast<int64_t>(b) * C + (sgn - 1)) * a.n;
    float *dst = tiles + static_cast<size_t>(tile_cap) * sgn;
    for (int i = threadIdx.x; i < tile_len; i += blockDim.x) {
      const int64_t g = s0 + i;
      dst[i] = (g >= 0 && g < a.n) ? __ldg(row + g) : 0.f;
    }
  }
  __syncthreads();

  const float len_f = a.lengths ? a.lengths[b] : static_cast<float>(T);
  const int len_i = static_cast<int>(len_f);
  double pit[C * C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) pit[i] = 0.0;

  for (int f = warp; f < nframes; f += nwarps) {
    const int t = t_lo + f;
    const bool owned = t >= own_lo;
    // mixture spectrum
    load_frame_packed(A, tiles + f * shift, a.win_half, M, lane);
    const float2 *Z = warp_fft<false>(A, B, a.tw_half, M, lane);
    for (int k = lane; k <= M; k += 32) XB[k] = real_split(Z, a.tw_full, M, k);
    // masks of this frame, all sources (coalesced rows of F floats)
    for (int c = 0; c < C; ++c) {
      const float *mrow = a.masks + ((static_cast<int64_t>(b) * C + c) * T + t) * F;
      for (int k = lane; k <= M; k += 32) MB[c * fpad + k] = __ldg(mrow + k);
    }
    __syncwarp();
    if (score && owned) {
      // PSA labels of every source: |S| cos(angle X - angle S) = Re(S conj X) / |X|
      for (int c = 0; c < C; ++c) {
        load_frame_packed(A, tiles + static_cast<size_t>(tile_cap) * (1 + c) + f * shift,
                          a.win_half, M, lane);
        const float2 *Zs = warp_fft<false>(A, B, a.tw_half, M, lane);
        for (int k = lane; k <= M; k += 32) {
          const float2 s = real_split(Zs, a.tw_full, M, k), x = XB[k];
          const float mag = hypotf(x.x, x.y);
          LB[c * fpad + k] = mag > 0.f ? fmaf(s.x, x.x, s.y * x.y) / mag : s.x;
        }
        __syncwarp();
      }
      // pair[i][j] += (m_t * mask_i |X| - label_j)^2     (cell 28 :1045-1052)
      const float gate = t < len_i ? 1.f : 0.f;
      float acc[C * C];
#pragma unroll
      for (int i = 0; i < C * C; ++i) acc[i] = 0.f;
      for (int k = lane; k <= M; k += 32) {
        const float2 x = XB[k];
        const float mag = hypotf(x.x, x.y) * gate;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float pred = MB[i * fpad + k] * mag;
#pragma unroll
          for (int j = 0; j < C; ++j) {
            const float d = pred - LB[j * fpad + k];
            acc[i * C + j] = fmaf(d, d, acc[i * C + j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < C * C; ++i) pit[i] += static_cast<double>(acc[i]);
    }
    // masked spectra -> time frames (spec_c = mask_c * X, cell 41) -> synthesis window
    for (int c = 0; c < C; ++c) {
      for (int k = lane; k < M; k += 32) {
        float2 yk = XB[k], ym = XB[M - k];
        const float mk = MB[c * fpad + k], mm = MB[c * fpad + M - k];
        yk.x *= mk; yk.y *= mk; ym.x *= mm; ym.y *= mm;
        if (k == 0) { yk.y = 0.f; ym.y = 0.f; }
        A[k] = real_merge(yk, ym, a.tw_full[k]);
      }
      __syncwarp();
      const float2 *z = warp_fft<true>(A, B, a.tw_half, M, lane);
      float2 *dst = reinterpret_cast<float2 *>(fb + (static_cast<size_t>(c) * slots + f) * size);
      for (int m = lane; m < M; m += 32) {
        const float2 w = reinterpret_cast<const float2 *>(a.syn)[m];
        dst[m] = make_float2(z[m].x * w.x, z[m].y * w.y);
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- overlap-add in frame order, write estimates, Gram partials ----
  double gram[C * C], ee[C], er[C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) gram[i] = 0.0;
#pragma unroll
  for (int i = 0; i < C; ++i) { ee[i] = 0.0; er[i] = 0.0; }
  const int64_t n_valid = a.valid ? min(static_cast<int64_t>(a.valid[b]), a.n) : a.n;
  const int span = (j1 - j0) * shift;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int j = j0 + i / shift, m = i % shift;
    const int64_t g = static_cast<int64_t>(j) * shift + m - a.pad;   // original sample index
    if (g >= a.n) continue;
    float e[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float *fbc = fb + static_cast<size_t>(c) * slots * size;
      float acc = 0.f;
      for (int t = max(j - R + 1, t_lo); t <= min(j, t_hi - 1); ++t)
        acc += fbc[(t - t_lo) * size + (j - t) * shift + m];
      e[c] = acc;
      if (a.est) a.est[(static_cast<int64_t>(b) * C + c) * a.n + g] = acc;
    }
    if (score && g < n_valid) {
      const int local = static_cast<int>(g - s0);
#pragma unroll
      for (int jr = 0; jr < C; ++jr) {
        const double r = static_cast<double>(tiles[static_cast<size_t>(tile_cap) * (1 + jr) + local]);
        er[jr] = fma(r, r, er[jr]);
#pragma unroll
        for (int ie = 0; ie < C; ++ie)
          gram[ie * C + jr] = fma(static_cast<double>(e[ie]), r, gram[ie * C + jr]);
      }
#pragma unroll
      for (int ie = 0; ie < C; ++ie) {
        const double v = static_cast<double>(e[ie]);
        ee[ie] = fma(v, v, ee[ie]);
      }
    }
  }
  if (!score) return;

  constexpr int NV = FusedVals<C>::NV;
  double v[NV];
#pragma unroll
  for (int i = 0; i < C * C; ++i) { v[i] = pit[i]; v[C * C + i] = gram[i]; }
#pragma unroll
  for (int i = 0; i < C; ++i) { v[2 * C * C + i] = ee[i]; v[2 * C * C + C + i] = er[i]; }
  block_sum<NV>(v, red);
  if (threadIdx.x == 0) {
    double *dst = a.partials + (static_cast<int64_t>(b) * a.tiles + tile) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = v[i];
  }
}

struct FusedCfg { int warps, tb; size_t smem; };

static bool pick_fused(const sep_plan *p, int C, FusedCfg *out) {
  const int R = p->hops, F = p->bins, M = p->half;
  const int fpad = (F + 1) & ~1;
  for (int warps : {8, 4, 2, 1}) {
    for (int tb : {32, 16, 8, 4, 2, 1}) {
      const int slots = tb + R - 1;
      const size_t tile_cap = (static_cast<size_t>(slots - 1) * p->shift + p->size + 3) & ~size_t(3);
      size_t bytes = tile_cap * 4 * (1 + C)                    // wave tiles (layout fixed at 1 + C)
                   + static_cast<size_t>(C) * slots * p->size * 4   // windowed time frames
                   + static_cast<size_t>(warps) * (2 * M + F + C * fpad) * 8
                   + static_cast<size_t>(warps) * (2 * C * C + 2 * C) * 8 + 64;
      if (bytes <= 200 * 1024) {
        *out = FusedCfg{warps, tb, bytes};
        return true;
      }
    }
  }
  return false;
}

template <int C>
static int run_fused(const sep_plan *p, FusedArgs a, int batch, double *d_scores, double *d_sums,
                     Scratch &s, cudaStream_t stream) {
  const bool score = a.refs != nullptr;
  FusedCfg cfg;
  if (!pick_fused(p, C, &cfg)) {
    set_error("fused path: size=%d shift=%d sources=%d does not fit in shared memory", p->size,
              p->shift, C);
    return SEP_ERR_UNSUPPORTED;
  }
  a.tb = cfg.tb;
  a.tiles = (a.T + cfg.tb - 1) / cfg.tb;
  constexpr int NV = FusedVals<C>::NV;
  double *partials = nullptr;
  int rc;
  if (score && (rc = s.alloc(&partials, static_cast<size_t>(batch) * a.tiles * NV))) return rc;
  a.partials = partials;
  SEP_CUDA(cudaFuncSetAttribute(fused_generic_kernel<C>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(cfg.smem)));
  dim3 grid(a.tiles, batch);
  profile_begin(stream, "fused_generic_kernel<C=%d> (size=%d shift=%d)", C, p->size, p->shift);
  fused_generic_kernel<C><<<grid, cfg.warps * 32, cfg.smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (score) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

}  // namespace sep

using namespace sep;

extern "C" int sep_fused_workspace_bytes(const sep_plan *p, int batch, int n_src, int64_t n_samples,
                                         int64_t *bytes) {
  SEP_REQUIRE(p && bytes, "sep_fused_workspace_bytes: null argument");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES && batch >= 1 && n_samples >= 1,
              "sep_fused_workspace_bytes: bad shape");
  int T = 0;
  sep_plan_frames(p, n_samples, &T);
  // worst case: one tile per frame, NV doubles per tile, plus alignment slack
  const int64_t nv = 2 * n_src * n_src + 2 * n_src;
  *bytes = static_cast<int64_t>(batch) * T * nv * 8 + (static_cast<int64_t>(batch) + 1) * 4 + 4096;
  return SEP_OK;
}

extern "C" int sep_fused_separate_f32(const sep_plan *p, const float *mix, const float *masks,
                                      const float *refs, const float *frame_lengths,
                                      const int32_t *valid_samples, int batch, int n_src,
                                      int64_t n_samples, float *est, double *scores, double *sums,
                                      int mem, void *stream) {
  return sep_fused_separate_ws_f32(p, mix, masks, refs, frame_lengths, valid_samples, batch, n_src,
                                   n_samples, est, scores, sums, nullptr, 0, mem, stream);
}

struct PushTarget {
  double *const *peers = nullptr;
  int world = 0, rank = 0, slot = 0, slots = 0;
};

static int fused_separate(const sep_plan *p, const float *mix, const float *masks, const float *refs,
                          const float *frame_lengths, const int32_t *valid_samples, int batch, int n_src,
                          int64_t n_samples, float *est, double *scores, double *sums, void *workspace,
                          int64_t workspace_bytes, int mem, void *stream_, const PushTarget &push);

extern "C" int sep_fused_separate_ws_f32(const sep_plan *p, const float *mix, const float *masks,
                                         const float *refs, const float *frame_lengths,
                                         const int32_t *valid_samples, int batch, int n_src,
                                         int64_t n_samples, float *est, double *scores,
                                         double *sums, void *workspace, int64_t workspace_bytes,
                                         int mem, void *stream_) {
  return fused_separate(p, mix, masks, refs, frame_lengths, valid_samples, batch, n_src, n_samples, est, scores, sums,
                        workspace, workspace_bytes, mem, stream_, PushTarget{});
}

extern "C" int sep_fused_separate_push_f32(const sep_plan *p, const float *mix, const float *masks,
                                           const float *refs, const float *frame_lengths,
                                           const int32_t *valid_samples, int batch, int n_src,
                                           int64_t n_samples, float *est, double *scores, double *sums,
                                           void *workspace, int64_t workspace_bytes,
                                           void *const *peer_inboxes, int world, int rank, int slot, int slots,
                                           void *stream_) {
  SEP_REQUIRE(peer_inboxes && world >= 1 && world <= 32 && rank >= 0 && rank < world && slots >= 1 && slot >= 0 &&
                  slot < slots,
              "sep_fused_separate_push_f32: bad push target (world=%d rank=%d slot=%d slots=%d)", world, rank, slot, slots);
  SEP_REQUIRE(refs && scores && sums, "sep_fused_separate_push_f32: the pushed sums need refs, scores and sums");
  PushTarget push;
  push.peers = reinterpret_cast<double *const *>(peer_inboxes);
  push.world = world;
  push.rank = rank;
  push.slot = slot;
  push.slots = slots;
  return fused_separate(p, mix, masks, refs, frame_lengths, valid_samples, batch, n_src, n_samples, est, scores, sums,
                        workspace, workspace_bytes, SEP_MEM_DEVICE, stream_, push);
}

static int fused_separate(const sep_plan *p, const float *mix, const float *masks, const float *refs,
                          const float *frame_lengths, const int32_t *valid_samples, int batch, int n_src,
                          int64_t n_samples, float *est, double *scores, double *sums, void *workspace,
                          int64_t workspace_bytes, int mem, void *stream_, const PushTarget &push) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(p && mix && masks, "sep_fused_separate_f32: null argument");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES, "sep_fused_separate_f32: n_src=%d out of range",
              n_src);
  SEP_REQUIRE(batch >= 1 && n_samples >= 1, "sep_fused_separate_f32: bad shape");
  SEP_REQUIRE(p->hops > 0 && p->fading,
              "fused path needs size %% shift == 0 and fading (size=%d shift=%d fading=%d)", p->size,
              p->shift, p->fading);
  if (!p->pow2) {
    set_error("fused path: size=%d is not a power of two >= 32 (only stft / istft take other sizes)", p->size);
    return SEP_ERR_UNSUPPORTED;
  }
  SEP_REQUIRE(refs != nullptr || (scores == nullptr && sums == nullptr),
              "sep_fused_separate_f32: scores/sums need refs");
  SEP_REQUIRE(refs == nullptr || scores != nullptr, "sep_fused_separate_f32: refs given but scores is null");
  SEP_REQUIRE(est != nullptr || refs != nullptr, "sep_fused_separate_f32: nothing to compute");
  int rc = check_mem(mem);
  if (rc) return rc;
  int T = 0;
  sep_plan_frames(p, n_samples, &T);
  const int C = n_src, stride = sep_score_stride(C);
  Scratch s(stream);
  if (workspace && mem == SEP_MEM_DEVICE) s.use_arena(workspace, static_cast<size_t>(workspace_bytes));
  FusedArgs a{};
  const size_t wave_count = static_cast<size_t>(batch) * n_samples;
  const size_t mask_count = static_cast<size_t>(batch) * C * T * p->bins;
  if ((rc = stage_in(s, mix, wave_count, mem, &a.mix))) return rc;
  if ((rc = stage_in(s, masks, mask_count, mem, &a.masks))) return rc;
  if ((rc = stage_in(s, refs, wave_count * C, mem, &a.refs))) return rc;
  if ((rc = stage_in(s, frame_lengths, static_cast<size_t>(batch), mem, &a.lengths))) return rc;
  if ((rc = stage_in(s, valid_samples, static_cast<size_t>(batch), mem, &a.valid))) return rc;
  double *d_scores, *d_sums;
  if ((rc = stage_out(s, est, wave_count * C, mem, &a.est))) return rc;
  if ((rc = stage_out(s, scores, static_cast<size_t>(batch) * stride, mem, &d_scores))) return rc;
  if ((rc = stage_out(s, sums, static_cast<size_t>(4), mem, &d_sums))) return rc;
  a.n = n_samples;
  a.batch = batch;
  a.peers = push.peers;
  a.world = push.world;
  a.rank = push.rank;
  a.push_slot = push.slot;
  a.push_slots = push.slots;
  a.lookahead = 0;
  a.T = T;
  a.size = p->size;
  a.shift = p->shift;
  a.pad = p->pad;
  a.win_half = p->d_win_half;
  a.syn = p->d_syn;
  a.tw_half = p->d_tw_half;
  a.tw_full = p->d_tw_full;
  a.tw16 = p->d_tw16;
  a.win_t = p->d_win_t;
  a.syn_t = p->d_syn_t;
  a.win2_t = p->d_win2_t;
  a.syn2_t = p->d_syn2_t;
  a.tw512_t = p->d_tw512_t;

  bool handled = false;
  if ((rc = fused_wstrip_try(p, a, batch, C, d_scores, d_sums, s, stream, &handled))) return rc;
  if (!handled && (rc = fused_strip_try(p, a, batch, C, d_scores, d_sums, s, stream, &handled))) return rc;
  if (!handled && (rc = fused_wstrip512_try(p, a, batch, C, d_scores, d_sums, s, stream, &handled))) return rc;
  if (!handled && (rc = fused_strip512_try(p, a, batch, C, d_scores, d_sums, s, stream, &handled))) return rc;
  if (!handled && (rc = fused_fast_try(p, a, batch, C, d_scores, d_sums, s, stream, &handled))) return rc;
  if (!handled) {
    switch (C) {
      case 1: rc = run_fused<1>(p, a, batch, d_scores, d_sums, s, stream); break;
      case 2: rc = run_fused<2>(p, a, batch, d_scores, d_sums, s, stream); break;
      case 3: rc = run_fused<3>(p, a, batch, d_scores, d_sums, s, stream); break;
      default: rc = run_fused<4>(p, a, batch, d_scores, d_sums, s, stream); break;
    }
    if (rc) return rc;
  }
  if ((rc = copy_back(s, est, a.est, wave_count * C, mem))) return rc;
  if ((rc = copy_back(s, scores, d_scores, static_cast<size_t>(batch) * stride, mem))) return rc;
  if ((rc = copy_back(s, sums, d_sums, static_cast<size_t>(4), mem))) return rc;
  return finish(s, mem);
}
