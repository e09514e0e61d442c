// stft.cu -- framing, batched STFT and the |X| / angle X / PSA-label feature kernel.
//
// Reference: segment_axis parallel_stft.py:37-123, stft :146-196, feature and
// label math :262-272.  A CTA owns a tile of consecutive frames of one
// utterance: the contiguous waveform span behind the tile is staged once in
// shared memory (the reference's strided "view"), each warp transforms one
// frame at a time, and the tile's output rows form one contiguous, coalesced
// span of the [T, F] spectrum.
#include "common.cuh"
#include "fft.cuh"

namespace sep {

struct TileCfg {
  int warps;       // warps per CTA
  int frames;      // frames per tile
  size_t smem;     // dynamic shared memory bytes
};

// signals = waveforms staged per tile; extra_f2_per_warp = float2 slots per warp
// besides the 2*M ping-pong.
static TileCfg pick_tile(const sep_plan *p, int signals, int extra_f2_per_warp) {
  TileCfg c{8, 32, 0};
  for (;;) {
    size_t tile = (static_cast<size_t>(c.frames - 1) * p->shift + p->size) * sizeof(float) * signals;
    size_t per_warp = (2 * static_cast<size_t>(p->half) + extra_f2_per_warp) * sizeof(float2);
    c.smem = tile + per_warp * c.warps;
    if (c.smem <= 96 * 1024 || (c.frames == 1 && c.warps == 1)) break;
    if (c.frames > c.warps) c.frames >>= 1; else c.warps >>= 1;
  }
  return c;
}

// Stages `tile_len` samples starting at original index s0 (may be negative or
// run past n: the fade / tail padding of parallel_stft.py:169-180 is zeros).
__device__ __forceinline__ void stage_wave(float *tile, const float *__restrict__ row, int64_t n,
                                           int64_t s0, int tile_len) {
  for (int i = threadIdx.x; i < tile_len; i += blockDim.x) {
    const int64_t g = s0 + i;
    tile[i] = (g >= 0 && g < n) ? __ldg(row + g) : 0.f;
  }
}

__global__ void stft_kernel(const float *__restrict__ wave, int64_t n, int64_t stride, int T,
                            int size, int shift, int pad, int frames_per_tile,
                            const float *__restrict__ win_half, const float2 *__restrict__ tw_half,
                            const float2 *__restrict__ tw_full, float2 *__restrict__ spec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = size >> 1, F = M + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * frames_per_tile;
  const int nframes = min(frames_per_tile, T - t0);
  const int tile_cap = (frames_per_tile - 1) * shift + size;
  const int tile_len = (nframes - 1) * shift + size;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float *tile = reinterpret_cast<float *>(smem_raw);
  float2 *bufs = reinterpret_cast<float2 *>(tile + ((tile_cap + 3) & ~3));
  float2 *A = bufs + static_cast<size_t>(warp) * 2 * M, *B = A + M;

  stage_wave(tile, wave + static_cast<int64_t>(b) * stride, n,
             static_cast<int64_t>(t0) * shift - pad, tile_len);
  __syncthreads();

  for (int f = warp; f < nframes; f += nwarps) {
    load_frame_packed(A, tile + f * shift, win_half, M, lane);
    const float2 *Z = warp_fft<false>(A, B, tw_half, M, lane);
    float2 *row = spec + (static_cast<int64_t>(b) * T + t0 + f) * F;
    for (int k = lane; k <= M; k += 32) row[k] = real_split(Z, tw_full, M, k);
    __syncwarp();
  }
}

// Any size (parallel_stft.py:146 takes a free `size`; every call site of the reference uses 256): the real DFT of a
// windowed frame by direct summation, X[k] = sum_m xw[m] W_n^(k m), k <= n / 2 -- O(n^2) per frame, the twiddle index
// walks the table modulo n without a division.  A compatibility path, not a hot path.
__global__ void stft_dft_kernel(const float *__restrict__ wave, int64_t n, int64_t stride, int T, int size, int shift,
                                int pad, int frames_per_tile, const float *__restrict__ win_full,
                                const float2 *__restrict__ tw_n, float2 *__restrict__ spec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = size / 2 + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * frames_per_tile;
  const int nframes = min(frames_per_tile, T - t0);
  const int tile_cap = (frames_per_tile - 1) * shift + size;
  const int tile_len = (nframes - 1) * shift + size;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float *tile = reinterpret_cast<float *>(smem_raw);
  float2 *tw = reinterpret_cast<float2 *>(tile + ((tile_cap + 3) & ~3));
  float *xw = reinterpret_cast<float *>(tw + size) + static_cast<size_t>(warp) * size;
  stage_wave(tile, wave + static_cast<int64_t>(b) * stride, n, static_cast<int64_t>(t0) * shift - pad, tile_len);
  for (int i = threadIdx.x; i < size; i += blockDim.x) tw[i] = tw_n[i];
  __syncthreads();
  for (int f = warp; f < nframes; f += nwarps) {
    for (int m = lane; m < size; m += 32) xw[m] = tile[f * shift + m] * win_full[m];
    __syncwarp();
    float2 *row = spec + (static_cast<int64_t>(b) * T + t0 + f) * F;
    for (int k = lane; k < F; k += 32) {
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int m = 0; m < size; ++m) {
        const float2 w = tw[idx];
        re = fmaf(xw[m], w.x, re);
        im = fmaf(xw[m], w.y, im);
        idx += k;
        if (idx >= size) idx -= size;
      }
      row[k] = make_float2(re, im);
    }
    __syncwarp();
  }
}

// feats [B, T, 2F] = |X| || angle X ; labels [B, T, C*F] = |S_c| cos(angle X - angle S_c).
__global__ void features_kernel(const float *__restrict__ mix, const float *__restrict__ refs,
                                int n_src, int64_t n, int T, int size, int shift, int pad,
                                int frames_per_tile, const float *__restrict__ win_half,
                                const float2 *__restrict__ tw_half,
                                const float2 *__restrict__ tw_full, float *__restrict__ feats,
                                float *__restrict__ labels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = size >> 1, F = M + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * frames_per_tile;
  const int nframes = min(frames_per_tile, T - t0);
  const int tile_cap = ((frames_per_tile - 1) * shift + size + 3) & ~3;
  const int tile_len = (nframes - 1) * shift + size;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float *tiles = reinterpret_cast<float *>(smem_raw);
  float2 *bufs = reinterpret_cast<float2 *>(tiles + static_cast<size_t>(tile_cap) * (1 + n_src));
  float2 *A = bufs + static_cast<size_t>(warp) * (2 * M + F), *B = A + M, *XB = B + M;
  const int64_t s0 = static_cast<int64_t>(t0) * shift - pad;

  stage_wave(tiles, mix + static_cast<int64_t>(b) * n, n, s0, tile_len);
  for (int c = 0; c < n_src; ++c)
    stage_wave(tiles + static_cast<size_t>(tile_cap) * (1 + c),
               refs + (static_cast<int64_t>(b) * n_src + c) * n, n, s0, tile_len);
  __syncthreads();

  for (int f = warp; f < nframes; f += nwarps) {
    const int64_t frame = static_cast<int64_t>(b) * T + t0 + f;
    load_frame_packed(A, tiles + f * shift, win_half, M, lane);
    const float2 *Z = warp_fft<false>(A, B, tw_half, M, lane);
    for (int k = lane; k <= M; k += 32) {
      const float2 x = real_split(Z, tw_full, M, k);
      XB[k] = x;
      if (feats) {
        feats[frame * 2 * F + k] = hypotf(x.x, x.y);          // np.abs, :262
        feats[frame * 2 * F + F + k] = atan2f(x.y, x.x);      // np.angle, :263
      }
    }
    __syncwarp();
    if (labels) {
      for (int c = 0; c < n_src; ++c) {
        load_frame_packed(A, tiles + static_cast<size_t>(tile_cap) * (1 + c) + f * shift,
                          win_half, M, lane);
        const float2 *Zs = warp_fft<false>(A, B, tw_half, M, lane);
        for (int k = lane; k <= M; k += 32) {
          const float2 s = real_split(Zs, tw_full, M, k), x = XB[k];
          const float mag = hypotf(x.x, x.y);
          // |S| cos(angle X - angle S) = Re(S conj X) / |X|; angle(0) = 0 -> Re S
          const float lab = mag > 0.f ? fmaf(s.x, x.x, s.y * x.y) / mag : s.x;
          labels[(frame * n_src + c) * F + k] = lab;
        }
        __syncwarp();
      }
    }
  }
}

__global__ void segment_axis_kernel(const float *__restrict__ in, int64_t n, int frames, int length,
                                    int hop, float *__restrict__ out) {
  const int b = blockIdx.y;
  const int64_t total = static_cast<int64_t>(frames) * length;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t f = i / length, j = i - f * length;
    out[b * total + i] = __ldg(in + b * n + f * hop + j);
  }
}

}  // namespace sep

using namespace sep;

extern "C" {

int sep_segment_axis_f32(const float *in, int batch, int64_t n, int length, int overlap,
                         float *out, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(in && out, "sep_segment_axis_f32: null buffer");
  SEP_REQUIRE(batch >= 1 && length >= 1 && overlap >= 0 && overlap < length && n >= length,
              "sep_segment_axis_f32: bad geometry (batch=%d n=%lld length=%d overlap=%d)", batch,
              (long long)n, length, overlap);
  const int hop = length - overlap;
  SEP_REQUIRE((n - length) % hop == 0,
              "sep_segment_axis_f32: (n - length) must be a multiple of the hop");
  int rc = check_mem(mem);
  if (rc) return rc;
  const int64_t frames = 1 + (n - length) / hop;
  Scratch s(stream);
  const float *d_in;
  float *d_out;
  if ((rc = stage_in(s, in, static_cast<size_t>(batch) * n, mem, &d_in))) return rc;
  if ((rc = stage_out(s, out, static_cast<size_t>(batch) * frames * length, mem, &d_out))) return rc;
  const int64_t total = frames * length;
  dim3 grid(static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 4096)), batch);
  segment_axis_kernel<<<grid, 256, 0, stream>>>(d_in, n, static_cast<int>(frames), length, hop, d_out);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, out, d_out, static_cast<size_t>(batch) * total, mem))) return rc;
  return finish(s, mem);
}

int sep_stft_f32(const sep_plan *p, const float *wave, int batch, int64_t n_samples,
                 int64_t wave_stride, float *spec, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(p && wave && spec, "sep_stft_f32: null argument");
  SEP_REQUIRE(batch >= 1 && n_samples >= 0 && wave_stride >= n_samples,
              "sep_stft_f32: bad shape (batch=%d n=%lld stride=%lld)", batch, (long long)n_samples,
              (long long)wave_stride);
  int rc = check_mem(mem);
  if (rc) return rc;
  int T = 0;
  sep_plan_frames(p, n_samples, &T);
  if (T == 0) return SEP_OK;  // nothing to write: the reference returns a [0, F] array
  Scratch s(stream);
  const float *d_wave;
  float *d_spec;
  if ((rc = stage_in(s, wave, static_cast<size_t>(batch) * wave_stride, mem, &d_wave))) return rc;
  const size_t out_count = static_cast<size_t>(batch) * T * p->bins * 2;
  if ((rc = stage_out(s, spec, out_count, mem, &d_spec))) return rc;
  if (!p->pow2) {
    const int warps = 4, frames = 8;
    const size_t smem = (((frames - 1) * static_cast<size_t>(p->shift) + p->size + 3) & ~size_t(3)) * sizeof(float) +
                        p->size * sizeof(float2) + static_cast<size_t>(warps) * p->size * sizeof(float) + 16;
    SEP_CUDA(cudaFuncSetAttribute(stft_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid((T + frames - 1) / frames, batch);
    profile_begin(stream, "stft_dft_kernel (direct DFT, size=%d)", p->size);
    stft_dft_kernel<<<grid, warps * 32, smem, stream>>>(d_wave, n_samples, wave_stride, T, p->size, p->shift, p->pad,
                                                        frames, p->d_win_full, p->d_tw_n,
                                                        reinterpret_cast<float2 *>(d_spec));
    profile_end(stream);
    SEP_LAUNCHED();
    if ((rc = copy_back(s, spec, d_spec, out_count, mem))) return rc;
    return finish(s, mem);
  }
  const TileCfg cfg = pick_tile(p, 1, 0);
  SEP_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(cfg.smem + 16)));
  dim3 grid((T + cfg.frames - 1) / cfg.frames, batch);
  stft_kernel<<<grid, cfg.warps * 32, cfg.smem + 16, stream>>>(
      d_wave, n_samples, wave_stride, T, p->size, p->shift, p->pad, cfg.frames, p->d_win_half,
      p->d_tw_half, p->d_tw_full, reinterpret_cast<float2 *>(d_spec));
  SEP_LAUNCHED();
  if ((rc = copy_back(s, spec, d_spec, out_count, mem))) return rc;
  return finish(s, mem);
}

int sep_stft_features_f32(const sep_plan *p, const float *mix, const float *refs, int batch,
                          int n_src, int64_t n_samples, float *feats, float *labels, int mem,
                          void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(p && mix, "sep_stft_features_f32: null argument");
  SEP_REQUIRE(n_src >= 0 && n_src <= SEP_MAX_SOURCES, "sep_stft_features_f32: n_src=%d out of range",
              n_src);
  SEP_REQUIRE(n_src == 0 || refs != nullptr, "sep_stft_features_f32: refs required when n_src > 0");
  SEP_REQUIRE(labels == nullptr || n_src > 0, "sep_stft_features_f32: labels need n_src > 0");
  SEP_REQUIRE(batch >= 1 && n_samples >= 0, "sep_stft_features_f32: bad shape");
  if (!p->pow2) {
    set_error("sep_stft_features_f32: size=%d is not a power of two >= 32 (only stft / istft take other sizes)", p->size);
    return SEP_ERR_UNSUPPORTED;
  }
  int rc = check_mem(mem);
  if (rc) return rc;
  int T = 0;
  sep_plan_frames(p, n_samples, &T);
  if (T == 0 || (!feats && !labels)) return SEP_OK;
  const int src = labels ? n_src : 0;
  Scratch s(stream);
  const float *d_mix, *d_refs;
  float *d_feats, *d_labels;
  if ((rc = stage_in(s, mix, static_cast<size_t>(batch) * n_samples, mem, &d_mix))) return rc;
  if ((rc = stage_in(s, src ? refs : nullptr, static_cast<size_t>(batch) * src * n_samples, mem,
                     &d_refs)))
    return rc;
  const size_t nf = static_cast<size_t>(batch) * T * 2 * p->bins;
  const size_t nl = static_cast<size_t>(batch) * T * src * p->bins;
  if ((rc = stage_out(s, feats, nf, mem, &d_feats))) return rc;
  if ((rc = stage_out(s, labels, nl, mem, &d_labels))) return rc;
  const TileCfg cfg = pick_tile(p, 1 + src, p->bins);
  SEP_CUDA(cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(cfg.smem + 64)));
  dim3 grid((T + cfg.frames - 1) / cfg.frames, batch);
  features_kernel<<<grid, cfg.warps * 32, cfg.smem + 64, stream>>>(
      d_mix, d_refs, src, n_samples, T, p->size, p->shift, p->pad, cfg.frames, p->d_win_half,
      p->d_tw_half, p->d_tw_full, d_feats, d_labels);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, feats, d_feats, nf, mem))) return rc;
  if ((rc = copy_back(s, labels, d_labels, nl, mem))) return rc;
  return finish(s, mem);
}

}  // extern "C"
