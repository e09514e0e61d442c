// istft.cu -- inverse STFT with atomic-free overlap-add, and the fused
// "cleaned magnitude x exp(j phase) -> waveform" back end.
//
// Reference: istft uPIT_baseline.ipynb:1269-1307 (cell 39); synthesis window
// :1234-1259 (cell 38, computed at plan creation); recombination
// spec_c = cleaned_c * exp(1j * angle) :1385-1388 (cell 41).
//
// A CTA owns a tile of consecutive output hop-blocks of one utterance.  It
// inverse-transforms every frame that touches the tile (TB + size/shift - 1
// frames: the halo on the left is recomputed, never exchanged), parks the
// windowed time frames in shared memory, and then each output sample sums its
// size/shift contributions in frame order -- no atomics, deterministic, every
// output sample written exactly once with coalesced stores.
#include "common.cuh"
#include "fft.cuh"

namespace sep {

struct OlaCfg {
  int warps;
  int blocks;   // output hop-blocks per tile (TB)
  size_t smem;
};

static bool pick_ola(const sep_plan *p, int n_src, int extra_f2_per_warp, OlaCfg *out) {
  const int R = p->hops;
  for (int warps = 8; warps >= 1; warps >>= 1) {
    for (int tb = 32; tb >= 1; tb >>= 1) {
      size_t frames = static_cast<size_t>(tb + R - 1) * p->size * sizeof(float) * n_src;
      size_t per_warp = (2 * static_cast<size_t>(p->half) + p->bins + extra_f2_per_warp) * sizeof(float2);
      size_t total = frames + per_warp * warps + 64;
      if (total <= 160 * 1024) {
        *out = OlaCfg{warps, tb, total};
        return true;
      }
    }
  }
  return false;
}

// MODE 0: spec [B, T, F] complex -> wave [B, L]          (n_src == 1)
// MODE 1: cleaned [B, T, C*F], phase [B, T, F] -> wave [B, C, L]
template <int MODE>
__global__ void istft_kernel(const float2 *__restrict__ spec, const float *__restrict__ cleaned,
                             const float *__restrict__ phase, int n_src, int T, int size, int shift,
                             int pad, int tb, int64_t out_len,
                             const float *__restrict__ syn, const float2 *__restrict__ tw_half,
                             const float2 *__restrict__ tw_full, float *__restrict__ wave) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = size >> 1, F = M + 1, R = size / shift;
  const int b = blockIdx.y;
  const int j_lo = pad / shift;                       // first hop-block that is kept
  const int j_hi = T + R - 1 - j_lo;                  // one past the last kept hop-block
  const int j0 = max(blockIdx.x * tb, j_lo), j1 = min((blockIdx.x + 1) * tb, j_hi);
  if (j0 >= j1) return;
  const int t_lo = max(j0 - R + 1, 0), t_hi = min(j1, T);   // frames [t_lo, t_hi) touch the tile
  const int nframes = t_hi - t_lo, slots = tb + R - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  float *fb = reinterpret_cast<float *>(smem_raw);   // [n_src][slots][size]
  float2 *bufs = reinterpret_cast<float2 *>(fb + static_cast<size_t>(n_src) * slots * size);
  const int per_warp = 2 * M + F + (MODE == 1 ? F : 0);
  float2 *A = bufs + static_cast<size_t>(warp) * per_warp, *B = A + M, *Y = B + M;
  float2 *CS = Y + F;  // MODE 1: (cos, sin) of the mixture phase

  for (int f = warp; f < nframes; f += nwarps) {
    const int64_t frame = static_cast<int64_t>(b) * T + t_lo + f;
    if (MODE == 1) {
      for (int k = lane; k <= M; k += 32) {
        float sn, cs;
        sincosf(__ldg(phase + frame * F + k), &sn, &cs);
        CS[k] = make_float2(cs, sn);
      }
    }
    for (int c = 0; c < n_src; ++c) {
      if (MODE == 0) {
        for (int k = lane; k <= M; k += 32) Y[k] = __ldg(spec + frame * F + k);
      } else {
        __syncwarp();
        for (int k = lane; k <= M; k += 32) {
          const float m = __ldg(cleaned + (frame * n_src + c) * F + k);
          Y[k] = make_float2(m * CS[k].x, m * CS[k].y);
        }
      }
      __syncwarp();
      for (int k = lane; k < M; k += 32) {
        float2 yk = Y[k], ym = Y[M - k];
        if (k == 0) { yk.y = 0.f; ym.y = 0.f; }   // irfft ignores Im of DC and Nyquist
        A[k] = real_merge(yk, ym, tw_full[k]);
      }
      __syncwarp();
      const float2 *z = warp_fft<true>(A, B, tw_half, M, lane);
      float2 *dst = reinterpret_cast<float2 *>(fb + (static_cast<size_t>(c) * slots + f) * size);
      for (int m = lane; m < M; m += 32) {
        const float2 w = reinterpret_cast<const float2 *>(syn)[m];
        dst[m] = make_float2(z[m].x * w.x, z[m].y * w.y);
      }
      __syncwarp();
    }
  }
  __syncthreads();

  const int span = (j1 - j0) * shift;
  for (int c = 0; c < n_src; ++c) {
    float *out = wave + (static_cast<int64_t>(b) * n_src + c) * out_len;
    const float *fbc = fb + static_cast<size_t>(c) * slots * size;
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
      const int j = j0 + i / shift, m = i % shift;
      float acc = 0.f;
      for (int t = max(j - R + 1, t_lo); t <= min(j, t_hi - 1); ++t)   // frame order, like :1300
        acc += fbc[(t - t_lo) * size + (j - t) * shift + m];
      out[static_cast<int64_t>(j) * shift + m - pad] = acc;
    }
  }
}

// Any size with size % shift == 0: inverse real DFT by direct summation (numpy's irfft: the imaginary parts of DC and,
// for even sizes, of the Nyquist bin are ignored), synthesis window, the same frame-ordered overlap-add.
__global__ void istft_dft_kernel(const float2 *__restrict__ spec, int T, int size, int shift, int pad, int tb,
                                 int64_t out_len, const float *__restrict__ syn, const float2 *__restrict__ tw_n,
                                 float *__restrict__ wave) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = size / 2 + 1, R = size / shift;
  const int b = blockIdx.y;
  const int j_lo = pad / shift, j_hi = T + R - 1 - j_lo;
  const int j0 = max(blockIdx.x * tb, j_lo), j1 = min((blockIdx.x + 1) * tb, j_hi);
  if (j0 >= j1) return;
  const int t_lo = max(j0 - R + 1, 0), t_hi = min(j1, T);
  const int nframes = t_hi - t_lo, slots = tb + R - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float *fb = reinterpret_cast<float *>(smem_raw);                         // [slots][size]
  float2 *tw = reinterpret_cast<float2 *>(fb + static_cast<size_t>(slots) * size);
  float2 *Y = tw + size + static_cast<size_t>(warp) * F;
  for (int i = threadIdx.x; i < size; i += blockDim.x) tw[i] = tw_n[i];
  __syncthreads();
  const bool even = (size & 1) == 0;
  const int kmax = even ? size / 2 - 1 : size / 2;                          // bins 1..kmax count twice
  for (int f = warp; f < nframes; f += nwarps) {
    const int64_t frame = static_cast<int64_t>(b) * T + t_lo + f;
    for (int k = lane; k < F; k += 32) Y[k] = __ldg(spec + frame * F + k);
    __syncwarp();
    for (int m = lane; m < size; m += 32) {
      float acc = Y[0].x;
      if (even) acc += (m & 1) ? -Y[size / 2].x : Y[size / 2].x;
      int idx = 0;
      for (int k = 1; k <= kmax; ++k) {
        idx += m;
        if (idx >= size) idx -= size;
        const float2 w = tw[idx];                                          // (cos, -sin) of 2 pi k m / n
        acc = fmaf(2.f * Y[k].x, w.x, acc);                                // Re(Y e^{+i theta}) = Yr cos - Yi sin
        acc = fmaf(2.f * Y[k].y, w.y, acc);
      }
      fb[static_cast<size_t>(f) * size + m] = acc * syn[m];
    }
    __syncwarp();
  }
  __syncthreads();
  const int span = (j1 - j0) * shift;
  float *out = wave + static_cast<int64_t>(b) * out_len;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int j = j0 + i / shift, m = i % shift;
    float acc = 0.f;
    for (int t = max(j - R + 1, t_lo); t <= min(j, t_hi - 1); ++t) acc += fb[(t - t_lo) * size + (j - t) * shift + m];
    out[static_cast<int64_t>(j) * shift + m - pad] = acc;
  }
}

template <int MODE>
static int launch_istft(const sep_plan *p, const float *in0, const float *in1, int batch, int n_src,
                        int frames, float *wave, int mem, cudaStream_t stream) {
  SEP_REQUIRE(p->hops > 0, "istft needs size %% shift == 0 (size=%d shift=%d)", p->size, p->shift);
  SEP_REQUIRE(batch >= 1 && frames >= 0, "istft: bad shape (batch=%d frames=%d)", batch, frames);
  int rc = check_mem(mem);
  if (rc) return rc;
  int64_t L = 0;
  sep_plan_istft_samples(p, frames, &L);
  if (L == 0 || frames == 0) return SEP_OK;
  if (!p->pow2) {
    if (MODE != 0) {
      set_error("recombine_istft: size=%d is not a power of two >= 32 (only stft / istft take other sizes)", p->size);
      return SEP_ERR_UNSUPPORTED;
    }
    const int warps = 4, tb = 8;
    Scratch s(stream);
    const float *d0;
    float *d_wave;
    const size_t TF = static_cast<size_t>(batch) * frames * p->bins;
    if ((rc = stage_in(s, in0, TF * 2, mem, &d0))) return rc;
    const size_t out_count = static_cast<size_t>(batch) * L;
    if ((rc = stage_out(s, wave, out_count, mem, &d_wave))) return rc;
    const size_t smem = static_cast<size_t>(tb + p->hops - 1) * p->size * sizeof(float) + p->size * sizeof(float2) +
                        static_cast<size_t>(warps) * p->bins * sizeof(float2) + 16;
    if (smem > 200 * 1024) {
      set_error("istft: size=%d shift=%d does not fit in shared memory (direct-DFT path)", p->size, p->shift);
      return SEP_ERR_UNSUPPORTED;
    }
    SEP_CUDA(cudaFuncSetAttribute(istft_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int j_hi = frames + p->hops - 1 - p->pad / p->shift;
    dim3 grid((j_hi + tb - 1) / tb, batch);
    profile_begin(stream, "istft_dft_kernel (direct inverse DFT, size=%d)", p->size);
    istft_dft_kernel<<<grid, warps * 32, smem, stream>>>(reinterpret_cast<const float2 *>(d0), frames, p->size, p->shift,
                                                         p->pad, tb, L, p->d_syn, p->d_tw_n, d_wave);
    profile_end(stream);
    SEP_LAUNCHED();
    if ((rc = copy_back(s, wave, d_wave, out_count, mem))) return rc;
    return finish(s, mem);
  }
  OlaCfg cfg;
  if (!pick_ola(p, n_src, MODE == 1 ? p->bins : 0, &cfg)) {
    set_error("istft: size=%d shift=%d sources=%d does not fit in shared memory", p->size, p->shift,
              n_src);
    return SEP_ERR_UNSUPPORTED;
  }
  Scratch s(stream);
  const float *d0, *d1;
  float *d_wave;
  const size_t TF = static_cast<size_t>(batch) * frames * p->bins;
  if (MODE == 0) {
    if ((rc = stage_in(s, in0, TF * 2, mem, &d0))) return rc;
    d1 = nullptr;
  } else {
    if ((rc = stage_in(s, in0, TF * n_src, mem, &d0))) return rc;
    if ((rc = stage_in(s, in1, TF, mem, &d1))) return rc;
  }
  const size_t out_count = static_cast<size_t>(batch) * n_src * L;
  if ((rc = stage_out(s, wave, out_count, mem, &d_wave))) return rc;
  SEP_CUDA(cudaFuncSetAttribute(istft_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(cfg.smem)));
  const int j_hi = frames + p->hops - 1 - p->pad / p->shift;
  dim3 grid((j_hi + cfg.blocks - 1) / cfg.blocks, batch);
  istft_kernel<MODE><<<grid, cfg.warps * 32, cfg.smem, stream>>>(
      MODE == 0 ? reinterpret_cast<const float2 *>(d0) : nullptr, MODE == 1 ? d0 : nullptr, d1,
      n_src, frames, p->size, p->shift, p->pad, cfg.blocks, L, p->d_syn, p->d_tw_half,
      p->d_tw_full, d_wave);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, wave, d_wave, out_count, mem))) return rc;
  return finish(s, mem);
}

}  // namespace sep

using namespace sep;

extern "C" {

int sep_istft_f32(const sep_plan *p, const float *spec, int batch, int frames, float *wave, int mem,
                  void *stream) {
  SEP_REQUIRE(p && spec && wave, "sep_istft_f32: null argument");
  return launch_istft<0>(p, spec, nullptr, batch, 1, frames, wave, mem,
                         static_cast<cudaStream_t>(stream));
}

int sep_recombine_istft_f32(const sep_plan *p, const float *cleaned, const float *phase, int batch,
                            int n_src, int frames, float *wave, int mem, void *stream) {
  SEP_REQUIRE(p && cleaned && phase && wave, "sep_recombine_istft_f32: null argument");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES, "sep_recombine_istft_f32: n_src=%d out of range",
              n_src);
  return launch_istft<1>(p, cleaned, phase, batch, n_src, frames, wave, mem,
                         static_cast<cudaStream_t>(stream));
}

}  // extern "C"
