// pcm.cu -- the sample-format steps either side of the path:
//   * int16 PCM -> float32 (what wavread / librosa.load give for the reference's 16-bit wavs:
//     metrics/evaluate_metrics.py:7-12, parallel_stft.py:213): x / 32768
//   * audiowrite's float -> int16 conversion (uPIT_baseline.ipynb:1317-1354, cell 40), per row:
//       if normalize: data /= max(|data|)          (float32 division)
//       data *= 32767                              (float32 product)
//       clipped = count(data > 32767);  data = clip(data, -32768, 32767);  int16 = trunc(data)
// Both are HBM-bound streams: 6 bytes per sample (+ 4 for the max pass when normalising).
#include <algorithm>

#include "common.cuh"

namespace sep {

__global__ void pcm16_to_f32_kernel(const int16_t *__restrict__ pcm, int64_t n, float *__restrict__ out) {
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i0 + 8 <= n && ((reinterpret_cast<uintptr_t>(pcm) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(pcm + i0));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = static_cast<float>(static_cast<int16_t>(w[k] & 0xFFFFu)) * (1.f / 32768.f);
      v[2 * k + 1] = static_cast<float>(static_cast<int16_t>(w[k] >> 16)) * (1.f / 32768.f);
    }
    reinterpret_cast<float4 *>(out + i0)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(out + i0)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    for (int64_t i = i0; i < min(i0 + 8, n); ++i) out[i] = static_cast<float>(pcm[i]) * (1.f / 32768.f);
  }
}

// |x| as an unsigned integer is monotonic in |x| (NaN sorts above everything, like numpy's max propagates it)
template <typename T> struct AbsBits;
template <> struct AbsBits<float> {
  using U = unsigned;
  static __device__ __forceinline__ U of(float v) { return __float_as_uint(v) & 0x7FFFFFFFu; }
  static __device__ __forceinline__ float back(U u) { return __uint_as_float(u); }
};
template <> struct AbsBits<double> {
  using U = unsigned long long;
  static __device__ __forceinline__ U of(double v) { return static_cast<U>(__double_as_longlong(v)) & 0x7FFFFFFFFFFFFFFFull; }
  static __device__ __forceinline__ double back(U u) { return __longlong_as_double(static_cast<long long>(u)); }
};

template <typename T, bool VEC>
__global__ void row_absmax_kernel(const T *__restrict__ data, int64_t n, typename AbsBits<T>::U *__restrict__ rowmax) {
  using B = AbsBits<T>;
  const int row = blockIdx.y;
  const T *src = data + static_cast<int64_t>(row) * n;
  typename B::U m = 0;
  constexpr int V = VEC ? 16 / sizeof(T) : 1;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * V; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * V) {
    if (VEC) {
      const float4 raw = __ldg(reinterpret_cast<const float4 *>(src + i));
      const T *v = reinterpret_cast<const T *>(&raw);
#pragma unroll
      for (int k = 0; k < V; ++k) m = max(m, B::of(v[k]));
    } else {
      m = max(m, B::of(__ldg(src + i)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(rowmax + row, m);
}

// One sample of cell 40 in the input's own precision (float32 data stays float32, float64 -- what the
// reference's istft returns -- stays float64), numpy's order of operations.  NaN (an all-zero row
// normalised by its zero peak, :1337) goes through np.clip unchanged and astype(np.int16) makes it 0.
template <typename T>
__device__ __forceinline__ int16_t audiowrite_one(T v, bool norm, T peak, unsigned &over) {
  if (norm) v = v / peak;                            // data /= np.max(np.abs(data)), :1337 (IEEE division)
  v = v * static_cast<T>(32767);                     // data *= int16_max, :1340
  over += v > static_cast<T>(32767);                 // :1342
  if (v != v) return 0;                              // NaN.astype(np.int16) == 0
  v = v < static_cast<T>(-32768) ? static_cast<T>(-32768) : (v > static_cast<T>(32767) ? static_cast<T>(32767) : v);   // np.clip, :1345
  return static_cast<int16_t>(static_cast<int>(v));  // astype(np.int16): toward zero
}

// VEC: n % 8 == 0 and 16-byte aligned rows -> 16-byte loads, one 16-byte store of eight int16
template <typename T, bool VEC>
__global__ void audiowrite_kernel(const T *__restrict__ data, int64_t n, const typename AbsBits<T>::U *__restrict__ rowmax,
                                  int16_t *__restrict__ out, unsigned long long *__restrict__ clipped) {
  const int row = blockIdx.y;
  const T *src = data + static_cast<int64_t>(row) * n;
  int16_t *dst = out + static_cast<int64_t>(row) * n;
  const bool norm = rowmax != nullptr;
  const T peak = norm ? AbsBits<T>::back(rowmax[row]) : static_cast<T>(1);
  unsigned over = 0;
  constexpr int V = VEC ? 8 : 1;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * V; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * V) {
    if (VEC) {
      constexpr int NL = 8 * sizeof(T) / 16;          // 16-byte loads for eight samples
      float4 raw[NL];
#pragma unroll
      for (int k = 0; k < NL; ++k) raw[k] = __ldg(reinterpret_cast<const float4 *>(src + i) + k);
      const T *v = reinterpret_cast<const T *>(raw);
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = static_cast<uint16_t>(audiowrite_one<T>(v[2 * k], norm, peak, over));
        const uint32_t hi = static_cast<uint16_t>(audiowrite_one<T>(v[2 * k + 1], norm, peak, over));
        w[k] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      dst[i] = audiowrite_one<T>(__ldg(src + i), norm, peak, over);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
  if ((threadIdx.x & 31) == 0 && over) atomicAdd(clipped + row, static_cast<unsigned long long>(over));
}

}  // namespace sep

using namespace sep;

extern "C" int sep_pcm16_to_f32(const int16_t *pcm, int64_t n, float *out, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(pcm && out && n >= 0, "sep_pcm16_to_f32: bad argument");
  int rc = check_mem(mem);
  if (rc) return rc;
  if (n == 0) return SEP_OK;
  Scratch s(stream);
  const int16_t *d_in;
  float *d_out;
  if ((rc = stage_in(s, pcm, static_cast<size_t>(n), mem, &d_in))) return rc;
  if ((rc = stage_out(s, out, static_cast<size_t>(n), mem, &d_out))) return rc;
  const int64_t threads = (n + 7) / 8;
  profile_begin(stream, "pcm16_to_f32_kernel");
  pcm16_to_f32_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(d_in, n, d_out);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, out, d_out, static_cast<size_t>(n), mem))) return rc;
  return finish(s, mem);
}

template <typename T>
static int audiowrite_impl(const T *data, int batch, int64_t n, int normalize, int16_t *out, int64_t *clipped, int mem,
                           cudaStream_t stream) {
  using U = typename AbsBits<T>::U;
  SEP_REQUIRE(data && out && batch >= 1 && n >= 1, "sep_audiowrite_i16: bad argument");
  int rc = check_mem(mem);
  if (rc) return rc;
  Scratch s(stream);
  const size_t count = static_cast<size_t>(batch) * n;
  const T *d_in;
  int16_t *d_out;
  unsigned long long *d_clip;
  U *d_max = nullptr;
  if ((rc = stage_in(s, data, count, mem, &d_in))) return rc;
  if ((rc = stage_out(s, out, count, mem, &d_out))) return rc;
  if ((rc = s.alloc(&d_clip, static_cast<size_t>(batch)))) return rc;
  SEP_CUDA(cudaMemsetAsync(d_clip, 0, sizeof(unsigned long long) * batch, stream));
  const bool vec = n % 8 == 0 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
  // enough CTAs to fill the machine a few times over, whatever the batch
  const int64_t per_block = 256 * (vec ? 8 : 1) * 4;
  const unsigned bx = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((n + per_block - 1) / per_block,
                                                                             (148 * 32 + batch - 1) / batch)));
  const dim3 grid(bx, batch);
  profile_begin(stream, "row_absmax_kernel + audiowrite_kernel<%s>", sizeof(T) == 4 ? "float" : "double");
  if (normalize) {
    if ((rc = s.alloc(&d_max, static_cast<size_t>(batch)))) return rc;
    SEP_CUDA(cudaMemsetAsync(d_max, 0, sizeof(U) * batch, stream));
    if (vec) row_absmax_kernel<T, true><<<grid, 256, 0, stream>>>(d_in, n, d_max);
    else row_absmax_kernel<T, false><<<grid, 256, 0, stream>>>(d_in, n, d_max);
    SEP_LAUNCHED();
  }
  if (vec) audiowrite_kernel<T, true><<<grid, 256, 0, stream>>>(d_in, n, d_max, d_out, d_clip);
  else audiowrite_kernel<T, false><<<grid, 256, 0, stream>>>(d_in, n, d_max, d_out, d_clip);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, out, d_out, count, mem))) return rc;
  if (clipped) {
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "clipped counter width");
    if (mem == SEP_MEM_HOST) {
      SEP_CUDA(cudaMemcpyAsync(clipped, d_clip, sizeof(int64_t) * batch, cudaMemcpyDeviceToHost, stream));
    } else {
      SEP_CUDA(cudaMemcpyAsync(clipped, d_clip, sizeof(int64_t) * batch, cudaMemcpyDeviceToDevice, stream));
    }
  }
  return finish(s, mem);
}

extern "C" int sep_audiowrite_i16_f32(const float *data, int batch, int64_t n, int normalize, int16_t *out,
                                      int64_t *clipped, int mem, void *stream_) {
  return audiowrite_impl<float>(data, batch, n, normalize, out, clipped, mem, static_cast<cudaStream_t>(stream_));
}

extern "C" int sep_audiowrite_i16_f64(const double *data, int batch, int64_t n, int normalize, int16_t *out,
                                      int64_t *clipped, int mem, void *stream_) {
  return audiowrite_impl<double>(data, batch, n, normalize, out, clipped, mem, static_cast<cudaStream_t>(stream_));
}
