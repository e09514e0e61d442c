// common.cuh -- shared host/device plumbing of libsepcore (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/sepcore.h"

namespace sep {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

#define SEP_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::sep::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                       cudaGetErrorString(e__));                                    \
      return SEP_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define SEP_REQUIRE(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::sep::set_error(__VA_ARGS__);                                                \
      return SEP_ERR_INVALID;                                                       \
    }                                                                               \
  } while (0)

void profile_begin(cudaStream_t stream, const char *fmt = nullptr, ...);
void profile_end(cudaStream_t stream);

// counts one kernel launch and checks the launch error
#define SEP_LAUNCHED()                                                              \
  do {                                                                              \
    ::sep::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    SEP_CUDA(cudaGetLastError());                                                   \
  } while (0)

}  // namespace sep

// ------------------------------------------------------------------ plan
struct sep_plan {
  int size = 0;     // n_fft
  int shift = 0;    // hop
  int bins = 0;     // size/2 + 1
  int half = 0;     // size/2 (length of the complex FFT behind a real one)
  int hops = 0;     // size/shift when divisible, else 0 (no iSTFT / fused path)
  int pow2 = 1;     // size is a power of two >= 32 (FFT kernels); otherwise only stft / istft run (direct DFT)
  int fading = 1;
  int pad = 0;      // size - shift when fading
  int device = 0;
  int sm_count = 148;
  int max_smem = 0;  // opt-in dynamic shared memory per block
  std::vector<double> window;  // analysis, float64 (host)
  std::vector<double> synth;   // biorthogonal synthesis window (cell 38), float64
  // device tables (float32)
  float *d_win_half = nullptr;   // 0.5 * analysis window   (half-size real FFT)
  float *d_win_full = nullptr;   // analysis window          (pair-packed FFT)
  float *d_syn = nullptr;        // size * synth / size = w/q, times 1/size (irfft norm)
  float2 *d_tw_half = nullptr;   // exp(-2 pi i k / half), k < half
  float2 *d_tw_full = nullptr;   // exp(-2 pi i k / size), k <= half
  float2 *d_tw16 = nullptr;      // exp(-2 pi i p*k / 256) [16][16] for the 16x16 FFT
  float2 *d_tw_n = nullptr;      // exp(-2 pi i k / size), k < size: direct-DFT path of a non-power-of-two size
  // size 256 only: windows as [16 lanes][18] (lane p holds taps p + 16 m at [p*18 + m]; the
  // pitch of 18 makes a half-warp's paired loads conflict-free)
  float *d_win_t = nullptr;      // 0.5 * analysis
  float *d_syn_t = nullptr;      // synthesis
  // size 512 only (half-size real transform on the 16x16 core): [16 lanes][18] float2, lane p entry m
  float2 *d_win2_t = nullptr;    // 0.5 * (w[2p + 32m], w[2p + 32m + 1])
  float2 *d_syn2_t = nullptr;    // (s[2p + 32m], s[2p + 32m + 1])
  float2 *d_tw512_t = nullptr;   // exp(-2 pi i (p + 16 m) / 512)
};

namespace sep {

// Stream-ordered scratch: every entry point allocates what it needs on its own
// stream, so calls are re-entrant.  The pool keeps freed blocks (threshold
// raised at plan creation), so steady-state calls do not hit the driver.
struct Scratch {
  cudaStream_t stream;
  std::vector<void *> blocks;
  // optional caller-provided arena (no allocation inside the call: CUDA graphs)
  unsigned char *arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  explicit Scratch(cudaStream_t s) : stream(s) { keep_pool_cached(); }
  // Stream-ordered frees go back to the device's default pool, not to the driver: without this every call
  // pays hundreds of microseconds of cudaMallocAsync (set once per process and device).
  static void keep_pool_cached() {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
  }
  ~Scratch() {
    for (void *p : blocks) cudaFreeAsync(p, stream);
  }
  void use_arena(void *base, size_t bytes) {
    arena = static_cast<unsigned char *>(base);
    arena_bytes = bytes;
    arena_used = 0;
  }
  template <typename T>
  int alloc(T **out, size_t count) {
    void *p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    if (arena) {
      const size_t start = (arena_used + 255) & ~size_t(255);
      if (start + bytes > arena_bytes) {
        set_error("workspace too small: need %zu bytes, have %zu", start + bytes, arena_bytes);
        return SEP_ERR_INVALID;
      }
      arena_used = start + bytes;
      *out = reinterpret_cast<T *>(arena + start);
      return SEP_OK;
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess) {
      set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? SEP_ERR_NOMEM : SEP_ERR_CUDA;
    }
    blocks.push_back(p);
    *out = static_cast<T *>(p);
    return SEP_OK;
  }
};

// Stages a caller buffer onto the device when it lives in host memory.
template <typename T>
inline int stage_in(Scratch &s, const T *src, size_t count, int mem, const T **dev) {
  if (src == nullptr) {
    *dev = nullptr;
    return SEP_OK;
  }
  if (mem == SEP_MEM_DEVICE) {
    *dev = src;
    return SEP_OK;
  }
  T *d = nullptr;
  int rc = s.alloc(&d, count);
  if (rc != SEP_OK) return rc;
  SEP_CUDA(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, s.stream));
  *dev = d;
  return SEP_OK;
}

template <typename T>
inline int stage_out(Scratch &s, T *dst, size_t count, int mem, T **dev) {
  if (dst == nullptr) {
    *dev = nullptr;
    return SEP_OK;
  }
  if (mem == SEP_MEM_DEVICE) {
    *dev = dst;
    return SEP_OK;
  }
  return s.alloc(dev, count);
}

template <typename T>
inline int copy_back(Scratch &s, T *dst, const T *dev, size_t count, int mem) {
  if (dst == nullptr || mem == SEP_MEM_DEVICE) return SEP_OK;
  SEP_CUDA(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, s.stream));
  return SEP_OK;
}

inline int finish(Scratch &s, int mem) {
  if (mem == SEP_MEM_HOST) SEP_CUDA(cudaStreamSynchronize(s.stream));
  return SEP_OK;
}

inline int check_mem(int mem) {
  if (mem != SEP_MEM_HOST && mem != SEP_MEM_DEVICE) {
    set_error("mem must be SEP_MEM_HOST or SEP_MEM_DEVICE, got %d", mem);
    return SEP_ERR_INVALID;
  }
  return SEP_OK;
}

inline int factorial(int n) {
  int f = 1;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Complex arithmetic on Blackwell's packed FP32 pipe: sm_100 FADD2 / FMUL2 / FFMA2
// work on (lo, hi) register pairs with per-operand half swizzles and per-half
// negation, so a complex add, an add with a +-i rotation or a conjugate costs ONE
// instruction and a complex multiply TWO (nvcc folds the make_float2 shuffles
// below into operand modifiers: FADD2 R, a.HI_LO, b.LO_HI.NP).
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) {   // a - i b
  return __fadd2_rn(a, make_float2(b.y, -b.x));
}
__device__ __forceinline__ float2 cadd_pi(float2 a, float2 b) {   // a + i b
  return __fadd2_rn(a, make_float2(-b.y, b.x));
}
__device__ __forceinline__ float2 cadd_conj(float2 a, float2 b) {  // a + conj(b)
  return __fadd2_rn(a, make_float2(b.x, -b.y));
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {      // a * b
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return __ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}

// MUFU.RSQ, one instruction (max relative error 2^-22); callers pass normal, positive arguments.
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of NV doubles per thread: warp butterflies, then
// warp 0 adds the per-warp partials in warp order.  `red` is shared scratch of
// at least NV * (blockDim.x / 32) doubles.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp * NV + i] = v[i];
  }
  __syncthreads();
  // warp 0: lane i adds value i over the warps, in warp order (deterministic)
  if (warp == 0) {
    for (int i = lane; i < NV; i += 32) {
      double s = 0.0;
      for (int w = 0; w < nwarp; ++w) s += red[w * NV + i];
      red[i] = s;     // safe: only lane i reads column i, and row 0 is read first
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = red[i];
  }
}

// ---- async copy helpers (sm_80+ LDGSTS, sm_90+ bulk prefetch) ----
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}
// Pulls [ptr, ptr + bytes) towards L2 without occupying registers or shared
// memory; ptr must be 16-byte aligned and bytes a multiple of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void *ptr, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// Lexicographic permutation p of n (n <= 4) into out[0..n).
__device__ __forceinline__ void nth_permutation(int n, int p, int *out) {
  int pool[SEP_MAX_SOURCES];
  for (int i = 0; i < n; ++i) pool[i] = i;
  int f = 1;
  for (int i = 2; i < n; ++i) f *= i;  // (n-1)!
  int left = n;
  for (int i = 0; i < n; ++i) {
    int q = p / f;
    p -= q * f;
    out[i] = pool[q];
    for (int k = q; k + 1 < left; ++k) pool[k] = pool[k + 1];
    --left;
    if (left > 1) f /= left;
  }
}

#endif  // __CUDACC__

// In-kernel finalisation counters ([batch + 1] ints) are zeroed before every launch -- also when they live in
// a caller workspace that the previous launch left at zero: a launch that was aborted half way must not poison
// every later call on that workspace.  The reset is a 4 (batch + 1)-byte memset node on the launch stream.
inline int reset_counters(int *counters, int batch, const Scratch &, cudaStream_t stream) {
  SEP_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * (static_cast<size_t>(batch) + 1), stream));
  return SEP_OK;
}

}  // namespace sep
