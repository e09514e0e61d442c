// plan.cu -- library state, error reporting and the immutable STFT plan.
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

namespace sep {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_error = buf;
}

// Optional per-kernel timing: when enabled, entry points bracket their dominant
// kernel with CUDA events on the launching stream (bench.py's roofline leg).
static std::mutex g_prof_mutex;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_events;   // begin/end pairs
static size_t g_prof_used = 0;

static void profile_mark(cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof_on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  if (g_prof_used == g_prof_events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof_events.push_back(e);
  }
  cudaEventRecord(g_prof_events[g_prof_used++], stream);
}
static thread_local char t_kernel[160] = "";
void profile_begin(cudaStream_t stream, const char *fmt, ...) {
  if (fmt) {                                     // name of the dominant kernel this entry point is about to launch
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_kernel, sizeof(t_kernel), fmt, ap);
    va_end(ap);
  }
  profile_mark(stream);
}
void profile_end(cudaStream_t stream) { profile_mark(stream); }

template <typename T>
static int upload(T **dst, const std::vector<T> &host) {
  SEP_CUDA(cudaMalloc(reinterpret_cast<void **>(dst), host.size() * sizeof(T)));
  SEP_CUDA(cudaMemcpy(*dst, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return SEP_OK;
}

}  // namespace sep

using namespace sep;

extern "C" {

int sep_version(void) { return 100; }

const char *sep_last_error(void) { return t_error.c_str(); }

const char *sep_last_kernel(void) { return t_kernel; }

int64_t sep_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- peer-visible device memory for the one-sided per-batch reduction (sep_fused_separate_push_f32) ----
int sep_peer_alloc(int64_t bytes, void **ptr, unsigned char *handle64) {
  SEP_REQUIRE(ptr && handle64 && bytes > 0, "sep_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SEP_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));          // a dedicated allocation: IPC handles name whole allocations
  SEP_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  SEP_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return SEP_OK;
}

int sep_peer_open(const unsigned char *handle64, void **ptr) {
  SEP_REQUIRE(ptr && handle64, "sep_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SEP_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));   // maps the peer's memory (NVLink P2P)
  return SEP_OK;
}

int sep_peer_close(void *ptr) {
  if (ptr) SEP_CUDA(cudaIpcCloseMemHandle(ptr));
  return SEP_OK;
}

int sep_peer_free(void *ptr) {
  if (ptr) SEP_CUDA(cudaFree(ptr));
  return SEP_OK;
}

int sep_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_prof_on = on != 0;
  g_prof_used = 0;
  return SEP_OK;
}

int sep_profile_collect(double *total_ms, int *launches) {
  SEP_REQUIRE(total_ms && launches, "sep_profile_collect: null argument");
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  double total = 0.0;
  int count = 0;
  for (size_t i = 0; i + 1 < g_prof_used; i += 2) {
    SEP_CUDA(cudaEventSynchronize(g_prof_events[i + 1]));
    float ms = 0.f;
    SEP_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]));
    total += ms;
    ++count;
  }
  *total_ms = total;
  *launches = count;
  g_prof_used = 0;
  return SEP_OK;
}

int sep_plan_create(sep_plan **out, int size, int shift, const double *window, int fading) {
  SEP_REQUIRE(out != nullptr && window != nullptr, "sep_plan_create: null argument");
  *out = nullptr;
  if (size < 2 || size > 4096) {
    set_error("sep_plan_create: size=%d unsupported (2 <= size <= 4096)", size);
    return SEP_ERR_UNSUPPORTED;
  }
  SEP_REQUIRE(shift >= 1 && shift <= size, "sep_plan_create: shift=%d must be in [1, size=%d]",
              shift, size);
  sep_plan *p = new (std::nothrow) sep_plan();
  if (!p) {
    set_error("sep_plan_create: out of host memory");
    return SEP_ERR_NOMEM;
  }
  p->size = size;
  // FFT kernels need a power of two >= 32; any other size (the reference's `size` is free, parallel_stft.py:146)
  // runs the direct-DFT kernels of stft / istft -- a compatibility path, O(size^2) per frame
  p->pow2 = size >= 32 && (size & (size - 1)) == 0 ? 1 : 0;
  p->shift = shift;
  p->half = size / 2;
  p->bins = size / 2 + 1;
  p->hops = (size % shift == 0) ? size / shift : 0;
  p->fading = fading ? 1 : 0;
  p->pad = fading ? size - shift : 0;
  p->window.assign(window, window + size);

  // Biorthogonal synthesis window, uPIT_baseline.ipynb:1234-1259 (cell 38):
  // q[m] = sum_k w[m + k*shift]^2 over taps with index + 1 < size (the last tap
  // never enters a sum); synth = w / q[n mod shift] / size.
  if (p->hops > 0) {
    std::vector<double> q(shift, 0.0);
    for (int m = 0; m < shift; ++m)
      for (int k = 0; k <= p->hops; ++k) {
        const int idx = m + k * shift;
        if (idx + 1 < size) q[m] += window[idx] * window[idx];
      }
    p->synth.resize(size);
    for (int n = 0; n < size; ++n) p->synth[n] = window[n] / q[n % shift] / size;
  }

  int rc = SEP_OK;
  do {
    cudaError_t e = cudaGetDevice(&p->device);
    if (e != cudaSuccess) {
      set_error("sep_plan_create: no usable CUDA device: %s", cudaGetErrorString(e));
      rc = SEP_ERR_CUDA;
      break;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, p->device);
    if (e != cudaSuccess) {
      set_error("sep_plan_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
      rc = SEP_ERR_CUDA;
      break;
    }
    p->sm_count = prop.multiProcessorCount;
    p->max_smem = static_cast<int>(prop.sharedMemPerBlockOptin);
    // keep stream-ordered scratch cached between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, p->device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }

    const int n = size, M = p->half;
    std::vector<float> wh(n), wf(n), ws(n, 0.f);
    for (int i = 0; i < n; ++i) {
      wh[i] = static_cast<float>(0.5 * window[i]);
      wf[i] = static_cast<float>(window[i]);
      // istft multiplies synth by size (cell 39 :1297) and numpy's irfft divides
      // by size; the two cancel, so the kernels run UNNORMALISED inverse
      // transforms and multiply by synth itself.
      if (p->hops > 0) ws[i] = static_cast<float>(p->synth[i]);
    }
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float2> th(M), tf(M + 1), t16(256);
    for (int k = 0; k < M; ++k)
      th[k] = make_float2(static_cast<float>(std::cos(two_pi * k / M)),
                          static_cast<float>(-std::sin(two_pi * k / M)));
    for (int k = 0; k <= M; ++k)
      tf[k] = make_float2(static_cast<float>(std::cos(two_pi * k / n)),
                          static_cast<float>(-std::sin(two_pi * k / n)));
    for (int a = 0; a < 16; ++a)
      for (int b = 0; b < 16; ++b)
        t16[a * 16 + b] = make_float2(static_cast<float>(std::cos(two_pi * a * b / 256.0)),
                                      static_cast<float>(-std::sin(two_pi * a * b / 256.0)));
    if ((rc = upload(&p->d_win_half, wh)) != SEP_OK) break;
    if ((rc = upload(&p->d_win_full, wf)) != SEP_OK) break;
    if ((rc = upload(&p->d_syn, ws)) != SEP_OK) break;
    if ((rc = upload(&p->d_tw_half, th)) != SEP_OK) break;
    if ((rc = upload(&p->d_tw_full, tf)) != SEP_OK) break;
    if ((rc = upload(&p->d_tw16, t16)) != SEP_OK) break;
    if (!p->pow2) {
      std::vector<float2> tn(n);
      for (int k = 0; k < n; ++k)
        tn[k] = make_float2(static_cast<float>(std::cos(two_pi * k / n)), static_cast<float>(-std::sin(two_pi * k / n)));
      if ((rc = upload(&p->d_tw_n, tn)) != SEP_OK) break;
    }
    if (size == 256) {
      std::vector<float> wt(16 * 18 + 8, 0.f), st(16 * 18 + 8, 0.f);
      for (int lane = 0; lane < 16; ++lane)
        for (int m = 0; m < 16; ++m) {
          wt[lane * 18 + m] = wh[lane + 16 * m];
          st[lane * 18 + m] = ws[lane + 16 * m];
        }
      if ((rc = upload(&p->d_win_t, wt)) != SEP_OK) break;
      if ((rc = upload(&p->d_syn_t, st)) != SEP_OK) break;
    }
    if (size == 512) {
      std::vector<float2> w2(16 * 18, make_float2(0.f, 0.f)), s2(16 * 18, make_float2(0.f, 0.f)),
          t2(16 * 18, make_float2(0.f, 0.f));
      for (int lane = 0; lane < 16; ++lane)
        for (int m = 0; m < 16; ++m) {
          const int i0 = 2 * lane + 32 * m;
          w2[lane * 18 + m] = make_float2(wh[i0], wh[i0 + 1]);
          s2[lane * 18 + m] = make_float2(ws[i0], ws[i0 + 1]);
          const int k = lane + 16 * m;
          t2[lane * 18 + m] = make_float2(static_cast<float>(std::cos(two_pi * k / 512.0)),
                                          static_cast<float>(-std::sin(two_pi * k / 512.0)));
        }
      if ((rc = upload(&p->d_win2_t, w2)) != SEP_OK) break;
      if ((rc = upload(&p->d_syn2_t, s2)) != SEP_OK) break;
      if ((rc = upload(&p->d_tw512_t, t2)) != SEP_OK) break;
    }
  } while (0);
  if (rc != SEP_OK) {
    sep_plan_destroy(p);
    return rc;
  }
  *out = p;
  return SEP_OK;
}

int sep_plan_destroy(sep_plan *p) {
  if (!p) return SEP_OK;
  cudaFree(p->d_win_half);
  cudaFree(p->d_win_full);
  cudaFree(p->d_syn);
  cudaFree(p->d_tw_half);
  cudaFree(p->d_tw_full);
  cudaFree(p->d_tw16);
  cudaFree(p->d_tw_n);
  cudaFree(p->d_win_t);
  cudaFree(p->d_syn_t);
  cudaFree(p->d_win2_t);
  cudaFree(p->d_syn2_t);
  cudaFree(p->d_tw512_t);
  delete p;
  return SEP_OK;
}

int sep_plan_frames(const sep_plan *p, int64_t n_samples, int *frames) {
  SEP_REQUIRE(p && frames, "sep_plan_frames: null argument");
  SEP_REQUIRE(n_samples >= 0, "sep_plan_frames: negative sample count");
  // parallel_stft.py:169-177: n1 = n + 2*pad; T = ceil((n1 - size + shift) / shift)
  const int64_t n1 = n_samples + 2 * static_cast<int64_t>(p->pad);
  const int64_t num = n1 - p->size + p->shift;
  int64_t t = num <= 0 ? 0 : (num + p->shift - 1) / p->shift;
  *frames = static_cast<int>(t);
  return SEP_OK;
}

int sep_plan_istft_samples(const sep_plan *p, int frames, int64_t *n_samples) {
  SEP_REQUIRE(p && n_samples, "sep_plan_istft_samples: null argument");
  SEP_REQUIRE(frames >= 0, "sep_plan_istft_samples: negative frame count");
  // cell 39 :1298-1305: T*shift + size - shift, minus the fade padding both ends
  int64_t n = static_cast<int64_t>(frames) * p->shift + p->size - p->shift - 2 * p->pad;
  *n_samples = n < 0 ? 0 : n;
  return SEP_OK;
}

int sep_plan_synthesis_window(const sep_plan *p, double *out) {
  SEP_REQUIRE(p && out, "sep_plan_synthesis_window: null argument");
  SEP_REQUIRE(p->hops > 0, "synthesis window needs size %% shift == 0 (size=%d shift=%d)",
              p->size, p->shift);
  std::memcpy(out, p->synth.data(), sizeof(double) * p->size);
  return SEP_OK;
}

int sep_score_stride(int n_src) {
  if (n_src < 1 || n_src > SEP_MAX_SOURCES) return SEP_ERR_INVALID;
  return 3 * n_src * n_src + factorial(n_src) + 6;
}

}  // extern "C"
