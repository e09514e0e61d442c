// conv1d.cu -- Keras Conv1D forward ("learned filterbank" front end), fp32 SIMT.
//
// Reference: Conv1D(filters=129, kernel_size=2, activation='sigmoid',
// padding='same') on [B, K, 40] segments, Raw_with_Convlayer.ipynb:389 (cell
// 13); segmentation :83-100 (cell 2).
//
// out[b, r, n] = act(bias[n] + sum_{j, c} x[b, r*stride + j - left, c] * W[j, c, n]).
// Because the rows of x are contiguous, the im2col row of output r is the
// contiguous span xflat[(r*stride - left) * c_in ... + taps*c_in): the
// contraction is a [rows_out, taps*c_in] x [taps*c_in, filters] GEMM whose A
// operand is an overlapping strided view -- no im2col buffer is materialised.
// 64 x 64 output tile per CTA, 4 x 4 register micro-tile per thread, K staged
// through shared memory 16 at a time.  (The tcgen05 variant for the BASELINE
// N=256/L=16 shape is tracked in DESIGN.md.)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace sep {

constexpr int kTM = 64, kTN = 64, kTK = 16;

__device__ __forceinline__ float activate(float v, int act) {
  if (act == SEP_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == SEP_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__global__ void __launch_bounds__(256)
conv1d_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
              int rows, int c_in, int taps, int filters, int stride, int left, int rows_out,
              int act, float *__restrict__ out) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int b = blockIdx.z, r0 = blockIdx.x * kTM, n0 = blockIdx.y * kTN;
  const int K = taps * c_in;
  const float *xb = x + static_cast<int64_t>(b) * rows * c_in;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kTK) {
    // A tile: As[kk][r] = x[b, (r0+r)*stride - left + j, c], (j, c) = divmod(k0+kk, c_in)
    for (int e = threadIdx.x; e < kTK * kTM; e += 256) {
      const int kk = e % kTK, r = e / kTK;
      const int k = k0 + kk, ro = r0 + r;
      float v = 0.f;
      if (k < K && ro < rows_out) {
        const int j = k / c_in, c = k - j * c_in;
        const int row = ro * stride - left + j;
        if (row >= 0 && row < rows) v = __ldg(xb + static_cast<int64_t>(row) * c_in + c);
      }
      As[kk][r] = v;
    }
    for (int e = threadIdx.x; e < kTK * kTN; e += 256) {
      const int n = e % kTN, kk = e / kTN;
      const int k = k0 + kk;
      Bs[kk][n] = (k < K && n0 + n < filters) ? __ldg(w + static_cast<int64_t>(k) * filters + n0 + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ro = r0 + ty * 4 + i;
    if (ro >= rows_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= filters) continue;
      const float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      out[(static_cast<int64_t>(b) * rows_out + ro) * filters + n] = activate(v, act);
    }
  }
}

// ---- weights-resident variant for small contractions (the reference shape: K = taps * c_in = 80, 129 filters) ----
// The whole kernel matrix W [K, filters] (41 KB) stays in shared memory for the lifetime of a persistent CTA; a tile is
// 128 output rows, whose im2col rows are ONE contiguous span of x ((128 - 1) * stride * c_in + K floats), staged with
// 16-byte loads.  256 threads = 16 row groups x 16 column groups; a thread owns 8 rows x NT columns (columns
// tx + 16 c: conflict-free weight reads, broadcast activation reads), i.e. 8 NT FFMAs per 8 + NT shared loads -- FMA-bound,
// and no column tile is ever 1/64 full (the 64 x 64 kernel spends a third of its CTAs on filter 128 alone).
// Exact fp32, same summation order per output as the generic kernel (k ascending).
template <int NT>
__global__ void __launch_bounds__(256)
conv1d_rows_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                   int batch, int rows, int c_in, int taps, int filters, int stride, int left, int rows_out,
                   int act, float *__restrict__ out) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int RT = 128;                                   // rows per tile
  const int K = taps * c_in, NP = 16 * NT, hop = stride * c_in;
  const int span = (RT - 1) * hop + K;
  float *Ws = smem_f;                                       // [K][NP], zero beyond `filters`
  float *xs = smem_f + K * NP;                              // [span] (+ pad)
  for (int e = threadIdx.x; e < K * NP; e += 256) {
    const int k = e / NP, n = e - k * NP;
    Ws[e] = n < filters ? __ldg(w + static_cast<int64_t>(k) * filters + n) : 0.f;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float bv[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) bv[c] = (bias && tx + 16 * c < filters) ? __ldg(bias + tx + 16 * c) : 0.f;
  const int tiles_per = (rows_out + RT - 1) / RT, tiles = batch * tiles_per;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int b = t / tiles_per, r0 = (t - b * tiles_per) * RT;
    const float *xb = x + static_cast<int64_t>(b) * rows * c_in;
    const int64_t g0 = static_cast<int64_t>(r0 * stride - left) * c_in;      // first x element of the span
    const int64_t n_x = static_cast<int64_t>(rows) * c_in;
    __syncthreads();                                        // the previous tile's reads are done (and Ws is complete)
    if (g0 >= 0 && g0 + span <= n_x && (reinterpret_cast<uintptr_t>(xb + g0) & 15) == 0) {
      for (int e = threadIdx.x; e < span / 4; e += 256)
        reinterpret_cast<float4 *>(xs)[e] = __ldg(reinterpret_cast<const float4 *>(xb + g0) + e);
      for (int e = (span / 4) * 4 + threadIdx.x; e < span; e += 256) xs[e] = __ldg(xb + g0 + e);
    } else {
      for (int e = threadIdx.x; e < span; e += 256) {
        const int64_t g = g0 + e;
        xs[e] = (g >= 0 && g < n_x) ? __ldg(xb + g) : 0.f;  // zero padding on both sides ('same') / beyond the last row
      }
    }
    __syncthreads();
    float acc[8][NT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < NT; ++c) acc[i][c] = 0.f;
    const float *xr = xs + (8 * ty) * hop;
    const float *wr = Ws + tx;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float a[8], bb[NT];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = xr[i * hop + k];
#pragma unroll
      for (int c = 0; c < NT; ++c) bb[c] = wr[k * NP + 16 * c];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < NT; ++c) acc[i][c] = fmaf(a[i], bb[c], acc[i][c]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ro = r0 + 8 * ty + i;
      if (ro >= rows_out) continue;
      float *orow = out + (static_cast<int64_t>(b) * rows_out + ro) * filters;
#pragma unroll
      for (int c = 0; c < NT; ++c) {
        const int n = tx + 16 * c;
        if (n < filters) orow[n] = activate(acc[i][c] + bv[c], act);
      }
    }
  }
}

// The reference layer itself: 129 filters = one block of 128 columns + a ragged tail of REM <= 8 columns.  A thread owns
// 8 rows x 8 columns tx + 16 c (so that the 16 threads of a row group store 16 consecutive floats: coalesced), but the
// weights sit PERMUTED in shared memory -- column tx + 16 c at 4 tx + c (c < 4) / 64 + 4 tx + c - 4 -- so the thread's
// eight weights of a k are two conflict-free LDS.128; activations are LDS.128 over four k at a time: 16 shared loads per
// 256 FFMAs.  The tail columns are spread over the threads (one row each for every second thread) instead of wasting a
// ninth strided column slot on every thread.  Needs stride * c_in and K to be multiples of 4.
template <int REM>
__global__ void __launch_bounds__(256)
conv1d_rows128_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                      int batch, int rows, int c_in, int taps, int filters, int stride, int left, int rows_out,
                      int act, float *__restrict__ out) {
  extern __shared__ __align__(16) float smem_f[];
  constexpr int RT = 128, NP = 128 + 8;
  const int K = taps * c_in, hop = stride * c_in;
  const int span = (RT - 1) * hop + K;
  float *Ws = smem_f;                                       // [K][136]: columns 0..127 permuted, then the tail (zero padded)
  float *xs = smem_f + K * NP;
  for (int e = threadIdx.x; e < K * NP; e += 256) {
    const int k = e / NP, n = e - k * NP;
    int pos = n;
    if (n < 128) {
      const int t = n & 15, c = n >> 4;                     // column n = t + 16 c
      pos = c < 4 ? 4 * t + c : 64 + 4 * t + (c - 4);
    }
    Ws[k * NP + pos] = n < filters ? __ldg(w + static_cast<int64_t>(k) * filters + n) : 0.f;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float bv[8], bt[REM];
#pragma unroll
  for (int c = 0; c < 8; ++c) bv[c] = bias ? __ldg(bias + tx + 16 * c) : 0.f;
#pragma unroll
  for (int c = 0; c < REM; ++c) bt[c] = (bias && 128 + c < filters) ? __ldg(bias + 128 + c) : 0.f;
  const int tiles_per = (rows_out + RT - 1) / RT, tiles = batch * tiles_per;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int b = t / tiles_per, r0 = (t - b * tiles_per) * RT;
    const float *xb = x + static_cast<int64_t>(b) * rows * c_in;
    const int64_t g0 = static_cast<int64_t>(r0 * stride - left) * c_in;
    const int64_t n_x = static_cast<int64_t>(rows) * c_in;
    __syncthreads();
    if (g0 >= 0 && g0 + span <= n_x && (reinterpret_cast<uintptr_t>(xb + g0) & 15) == 0) {
      for (int e = threadIdx.x; e < span / 4; e += 256)
        reinterpret_cast<float4 *>(xs)[e] = __ldg(reinterpret_cast<const float4 *>(xb + g0) + e);
    } else {
      for (int e = threadIdx.x; e < span; e += 256) {
        const int64_t g = g0 + e;
        xs[e] = (g >= 0 && g < n_x) ? __ldg(xb + g) : 0.f;
      }
    }
    __syncthreads();
    float acc[8][8], tail[REM];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int c = 0; c < REM; ++c) tail[c] = 0.f;
    const float *xr = xs + (8 * ty) * hop;
    const float *wr = Ws + 4 * tx;
    const float *xt = xs + (8 * ty + (tx >> 1)) * hop;      // the tail row of this thread (threads with even tx)
    for (int k = 0; k < K; k += 4) {
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(xr + i * hop + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b0 = *reinterpret_cast<const float4 *>(wr + (k + kk) * NP);
        const float4 b1 = *reinterpret_cast<const float4 *>(wr + (k + kk) * NP + 64);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(av, bb[c], acc[i][c]);
        }
      }
      if ((tx & 1) == 0) {
        const float4 av = *reinterpret_cast<const float4 *>(xt + k);
        const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int c = 0; c < REM; ++c) tail[c] = fmaf(a[kk], Ws[(k + kk) * NP + 128 + c], tail[c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ro = r0 + 8 * ty + i;
      if (ro >= rows_out) continue;
      float *orow = out + (static_cast<int64_t>(b) * rows_out + ro) * filters + tx;
#pragma unroll
      for (int c = 0; c < 8; ++c) orow[16 * c] = activate(acc[i][c] + bv[c], act);
    }
    if ((tx & 1) == 0) {
      const int ro = r0 + 8 * ty + (tx >> 1);
      if (ro < rows_out) {
        float *orow = out + (static_cast<int64_t>(b) * rows_out + ro) * filters + 128;
#pragma unroll
        for (int c = 0; c < REM; ++c)
          if (128 + c < filters) orow[c] = activate(tail[c] + bt[c], act);
      }
    }
  }
}

static int launch_conv1d_rows128(const float *d_x, const float *d_w, const float *d_b, int batch, int rows, int c_in,
                                 int taps, int filters, int stride, int left, int rows_out, int act, float *d_out,
                                 cudaStream_t stream) {
  const int K = taps * c_in, hop = stride * c_in;
  const size_t smem = sizeof(float) * (static_cast<size_t>(K) * 136 + static_cast<size_t>(127) * hop + K + 8);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = static_cast<int64_t>(batch) * ((rows_out + 127) / 128);
  const int rem = filters - 128;
  // persistent grid = exactly the CTAs that are resident at once (a larger grid would run a second, half-empty wave)
  int per_sm = 1;
  if (rem <= 1) {
    SEP_CUDA(cudaFuncSetAttribute(conv1d_rows128_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SEP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv1d_rows128_kernel<1>, 256, smem));
  } else {
    SEP_CUDA(cudaFuncSetAttribute(conv1d_rows128_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SEP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv1d_rows128_kernel<8>, 256, smem));
  }
  const int grid = static_cast<int>(std::min<int64_t>(tiles, static_cast<int64_t>(sms) * std::max(per_sm, 1)));
  profile_begin(stream, "conv1d_rows128_kernel<REM=%d> (fp32 SIMT, weights resident in shared memory, 128-row tiles, 8 x 8 "
                "micro-tiles; taps=%d c_in=%d filters=%d stride=%d)", rem <= 1 ? 1 : 8, taps, c_in, filters, stride);
  if (rem <= 1) {
    conv1d_rows128_kernel<1><<<grid, 256, smem, stream>>>(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left,
                                                          rows_out, act, d_out);
  } else {
    conv1d_rows128_kernel<8><<<grid, 256, smem, stream>>>(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left,
                                                          rows_out, act, d_out);
  }
  profile_end(stream);
  SEP_LAUNCHED();
  return SEP_OK;
}

template <int NT>
static int launch_conv1d_rows(const float *d_x, const float *d_w, const float *d_b, int batch, int rows, int c_in, int taps,
                              int filters, int stride, int left, int rows_out, int act, float *d_out, size_t smem,
                              cudaStream_t stream) {
  SEP_CUDA(cudaFuncSetAttribute(conv1d_rows_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = static_cast<int64_t>(batch) * ((rows_out + 127) / 128);
  int per_sm = 1;
  SEP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv1d_rows_kernel<NT>, 256, smem));
  const int grid = static_cast<int>(std::min<int64_t>(tiles, static_cast<int64_t>(sms) * std::max(per_sm, 1)));
  profile_begin(stream, "conv1d_rows_kernel<NT=%d> (fp32 SIMT, weights resident in shared memory, 128-row tiles; taps=%d c_in=%d "
                "filters=%d stride=%d)", NT, taps, c_in, filters, stride);
  conv1d_rows_kernel<NT><<<grid, 256, smem, stream>>>(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left, rows_out,
                                                     act, d_out);
  profile_end(stream);
  SEP_LAUNCHED();
  return SEP_OK;
}

int conv1d_tc_try(const float *d_x, const float *d_w, const float *d_b, int batch, int rows, int c_in, int taps,
                  int filters, int stride, int left, int rows_out, int act, float *d_out, cudaStream_t stream,
                  bool *handled);

}  // namespace sep

using namespace sep;

extern "C" int sep_conv1d_f32(const float *x, const float *kernel, const float *bias, int batch,
                              int rows, int c_in, int taps, int filters, int stride, int padding,
                              int activation, float *out, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(x && kernel && out, "sep_conv1d_f32: null argument");
  SEP_REQUIRE(batch >= 1 && rows >= 1 && c_in >= 1 && taps >= 1 && filters >= 1 && stride >= 1,
              "sep_conv1d_f32: bad shape");
  SEP_REQUIRE(padding == SEP_PAD_VALID || padding == SEP_PAD_SAME, "sep_conv1d_f32: bad padding code %d", padding);
  SEP_REQUIRE(activation >= SEP_ACT_LINEAR && activation <= SEP_ACT_RELU, "sep_conv1d_f32: bad activation code %d", activation);
  int rc = check_mem(mem);
  if (rc) return rc;
  int rows_out, left = 0;
  if (padding == SEP_PAD_SAME) {   // TensorFlow 'same': left = total / 2, surplus on the right
    rows_out = (rows + stride - 1) / stride;
    const int total = std::max((rows_out - 1) * stride + taps - rows, 0);
    left = total / 2;
  } else {
    SEP_REQUIRE(rows >= taps, "sep_conv1d_f32: 'valid' padding needs rows >= taps");
    rows_out = (rows - taps) / stride + 1;
  }
  Scratch s(stream);
  const float *d_x, *d_w, *d_b;
  float *d_out;
  if ((rc = stage_in(s, x, static_cast<size_t>(batch) * rows * c_in, mem, &d_x))) return rc;
  if ((rc = stage_in(s, kernel, static_cast<size_t>(taps) * c_in * filters, mem, &d_w))) return rc;
  if ((rc = stage_in(s, bias, static_cast<size_t>(filters), mem, &d_b))) return rc;
  const size_t n_out = static_cast<size_t>(batch) * rows_out * filters;
  if ((rc = stage_out(s, out, n_out, mem, &d_out))) return rc;
  // the reference layer over a dataset shard: 3xTF32 on the tensor cores (conv1d_tc.cu); SEPCORE_CONV_SIMT=1 keeps the
  // exact-fp32 SIMT kernels below (the cross-check)
  const char *simt_env = getenv("SEPCORE_CONV_SIMT");
  const bool simt_only = simt_env && atoi(simt_env) != 0;
  if (!simt_only) {
    bool handled = false;
    if ((rc = conv1d_tc_try(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left, rows_out, activation, d_out,
                            stream, &handled)))
      return rc;
    if (handled) {
      if ((rc = copy_back(s, out, d_out, n_out, mem))) return rc;
      return finish(s, mem);
    }
  }
  {
    // small contraction, many rows: the weights-resident kernel (the reference's Conv1D(129, 2) on [B, K, 40])
    const int K = taps * c_in, NT = (filters + 15) / 16, hop = stride * c_in;
    const size_t smem = sizeof(float) * (static_cast<size_t>(K) * 16 * NT + static_cast<size_t>(127) * hop + K + 8);
    const bool big = static_cast<int64_t>(batch) * rows_out >= 4096;
    if (big && filters >= 128 && filters <= 136 && K % 4 == 0 && hop % 4 == 0 &&
        sizeof(float) * (static_cast<size_t>(K) * 136 + static_cast<size_t>(127) * hop + K + 8) <= 100 * 1024) {
      if ((rc = launch_conv1d_rows128(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left, rows_out, activation,
                                      d_out, stream)))
        return rc;
      if ((rc = copy_back(s, out, d_out, n_out, mem))) return rc;
      return finish(s, mem);
    }
    if (NT >= 1 && NT <= 9 && smem <= 100 * 1024 && big) {
      switch (NT) {
#define SEP_ROWS_CASE(N)                                                                                              \
  case N:                                                                                                             \
    rc = launch_conv1d_rows<N>(d_x, d_w, d_b, batch, rows, c_in, taps, filters, stride, left, rows_out, activation,  \
                               d_out, smem, stream);                                                                  \
    break;
        SEP_ROWS_CASE(1) SEP_ROWS_CASE(2) SEP_ROWS_CASE(3) SEP_ROWS_CASE(4) SEP_ROWS_CASE(5)
        SEP_ROWS_CASE(6) SEP_ROWS_CASE(7) SEP_ROWS_CASE(8) SEP_ROWS_CASE(9)
#undef SEP_ROWS_CASE
      }
      if (rc) return rc;
      if ((rc = copy_back(s, out, d_out, n_out, mem))) return rc;
      return finish(s, mem);
    }
  }
  dim3 grid((rows_out + kTM - 1) / kTM, (filters + kTN - 1) / kTN, batch);
  profile_begin(stream, "conv1d_kernel (fp32 SIMT, 64x64 tiles; taps=%d c_in=%d filters=%d stride=%d)", taps, c_in, filters, stride);
  conv1d_kernel<<<grid, 256, 0, stream>>>(d_x, d_w, d_b, rows, c_in, taps, filters, stride, left,
                                          rows_out, activation, d_out);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, out, d_out, n_out, mem))) return rc;
  return finish(s, mem);
}
