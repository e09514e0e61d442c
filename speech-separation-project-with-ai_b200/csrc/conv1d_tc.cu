// conv1d_tc.cu -- the reference's learned front end, Conv1D(129, 2, activation='sigmoid', padding='same') on
// [B, K, 40] segments (Raw_with_Convlayer.ipynb:389, cell 13; segmentation :83-100), on the 5th-generation tensor cores.
//
// out[b, r, n] = act(bias[n] + sum_k A[r, k] W[k, n]) with A[r, :] = the CONTIGUOUS span xflat[(r stride - left) c_in ..
// + taps c_in) (rows of x are contiguous, so the im2col row is an overlapping strided view: no im2col buffer).  With
// exact fp32 the layer is FMA-bound on SIMT (16.9 GFLOP per 1024 utterances against 554 MB); as a
// 3xTF32 contraction on the tensor pipe -- x = hi + lo with hi exactly representable in tf32, hi*hi + lo*hi + hi*lo
// accumulated in fp32 in tensor memory: fp32-level accuracy, 1e-6 -- it is HBM-bound.
//
// One persistent CTA per SM, warp-specialised, hand-offs by mbarrier only:
//   warps 8-11 staging: thread s owns output row s of the 128-row tile: its K-float span -> registers (prefetched a
//              tile ahead, 20 x 16-byte loads in flight per thread) -> split hi / lo -> shared memory in the canonical
//              K-major UMMA layout (8-row x 16-byte core matrices)
//   warp 12    one thread issues 3 x (K + 8)/8 tcgen05.mma (kind::tf32, M = 128, N = filters rounded up to 16) into one
//              of two accumulator buffers in tensor memory; tcgen05.commit frees the A tile and publishes the accumulator
//   warps 0-7  epilogue (the critical path: 129 activations per row): TMEM lane = row, two warps per lane quarter
//              taking alternate 16-column chunks: tcgen05.ld -> activation -> a 32 x 17 shared transpose per warp ->
//              stores of 16 consecutive floats per row (rows of 129 floats are 4-byte aligned only: a thread-per-row
//              store would touch 32 sectors per instruction)
// The BIAS rides in the contraction: K is extended by one 8-wide step whose A column is the constant 1 and whose W row is
// the bias, so the tensor pipe adds it and the epilogue does not load it.  The weights ((K + 8) x N, split hi / lo,
// 101 KB for the reference layer) stay in shared memory for the lifetime of the CTA.  Waiting roles back off with
// nanosleep: a spinning mbarrier loop on 5 warps was taking issue slots from the epilogue (ncu, first version).
#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace sep {

constexpr int kCtM = 128;                         // rows per tile (MMA M)
constexpr int kCtThreads = 256 + 128 + 32;        // epilogue (2 warpgroups), staging, MMA warp

__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  while (!mbar_try(bar, parity)) __nanosleep(200);
}

struct ConvTcArgs {
  const float *x, *w, *bias;
  float *out;
  int batch, rows, c_in, taps, filters, stride, left, rows_out, act, npad;
};

__device__ __forceinline__ float activate_fast(float v, int act) {
  // sigmoid through MUFU.EX2 / MUFU.RCP: absolute error < 3e-7 on (0, 1), far inside the 1e-5 parity tolerance
  if (act == SEP_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-v));
  if (act == SEP_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

enum : int { kCtAFull = 0, kCtAEmpty = 1, kCtDFull = 2, kCtDEmpty = 4 };

// KQ = K / 4 (16-byte chunks of an im2col row); K = taps * c_in must be a multiple of 8
template <int KQ>
__global__ void __launch_bounds__(kCtThreads, 1) conv1d_tc_kernel(const ConvTcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int K = 4 * KQ, KQA = KQ + 2, KA = 4 * KQA;            // contraction extended by the bias step
  constexpr int kABytes = KQA * (kCtM / 8) * 128;                  // one A operand (hi or lo)
  const int NP = a.npad, bbytes = KQA * (NP / 8) * 128;            // one B operand
  unsigned char *Ahi = sm, *Alo = sm + kABytes, *Bhi = sm + 2 * kABytes, *Blo = Bhi + bbytes;
  float *tr = reinterpret_cast<float *>(Blo + bbytes);             // [8 warps][32][17] epilogue transpose
  uint64_t *bars = reinterpret_cast<uint64_t *>(tr + 8 * 32 * 17);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
  const uint32_t bar0 = smem_u32(bars), sm0 = smem_u32(sm);
#define CT_BAR(i) (bar0 + 8u * static_cast<uint32_t>(i))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per = (a.rows_out + kCtM - 1) / kCtM, n_tiles = a.batch * tiles_per;
  const uint32_t lbo_a = (kCtM / 8) * 128, lbo_b = (NP / 8) * 128;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) {
    mbar_init(CT_BAR(kCtAFull), 128);
    mbar_init(CT_BAR(kCtAEmpty), 1);
    mbar_init(CT_BAR(kCtDFull), 1);
    mbar_init(CT_BAR(kCtDFull + 1), 1);
    mbar_init(CT_BAR(kCtDEmpty), 256);
    mbar_init(CT_BAR(kCtDEmpty + 1), 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights W [K][filters] row-major -> B[n][k] (N rows, K-major), split hi / lo; row K = the bias, rows K+1.. and the
  // columns beyond `filters` are zero
  for (int e = threadIdx.x; e < KA * NP; e += kCtThreads) {
    const int k = e / NP, n = e - k * NP;
    float hi = 0.f, lo = 0.f;
    if (n < a.filters) {
      if (k < K) split_tf32(__ldg(a.w + static_cast<int64_t>(k) * a.filters + n), hi, lo);
      else if (k == K && a.bias) split_tf32(__ldg(a.bias + n), hi, lo);
    }
    const int o = kmajor_off(n, k, NP / 8);
    *reinterpret_cast<float *>(Bhi + o) = hi;
    *reinterpret_cast<float *>(Blo + o) = lo;
  }
  // the constant part of A: column K is 1 (exact in tf32), K+1.. are 0; the staging warps never touch these chunks
  for (int e = threadIdx.x; e < kCtM * 8; e += kCtThreads) {
    const int row = e >> 3, k = K + (e & 7);
    const int o = kmajor_off(row, k, kCtM / 8);
    *reinterpret_cast<float *>(Ahi + o) = k == K ? 1.f : 0.f;
    *reinterpret_cast<float *>(Alo + o) = 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 12) {
    // =========================== MMA warp: one thread issues ===========================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kCtM, NP);
      uint32_t i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        const uint32_t dbuf = i & 1, use = i >> 1;
        mbar_wait_backoff(CT_BAR(kCtAFull), i & 1);                  // the tile's rows are staged
        if (use > 0) mbar_wait_backoff(CT_BAR(kCtDEmpty + dbuf), (use - 1) & 1);   // the epilogue has drained this accumulator
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {                       // hi*hi, lo*hi, hi*lo
          const uint32_t aoff = sm0 + (pass == 1 ? kABytes : 0);
          const uint32_t boff = sm0 + 2 * kABytes + (pass == 2 ? bbytes : 0);
#pragma unroll
          for (int ks = 0; ks < KA / 8; ++ks) {
            umma_tf32(tmem + dbuf * NP, umma_desc(aoff + ks * 2 * lbo_a, lbo_a, 128),
                      umma_desc(boff + ks * 2 * lbo_b, lbo_b, 128), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(CT_BAR(kCtAEmpty));                              // the A tile may be overwritten
        umma_commit(CT_BAR(kCtDFull + dbuf));                        // the accumulator is complete
      }
    }
  } else if (warp >= 8) {
    // =========================== staging warps: thread s = row s of the tile ===========================
    const int s = threadIdx.x - 256;
    const int64_t n_x = static_cast<int64_t>(a.rows) * a.c_in;
    const int hop = a.stride * a.c_in;
    float4 pre[KQ];
    auto fetch = [&](int t) {
      const int b = t / tiles_per, ro = (t - b * tiles_per) * kCtM + s;
      const float *xb = a.x + static_cast<int64_t>(b) * n_x;
      const int64_t g0 = static_cast<int64_t>(ro) * hop - static_cast<int64_t>(a.left) * a.c_in;
      if (ro < a.rows_out && g0 >= 0 && g0 + K <= n_x && ((reinterpret_cast<uintptr_t>(xb + g0) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < KQ; ++q) pre[q] = __ldg(reinterpret_cast<const float4 *>(xb + g0) + q);
      } else {
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int64_t g = g0 + 4 * q + e;
            v[e] = (ro < a.rows_out && g >= 0 && g < n_x) ? __ldg(xb + g) : 0.f;   // 'same' padding / beyond the last row
          }
          pre[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    if (blockIdx.x < n_tiles) fetch(blockIdx.x);
    uint32_t i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      if (i > 0) mbar_wait_backoff(CT_BAR(kCtAEmpty), (i - 1) & 1);  // the MMAs of the previous tile have read A
      const int o = (s >> 3) * 128 + (s & 7) * 16;
#pragma unroll
      for (int q = 0; q < KQ; ++q) {
        float4 hi, lo;
        split_tf32(pre[q].x, hi.x, lo.x); split_tf32(pre[q].y, hi.y, lo.y);
        split_tf32(pre[q].z, hi.z, lo.z); split_tf32(pre[q].w, hi.w, lo.w);
        *reinterpret_cast<float4 *>(Ahi + q * (kCtM / 8) * 128 + o) = hi;
        *reinterpret_cast<float4 *>(Alo + q * (kCtM / 8) * 128 + o) = lo;
      }
      fence_async_smem();                                            // generic-proxy stores -> visible to the tensor pipe
      mbar_arrive(CT_BAR(kCtAFull));
      if (t + static_cast<int>(gridDim.x) < n_tiles) fetch(t + gridDim.x);   // the next tile's rows, in flight during the MMAs
    }
  } else {
    // =========================== epilogue warps: TMEM lane = row ===========================
    // warps w and w + 4 share TMEM lanes 32 (w & 3) .. and take alternate 16-column chunks
    float *trw = tr + warp * 32 * 17;
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const int col = lane & 15, rsel = lane >> 4;
    uint32_t i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      const uint32_t dbuf = i & 1, use = i >> 1;
      const int b = t / tiles_per, r0 = (t - b * tiles_per) * kCtM + quarter * 32;
      float *obase = a.out + (static_cast<int64_t>(b) * a.rows_out + r0) * a.filters;
      const bool full_rows = r0 + 32 <= a.rows_out;
      mbar_wait(CT_BAR(kCtDFull + dbuf), use & 1);
      tc_fence_after();
      for (int c0 = 16 * half; c0 < NP; c0 += 32) {
        float d[16];
        __syncwarp();                                                // tcgen05.ld is warp-collective; trw reads are done
        tmem_ld16(lane_addr + dbuf * NP + c0, d);
#pragma unroll
        for (int e = 0; e < 16; ++e) trw[lane * 17 + e] = activate_fast(d[e], a.act);
        __syncwarp();
        // 16 consecutive floats of a row per half-warp: two rows per instruction
        float *op = obase + static_cast<int64_t>(rsel) * a.filters + c0 + col;
        const float *tp = trw + rsel * 17 + col;
        if (full_rows && c0 + 16 <= a.filters) {
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) op[static_cast<int64_t>(2 * rr) * a.filters] = tp[2 * rr * 17];
        } else {
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            if (r0 + 2 * rr + rsel < a.rows_out && c0 + col < a.filters) op[static_cast<int64_t>(2 * rr) * a.filters] = tp[2 * rr * 17];
        }
      }
      tc_fence_before();
      mbar_arrive(CT_BAR(kCtDEmpty + dbuf));                         // this thread's reads of the accumulator are done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
  }
#undef CT_BAR
}

// Takes the call when the shape fits; *handled tells.  Needs K = taps * c_in == 80 (the reference layer; other
// contractions stay on the SIMT kernels), filters <= 256, and enough rows to fill the machine.
int conv1d_tc_try(const float *d_x, const float *d_w, const float *d_b, int batch, int rows, int c_in, int taps,
                  int filters, int stride, int left, int rows_out, int act, float *d_out, cudaStream_t stream,
                  bool *handled) {
  *handled = false;
  constexpr int KQ = 20;
  const int K = taps * c_in, NP = (filters + 15) / 16 * 16;
  if (K != 4 * KQ || NP > 256 || static_cast<int64_t>(batch) * rows_out < 4096) return SEP_OK;
  const size_t smem = 2 * static_cast<size_t>(KQ + 2) * (kCtM / 8) * 128 + 2 * static_cast<size_t>(KQ + 2) * (NP / 8) * 128 +
                      8 * 32 * 17 * sizeof(float) + 128;
  if (smem > 225 * 1024) return SEP_OK;
  *handled = true;
  ConvTcArgs a{d_x, d_w, d_b, d_out, batch, rows, c_in, taps, filters, stride, left, rows_out, act, NP};
  SEP_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = static_cast<int64_t>(batch) * ((rows_out + kCtM - 1) / kCtM);
  const int grid = static_cast<int>(std::min<int64_t>(tiles, sms));
  profile_begin(stream, "conv1d_tc_kernel<K=%d> (tcgen05 kind::tf32 x3, M=128 N=%d, TMEM accumulators x2; taps=%d c_in=%d "
                "filters=%d stride=%d)", K, NP, taps, c_in, filters, stride);
  conv1d_tc_kernel<KQ><<<grid, kCtThreads, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  return SEP_OK;
}

}  // namespace sep
