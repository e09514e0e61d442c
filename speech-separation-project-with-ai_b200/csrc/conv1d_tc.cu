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
#include <cstdlib>

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
  // the warp index through a shuffle: ptxas then KNOWS that it is warp-uniform, so the role branches are uniform and the
  // addresses / descriptors of the MMA warps can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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


// ------------------------------------------------------------------------------------------------------------------
// conv1d_tc2_kernel -- the same contraction with the OUTPUT TILE AS ONE BULK STORE (odd `filters`, e.g. the
// reference's 129).  Rows are numbered across the whole [batch * rows_out] output, so a 128-row tile is one
// contiguous run of 128 * filters floats in HBM:
//   * the A operand (the tile's im2col rows, hi | lo) lives in TENSOR MEMORY (lane = row; the staging thread that owns
//     a row writes it with tcgen05.st), which frees the shared memory for two output tiles;
//   * the epilogue thread of a row applies the activation and stores its values into a shared-memory image of the
//     output tile (pitch = filters floats, odd, so the 32 rows of a warp hit 32 different banks); when the three warps
//     of a 32-row quarter are done, the store warp hands the quarter (32 * filters * 4 bytes, contiguous in HBM) to the
//     TMA engine (cp.async.bulk shared -> global).  No transposes, no per-element global stores;
//   * the bias is folded into the activation's first FFMA (sigmoid: ex2(d * -log2e - bias * log2e));
//   * hand-offs: the A operand is handed over per K half (own full / empty barriers: the staging of the next tile's
//     first half overlaps the MMAs on the second half), the accumulator is double buffered, a store warp issues the
//     bulk stores so that no epilogue warp waits for another one, and every wait is a hardware-suspending try_wait.
// Roles: warps 0-11 epilogue (three per 32-row quarter, a third of the columns each), 12-15 staging, 16 MMA issue
// (converged warp, one elected lane), 17 bulk stores.
// TMEM columns: D0 [0, NP), D1 [NP, 2 NP), A_hi [2 NP, 2 NP + K), A_lo [2 NP + K, 2 NP + 2 K)  (448 of 512 here).
// What bounds it (in-kernel clock64 timelines, one-role-removed builds and microbenchmarks, round 2): the epilogue.  A
// scheduler finishes one 16-column chunk per ~500 cycles whether two or three epilogue warps share it -- twice the 264
// cycles the same code takes in isolation, where it sits on the MUFU floor (8 cycles per warp instruction, 32 per chunk)
// -- and MMA issue (56 cycles each from the converged warp, 30 per tile), staging (~1600 cycles per K half, L2 latency
// exposed: registers hold one half) and the epilogue wait for each other through one tile of slack.  With the same roles
// running UNSYNCHRONISED (tools/ubench/epilogue_ldtm.cu) a scheduler finishes a chunk per 280-297 cycles next to
// full-rate MMAs, tile stores and staging: no shared unit is the limit, the hand-offs are.  Tried and measured slower or
// equal: two N / 2 accumulators with two issuing warps (60 MMAs per tile, each paying the issue cost), 32-column epilogue
// blocks, all columns read in one tcgen05.ld burst with the accumulator released at once, half of the reciprocals on the
// FMA pipe, a software-pipelined epilogue (EX2 of chunk c next to RCP of chunk c - 1), staggered group starts, two
// staging sets (one per K half, 80 registers).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// the registers pass THROUGH the wait, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ float ex2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_pinned(float x) {
  float y;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// v[e] = act(d[e] + bias[e]) for N values; bb holds the bias (sigmoid: -log2e * bias)
template <int ACT, int N>
__device__ __forceinline__ void activate16(const uint32_t (&r)[16], const float (&bb)[16], float (&v)[16]) {
  constexpr float kLog2e = 1.4426950408889634f;
  if (ACT == SEP_ACT_SIGMOID) {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = fmaf(__uint_as_float(r[e]), -kLog2e, bb[e]);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = ex2_pinned(v[e]);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = 1.f + v[e];
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = rcp_pinned(v[e]);
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const float d = __uint_as_float(r[e]) + bb[e];
      v[e] = ACT == SEP_ACT_RELU ? fmaxf(d, 0.f) : d;
    }
  }
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4_wait(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
template <int ACT>
__device__ __forceinline__ void activate4(const uint32_t (&r)[4], const float (&bb)[4], float (&v)[4]) {
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float d = __uint_as_float(r[e]);
    if (ACT == SEP_ACT_SIGMOID) v[e] = rcp_pinned(1.f + ex2_pinned(fmaf(d, -kLog2e, bb[e])));
    else if (ACT == SEP_ACT_RELU) v[e] = fmaxf(d + bb[e], 0.f);
    else v[e] = d + bb[e];
  }
}
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
enum : int { kC2AFull = 0, kC2AEmpty = 2, kC2DFull = 4, kC2DEmpty = 6, kC2OEmpty = 8, kC2OFull = 10, kC2Bars = 18,
              // XS layout: six output quarter slots and the x tile
              kC3OEmpty = 8, kC3OFull = 14, kC3XFull = 20, kC3Bars = 21, kC3Slots = 6 };
constexpr int kC2Groups = 3;                        // epilogue warps per 32-row quarter (column groups)
constexpr int kCt2Threads = 32 * (4 * kC2Groups + 4 + 2);   // epilogue, staging, MMA warp, store warp

// XS = true (the reference geometry: two taps, stride 1, no left padding): the x rows of a tile are ONE contiguous span (row r
// is floats [40 r, 40 r + 80): its second K half is the first K half of row r + 1), brought into shared memory by one
// cp.async.bulk instead of 20 LDG.128 per thread whose lanes are 160 bytes apart (32 sectors per request: as many LSU
// wavefronts as the whole epilogue), and the output image shrinks from two tiles to a ring of six 32-row quarters to make
// room for it.
template <int KQ, int ACT, bool XS>
__global__ void __launch_bounds__(kCt2Threads, 1) conv1d_tc2_kernel(const ConvTcArgs a) {
  static_assert(KQ % 4 == 0, "the staging threads write 8 columns of tensor memory at a time, per K half");
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int K = 4 * KQ;
  const int NP = a.npad, bbytes = KQ * (NP / 8) * 128;             // one B operand
  const int tile_bytes = kCtM * a.filters * 4, qbytes = 32 * a.filters * 4;
  unsigned char *Bhi = sm, *Blo = sm + bbytes, *ob = Blo + bbytes; // output image behind the weights: two tiles / six quarters
  const int xbytes = XS ? ((kCtM + 1) * a.c_in * 4 + 127) / 128 * 128 : 0;
  unsigned char *xbuf = ob + (XS ? kC3Slots * qbytes : 2 * tile_bytes);   // XS: rows R0 .. R0 + 128 of x, contiguous
  float *nb = reinterpret_cast<float *>(xbuf + xbytes);            // per column: the bias term of the activation
  uint64_t *bars = reinterpret_cast<uint64_t *>(nb + NP + 32);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + (XS ? kC3Bars : kC2Bars));
  const uint32_t bar0 = smem_u32(bars), sm0 = smem_u32(sm);
#define CT_BAR(i) (bar0 + 8u * static_cast<uint32_t>(i))
  // the warp index through a shuffle: ptxas then KNOWS that it is warp-uniform, so the role branches are uniform and the
  // addresses / descriptors of the MMA warps can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t total_rows = static_cast<int64_t>(a.batch) * a.rows_out;
  const int n_tiles = static_cast<int>((total_rows + kCtM - 1) / kCtM);
  const uint32_t lbo_b = (NP / 8) * 128;
  const int nch = NP / 16;                                         // 16-column chunks, dealt to the epilogue column groups
  constexpr float kLog2e = 1.4426950408889634f;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) {
    for (int h = 0; h < 2; ++h) {
      mbar_init(CT_BAR(kC2AFull + h), 128);
      mbar_init(CT_BAR(kC2AEmpty + h), 1);
      mbar_init(CT_BAR(kC2DFull + h), 1);
      mbar_init(CT_BAR(kC2DEmpty + h), 128 * kC2Groups);
      if (!XS) {
        mbar_init(CT_BAR(kC2OEmpty + h), 1);
        for (int q = 0; q < 4; ++q) mbar_init(CT_BAR(kC2OFull + 2 * q + h), 32 * kC2Groups);
      }
    }
    if (XS) {
      for (int sl = 0; sl < kC3Slots; ++sl) {
        mbar_init(CT_BAR(kC3OEmpty + sl), 1);
        mbar_init(CT_BAR(kC3OFull + sl), 32 * kC2Groups);
      }
      mbar_init(CT_BAR(kC3XFull), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights W [K][filters] row-major -> B[n][k] (N rows, K-major), split hi / lo; columns beyond `filters` are zero
  // (eight loads in flight per thread: a one-load-at-a-time loop took 5 us of every launch)
  for (int e0 = threadIdx.x; e0 < K * NP; e0 += 8 * kCt2Threads) {
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * kCt2Threads, k = e / NP, n = e - k * NP;
      wv[u] = (e < K * NP && n < a.filters) ? __ldg(a.w + static_cast<int64_t>(k) * a.filters + n) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * kCt2Threads, k = e / NP, n = e - k * NP;
      if (e < K * NP) {
        float hi, lo;
        split_tf32(wv[u], hi, lo);
        const int o = kmajor_off(n, k, NP / 8);
        *reinterpret_cast<float *>(Bhi + o) = hi;
        *reinterpret_cast<float *>(Blo + o) = lo;
      }
    }
  }
  for (int n = threadIdx.x; n < NP + 32; n += kCt2Threads) {
    const float bv = (a.bias && n < a.filters) ? __ldg(a.bias + n) : 0.f;
    nb[n] = ACT == SEP_ACT_SIGMOID ? -kLog2e * bv : bv;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_a = tmem + 2 * NP;                             // A: hi K columns | lo K columns

  if (warp == 4 * kC2Groups + 5) {
    // =========================== store warp: one thread hands finished quarters to the TMA engine ===========================
    if (lane == 0 && XS) {
      // one bulk group per 32-row quarter; a slot is free again once the group that read it is two groups back
      uint32_t n = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int q = 0; q < 4; ++q, ++n) {
          const uint32_t slot = n % kC3Slots;
          mbar_wait_suspend(CT_BAR(kC3OFull + slot), (n / kC3Slots) & 1);       // all column groups have written these 32 rows
          const int64_t R0 = static_cast<int64_t>(t) * kCtM + q * 32;
          const int rows_here = static_cast<int>(std::min<int64_t>(32, total_rows - R0));
          const uint32_t bytes = static_cast<uint32_t>(rows_here > 0 ? rows_here : 0) * a.filters * 4u;
          if (bytes > 0 && (bytes & 15u) == 0)                       // (a ragged last quarter is stored by its writers)
            bulk_store(a.out + R0 * a.filters, smem_u32(ob + slot * qbytes), bytes);
          bulk_commit();
          if (n >= 2) {
            bulk_wait_read<2>();                                     // group n - 2 has been read:
            mbar_arrive(CT_BAR(kC3OEmpty + (n - 2) % kC3Slots));     // the epilogue may write its slot again
          }
        }
      }
      bulk_wait_read<0>();                                           // shared memory stays valid until the engine has read it
    } else if (lane == 0) {
      uint32_t i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        if (i > 0) {
          bulk_wait_read<0>();                                       // the engine has read the previous tile's buffer:
          mbar_arrive(CT_BAR(kC2OEmpty + ((i - 1) & 1)));            // the epilogue may write it again
        }
        unsigned char *otile = ob + (i & 1) * tile_bytes;
        for (int q = 0; q < 4; ++q) {
          mbar_wait_suspend(CT_BAR(kC2OFull + 2 * q + (i & 1)), (i >> 1) & 1);   // all column groups have written these 32 rows
          const int64_t R0 = static_cast<int64_t>(t) * kCtM + q * 32;
          const int rows_here = static_cast<int>(std::min<int64_t>(32, total_rows - R0));
          const uint32_t bytes = static_cast<uint32_t>(rows_here > 0 ? rows_here : 0) * a.filters * 4u;
          if (bytes > 0 && (bytes & 15u) == 0)                       // (a ragged last quarter is stored by its writers)
            bulk_store(a.out + R0 * a.filters, smem_u32(otile + q * qbytes), bytes);
        }
        bulk_commit();
      }
      bulk_wait_read<0>();                                           // shared memory stays valid until the engine has read it
    }
  } else if (warp == 4 * kC2Groups + 4) {
    // =========================== MMA warp: runs converged, one elected lane issues ===========================
    // ONE accumulator of N = NP columns: with the A operand in tensor memory an MMA costs max(N / 2, 64) cycles (the
    // 4 KB of A are read at the 64 B per cycle of the tensor-memory read port), so two N / 2 halves cost twice as much
    {
      const uint32_t idesc = umma_idesc_tf32(kCtM, NP);
      uint32_t i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        const uint32_t dbuf = i & 1;
        if (i > 1) mbar_wait_suspend(CT_BAR(kC2DEmpty + dbuf), ((i >> 1) - 1) & 1);   // the epilogue has drained this accumulator
        const uint32_t d = tmem + dbuf * NP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                // the two K halves of A: own full / empty barriers
          mbar_wait_suspend(CT_BAR(kC2AFull + h), i & 1);            // this half of the tile's rows is in tensor memory
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < K / 16; ++kk) {                      // hi*hi, lo*hi, hi*lo per 8-wide k-step
            const int ks = h * (K / 16) + kk;
            const uint64_t bhi = umma_desc(sm0 + ks * 2 * lbo_b, lbo_b, 128);
            const uint64_t blo = umma_desc(sm0 + bbytes + ks * 2 * lbo_b, lbo_b, 128);
            umma_tf32_ts_elect(d, tm_a + 8 * ks, bhi, idesc, ks > 0 ? 1u : 0u);
            umma_tf32_ts_elect(d, tm_a + K + 8 * ks, bhi, idesc, 1u);
            umma_tf32_ts_elect(d, tm_a + 8 * ks, blo, idesc, 1u);
          }
          umma_commit_elect(CT_BAR(kC2AEmpty + h));                  // this half of A may be overwritten
        }
        umma_commit_elect(CT_BAR(kC2DFull + dbuf));                  // the accumulator is complete
      }
    }
  } else if (warp >= 4 * kC2Groups && warp < 4 * kC2Groups + 4) {
    // =========================== staging warps: thread s = row s of the tile = TMEM lane s ===========================
    const int s = threadIdx.x - 128 * kC2Groups;
    const int64_t n_x = static_cast<int64_t>(a.rows) * a.c_in;
    const int hop = a.stride * a.c_in;
    const uint32_t a_addr = tm_a + (static_cast<uint32_t>(s & ~31) << 16);
    float4 pre[KQ / 2];                                              // one K half of this thread's row
    auto row_of = [&](int t, const float *&xb, int64_t &g0) -> bool {
      const int64_t R = static_cast<int64_t>(t) * kCtM + s;
      const int64_t b = R / a.rows_out;
      const int ro = static_cast<int>(R - b * a.rows_out);
      xb = a.x + b * n_x;
      g0 = static_cast<int64_t>(ro) * hop - static_cast<int64_t>(a.left) * a.c_in;
      return R < total_rows;
    };
    auto fetch_half = [&](int t, int h) {                            // (L2 hits: the tile was prefetched a tile ago)
      const float *xb;
      int64_t g0;
      const bool live = row_of(t, xb, g0);
      if (live && g0 >= 0 && g0 + K <= n_x && ((reinterpret_cast<uintptr_t>(xb + g0) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < KQ / 2; ++q) pre[q] = __ldg(reinterpret_cast<const float4 *>(xb + g0) + h * (KQ / 2) + q);
      } else {
#pragma unroll
        for (int q = 0; q < KQ / 2; ++q) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int64_t g = g0 + 4 * (h * (KQ / 2) + q) + e;
            v[e] = (live && g >= 0 && g < n_x) ? __ldg(xb + g) : 0.f;   // 'same' padding / beyond the last row
          }
          pre[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    };
    auto prefetch_l2 = [&](int t) {                                  // the rows of a later tile: on their way into L2
      const float *xb;
      int64_t g0;
      if (t < n_tiles && row_of(t, xb, g0) && g0 >= 0 && g0 + K <= n_x) {
        const char *p = reinterpret_cast<const char *>(xb + g0);
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p + 4 * K - 4));
      }
    };
    if constexpr (XS) {
      // the tile's rows R0 .. R0 + 128 of x are one contiguous run of floats [40 R0, 40 (R0 + 129)): one bulk copy (clipped at
      // the end of x); thread s reads K half h of its row at chunk s + h of the buffer
      const int64_t x_floats = static_cast<int64_t>(a.batch) * n_x;
      const int chunk = a.c_in * 4;                                  // bytes of a K half (one row of x)
      auto issue_x = [&](int t) {                                    // one thread
        const int64_t f0 = static_cast<int64_t>(t) * kCtM * a.c_in;
        const uint32_t bytes = static_cast<uint32_t>(std::min<int64_t>((kCtM + 1) * a.c_in, x_floats - f0) * 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer's readers are done (barrier below)
        mbar_expect_tx(CT_BAR(kC3XFull), bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(xbuf)), "l"(a.x + f0), "r"(bytes), "r"(CT_BAR(kC3XFull)) : "memory");
      };
      if (s == 0 && blockIdx.x < n_tiles) issue_x(blockIdx.x);
      uint32_t i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        const int64_t R = static_cast<int64_t>(t) * kCtM + s;
        const bool live = R < total_rows;
        const bool last_row = live && (R % a.rows_out) == a.rows_out - 1;   // 'same': its second tap is the zero row
        mbar_wait_suspend(CT_BAR(kC3XFull), i & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 *src = reinterpret_cast<const float4 *>(xbuf + (s + h) * chunk);
          const bool zero = !live || (h == 1 && last_row);
#pragma unroll
          for (int q = 0; q < KQ / 2; ++q) pre[q] = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : src[q];
          if (i > 0) {
            mbar_wait_suspend(CT_BAR(kC2AEmpty + h), (i - 1) & 1);   // the previous tile's MMAs have read this half of A
            tc_fence_after();
          }
          __syncwarp();                                              // tcgen05.st is warp-collective
#pragma unroll
          for (int gg = 0; gg < KQ / 4; ++gg) {
            const int g = h * (KQ / 4) + gg;
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float4 v = pre[2 * gg + q];
              split_tf32(v.x, hi[4 * q], lo[4 * q]); split_tf32(v.y, hi[4 * q + 1], lo[4 * q + 1]);
              split_tf32(v.z, hi[4 * q + 2], lo[4 * q + 2]); split_tf32(v.w, hi[4 * q + 3], lo[4 * q + 3]);
            }
            tmem_st8(a_addr + 8 * g, hi);
            tmem_st8(a_addr + K + 8 * g, lo);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(CT_BAR(kC2AFull + h));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");               // every staging thread has read the buffer
        if (s == 0 && t + static_cast<int>(gridDim.x) < n_tiles) issue_x(t + gridDim.x);
      }
    } else {
    prefetch_l2(blockIdx.x + gridDim.x);
      uint32_t i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
  #pragma unroll
        for (int h = 0; h < 2; ++h) {
          fetch_half(t, h);                                            // in flight while the MMAs still read this half
          if (i > 0) {
            mbar_wait_suspend(CT_BAR(kC2AEmpty + h), (i - 1) & 1);     // the previous tile's MMAs have read this half of A
            tc_fence_after();
          }
          __syncwarp();                                                // tcgen05.st is warp-collective
  #pragma unroll
          for (int gg = 0; gg < KQ / 4; ++gg) {
            const int g = h * (KQ / 4) + gg;
            float hi[8], lo[8];
  #pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float4 v = pre[2 * gg + q];
              split_tf32(v.x, hi[4 * q], lo[4 * q]); split_tf32(v.y, hi[4 * q + 1], lo[4 * q + 1]);
              split_tf32(v.z, hi[4 * q + 2], lo[4 * q + 2]); split_tf32(v.w, hi[4 * q + 3], lo[4 * q + 3]);
            }
            tmem_st8(a_addr + 8 * g, hi);
            tmem_st8(a_addr + K + 8 * g, lo);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(CT_BAR(kC2AFull + h));
        }
        prefetch_l2(t + 2 * gridDim.x);
      }
    }
  } else {
    // =========================== epilogue warps: TMEM lane = row ===========================
    // warps w, w + 4, w + 8 (column groups 0, 1, 2) share TMEM lanes 32 (w & 3) ..
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    // group g takes nch / G chunks, the first nch % G groups one more (a group may be empty: it only arrives)
    const int cbeg = 16 * (grp * (nch / kC2Groups) + min(grp, nch % kC2Groups));
    const int cend = cbeg + 16 * (nch / kC2Groups + (grp < nch % kC2Groups ? 1 : 0));
    uint32_t i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      const uint32_t qn = 4 * i + quarter, slot = qn % kC3Slots;      // XS: this quarter's slot of the output ring
      unsigned char *otile = XS ? ob + slot * qbytes - quarter * qbytes : ob + (i & 1) * tile_bytes;   // (+ quarter * qbytes below)
      float *orow = reinterpret_cast<float *>(otile) + (quarter * 32 + lane) * a.filters;
      if constexpr (XS) {
        if (qn >= kC3Slots) mbar_wait_suspend(CT_BAR(kC3OEmpty + slot), (qn / kC3Slots - 1) & 1);   // the store that read this slot is done
      } else {
        if (i > 1) mbar_wait_suspend(CT_BAR(kC2OEmpty + (i & 1)), ((i >> 1) - 1) & 1);   // the store of tile i - 2 has read this buffer
      }
      mbar_wait_suspend(CT_BAR(kC2DFull + (i & 1)), (i >> 1) & 1);
      tc_fence_after();
      uint32_t r0[16], r1[16], rx[4];
      const uint32_t dcol = lane_addr + (i & 1) * NP + cbeg;
      // chunks of 16 columns that exist in full, then `rem` more columns; up to 4 of those (the reference's 129th
      // column) ride with the last full chunk instead of costing a chunk of their own
      const int cols = max(0, min(cend, a.filters) - cbeg), full = cols >> 4, rem = cols & 15;
      const bool ride = rem > 0 && rem <= 4 && full > 0;
      const int n_chunks = full + ((rem > 0 && !ride) ? 1 : 0);
      // 16 columns of this row.  The order of the special-function instructions is PINNED (asm volatile): 16 x EX2, then
      // 16 x RCP.  Left to itself ptxas puts each dependent FADD / RCP two instructions behind its EX2, and with two
      // epilogue warps per scheduler nothing hides the 20-cycle MUFU latency.
      auto finish_chunk = [&](const uint32_t (&r)[16], int ch) {
        const int c0 = cbeg + 16 * ch;
        const float4 *nb4 = reinterpret_cast<const float4 *>(nb + c0);
        float bb[16], v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bq = nb4[q];
          bb[4 * q] = bq.x; bb[4 * q + 1] = bq.y; bb[4 * q + 2] = bq.z; bb[4 * q + 3] = bq.w;
        }
        activate16<ACT, 16>(r, bb, v);
        if (ch < full) {
#pragma unroll
          for (int e = 0; e < 16; ++e) orow[c0 + e] = v[e];
          if (ride && ch == full - 1) {                              // + the columns that ride along
            const float4 bq = nb4[4];
            const float bx[4] = {bq.x, bq.y, bq.z, bq.w};
            float vx[4];
            activate4<ACT>(rx, bx, vx);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < rem) orow[c0 + 16 + e] = vx[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e < rem) orow[c0 + e] = v[e];
        }
      };
      auto drained = [&]() {                                         // this thread's last read of the accumulator is done
        tc_fence_before();
        mbar_arrive(CT_BAR(kC2DEmpty + (i & 1)));
      };
      auto load_chunk = [&](int ch, uint32_t (&r)[16]) {
        tmem_ld16_issue(dcol + 16 * ch, r);
        if (ride && ch == full - 1) tmem_ld4_issue(dcol + 16 * full, rx);
      };
      auto wait_chunk = [&](int ch, uint32_t (&r)[16]) {
        tmem_ld16_wait(r);
        if (ride && ch == full - 1) tmem_ld4_wait(rx);
      };
      if (n_chunks == 0) drained();                                  // nothing to read for this group
      else {
        load_chunk(0, r0);
        wait_chunk(0, r0);
      }
#pragma unroll 1
      for (int it = 0; it < n_chunks; it += 2) {                     // two chunks per trip: the register sets alternate
        const bool second = it + 1 < n_chunks, third = it + 2 < n_chunks;
        if (second) load_chunk(it + 1, r1);                          // in flight during the activations
        else drained();
        finish_chunk(r0, it);
        if (second) {
          wait_chunk(it + 1, r1);
          if (third) load_chunk(it + 2, r0);
          else drained();
          finish_chunk(r1, it + 1);
          if (third) wait_chunk(it + 2, r0);
        }
      }
      const int64_t R0 = static_cast<int64_t>(t) * kCtM + quarter * 32;
      const int rows_here = static_cast<int>(std::min<int64_t>(32, total_rows - R0));
      if (rows_here > 0 && ((static_cast<uint32_t>(rows_here) * a.filters * 4u) & 15u) != 0) {
        // a ragged last quarter (its byte count is no multiple of 16): plain stores of this thread's own columns
        float *grow = a.out + (R0 + lane) * a.filters;
        if (lane < rows_here)
          for (int c = cbeg; c < cend && c < a.filters; ++c) grow[c] = orow[c];
      }
      fence_async_smem();                                            // this thread's tile values -> visible to the TMA engine
      mbar_arrive(CT_BAR(XS ? kC3OFull + slot : kC2OFull + 2 * quarter + (i & 1)));
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
// contractions stay on the SIMT kernels), filters <= 256, and enough rows to fill the machine.  Odd filter counts
// (the reference's 129) run on conv1d_tc2_kernel (bulk-stored output tiles); SEPCORE_CONV_TC1=1 keeps the first kernel.
int conv1d_tc_try(const float *d_x, const float *d_w, const float *d_b, int batch, int rows, int c_in, int taps,
                  int filters, int stride, int left, int rows_out, int act, float *d_out, cudaStream_t stream,
                  bool *handled) {
  *handled = false;
  constexpr int KQ = 20;
  const int K = taps * c_in, NP = (filters + 15) / 16 * 16;
  if (K != 4 * KQ || NP > 256 || static_cast<int64_t>(batch) * rows_out < 4096) return SEP_OK;
  ConvTcArgs a{d_x, d_w, d_b, d_out, batch, rows, c_in, taps, filters, stride, left, rows_out, act, NP};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const char *v1_env = getenv("SEPCORE_CONV_TC1");
  const bool v1_only = v1_env && atoi(v1_env) != 0;
  const size_t smem2 = 2 * static_cast<size_t>(KQ) * (NP / 8) * 128 + 2 * static_cast<size_t>(kCtM) * filters * 4 +
                       static_cast<size_t>(NP + 32) * 4 + 8 * kC2Bars + 64;
  if (!v1_only && (filters & 1) && NP >= 32 && 2 * NP + 2 * K <= 512 && smem2 <= 227 * 1024 &&
      (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
    *handled = true;
    // the reference geometry (two taps, stride 1, 'same' with the pad on the right) takes the x tile by one bulk copy
    const char *xs_env = getenv("SEPCORE_CONV_XS");
    const bool xs = !(xs_env && atoi(xs_env) == 0) && taps == 2 && stride == 1 && left == 0 && rows_out == rows &&
                    (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (c_in * 4) % 16 == 0;
    const size_t smem3 = 2 * static_cast<size_t>(KQ) * (NP / 8) * 128 + kC3Slots * 32 * static_cast<size_t>(filters) * 4 +
                         ((kCtM + 1) * static_cast<size_t>(c_in) * 4 + 127) / 128 * 128 + static_cast<size_t>(NP + 32) * 4 +
                         8 * kC3Bars + 64;
    void (*kern)(ConvTcArgs);
    if (xs) kern = act == SEP_ACT_SIGMOID ? conv1d_tc2_kernel<KQ, SEP_ACT_SIGMOID, true>
                   : act == SEP_ACT_RELU  ? conv1d_tc2_kernel<KQ, SEP_ACT_RELU, true>
                                          : conv1d_tc2_kernel<KQ, SEP_ACT_LINEAR, true>;
    else kern = act == SEP_ACT_SIGMOID ? conv1d_tc2_kernel<KQ, SEP_ACT_SIGMOID, false>
                : act == SEP_ACT_RELU  ? conv1d_tc2_kernel<KQ, SEP_ACT_RELU, false>
                                       : conv1d_tc2_kernel<KQ, SEP_ACT_LINEAR, false>;
    const size_t smem_k = xs ? smem3 : smem2;
    SEP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_k)));
    const int64_t tiles = (static_cast<int64_t>(batch) * rows_out + kCtM - 1) / kCtM;
    const int grid = static_cast<int>(std::min<int64_t>(tiles, sms));
    profile_begin(stream, "conv1d_tc2_kernel<K=%d%s> (tcgen05 kind::tf32 x3, A in TMEM, M=128 N=%d, output tiles by "
                  "cp.async.bulk; taps=%d c_in=%d filters=%d stride=%d)", K, xs ? ",XS" : "", NP, taps, c_in, filters, stride);
    kern<<<grid, kCt2Threads, smem_k, stream>>>(a);
    profile_end(stream);
    SEP_LAUNCHED();
    return SEP_OK;
  }
  const size_t smem = 2 * static_cast<size_t>(KQ + 2) * (kCtM / 8) * 128 + 2 * static_cast<size_t>(KQ + 2) * (NP / 8) * 128 +
                      8 * 32 * 17 * sizeof(float) + 128;
  if (smem > 225 * 1024) return SEP_OK;
  *handled = true;
  SEP_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
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
