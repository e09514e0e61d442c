// filterbank.cu -- learned 1-D conv encoder / mask / transposed-conv decoder on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), BASELINE config 5:
// N = 256 filters, L = 16 taps, stride 8.
//
// The reference's Raw_with_Convlayer (Raw_with_Convlayer.ipynb:389, cell 13) is the
// encoder half of this idea on non-overlapping segments (served by sep_conv1d_f32);
// BASELINE.json generalises it to the Conv-TasNet shape with a decoder, which the
// reference does not have.  Semantics (oracle/signal_path.py: filterbank_separate):
//   frames[k]  = wave[k*8 : k*8+16]                   K = (n - 16)/8 + 1
//   code       = relu(frames @ enc)                   [K, 256]
//   est_c      = overlap_add((code * mask_c) @ dec)   hop 8, length (K-1)*8 + 16
//
// One CTA = 128 consecutive frames of one utterance = the M = 128 rows of the MMAs;
// thread m of the CTA owns frame row m (TMEM lane m).
//   GEMM 1  D1[128 x 256] = A1[128 x 16] * enc           (2 k-steps of K = 8, tf32)
//   epilogue: D1 -> registers (tcgen05.ld), relu, times mask_c, split, -> smem as the
//             A operand of GEMM 2, 32 columns at a time, double buffered
//   GEMM 2  D2[128 x 16]  += A2[128 x 32] * dec[32 x 16] (4 k-steps per chunk)
//   epilogue: D2 -> registers -> overlap-add of neighbouring frames -> est
// The masked code never touches HBM.  fp32 accuracy on a tf32 pipe: every operand is
// split x = hi + lo with hi exactly representable in tf32, and hi*hi + lo*hi + hi*lo
// is accumulated (3xTF32), so results agree with an fp32 reference to ~1e-6.
//
// Operands live in shared memory in the canonical K-major, no-swizzle UMMA layout:
// 8-row x 16-byte core matrices; LBO = distance between the two 16-byte K chunks of one
// MMA, SBO = distance between 8-row groups (cute/arch/mma_sm100_desc.hpp).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace sep {

constexpr int kFbM = 128;        // frames per tile (MMA M)
constexpr int kFbN = 256;        // filters
constexpr int kFbL = 16;         // taps
constexpr int kFbHop = 8;
constexpr int kFbChunk = 16;     // code columns per decoder work item
constexpr int kFbWG = 4;         // warpgroups per CTA; warpgroup w takes work items w, w + 4, ...
constexpr int kFbThreads = 128 * kFbWG;
constexpr int kFbStages = 2;     // mask tiles in flight per warpgroup (TMA ring, one mbarrier per slot)

// ---- shared memory map (bytes) ----
constexpr int kA1Bytes = kFbM * kFbL * 4;            // 8 KB   [4 k-chunks][16 row groups][128 B]
constexpr int kB1Bytes = kFbN * kFbL * 4;            // 16 KB  [4][32][128 B]
constexpr int kB2Bytes = kFbL * kFbN * 4;            // 16 KB  [64 k-chunks][2][128 B]
constexpr int kA2Bytes = kFbM * kFbChunk * 4;        // 8 KB   [4 k-chunks][16][128 B]
constexpr int kMaskTile = kFbM * kFbChunk * 4;       // 8 KB   TMA box [128 rows][64 B], SWIZZLE_64B
constexpr int kOffA1Hi = 0, kOffA1Lo = kOffA1Hi + kA1Bytes;
constexpr int kOffB1Hi = kOffA1Lo + kA1Bytes, kOffB1Lo = kOffB1Hi + kB1Bytes;
constexpr int kOffB2Hi = kOffB1Lo + kB1Bytes, kOffB2Lo = kOffB2Hi + kB2Bytes;
constexpr int kOffWG = kOffB2Lo + kB2Bytes;          // per warpgroup: A2 hi, A2 lo, mask ring, `up`
constexpr int kWGBytes = 2 * kA2Bytes + kFbStages * kMaskTile + kFbM * 8 * 4;
constexpr int kOffBar = kOffWG + kFbWG * kWGBytes;   // mbarriers + tmem address + has[][] table
constexpr int kFbSmem = kOffBar + 256;

constexpr uint32_t kLboA1 = 16 * 128, kLboB1 = 32 * 128, kLboB2 = 2 * 128, kLboA2 = 16 * 128, kSbo = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);   // version 1, SWIZZLE_NONE
}
// kind::tf32, D = F32, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// 2-D tiled bulk tensor copy global -> shared, completion on an mbarrier (TMA)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&d)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&d)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = __uint_as_float(r[i]);
}

// x = hi + lo with hi exactly representable in tf32 (10-bit mantissa: low 13 bits clear)
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// byte offset of element (row, k) in a K-major no-swizzle operand with `groups` 8-row groups
__device__ __forceinline__ int kmajor_off(int row, int k, int groups) {
  return (k >> 2) * groups * 128 + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4;
}

struct FbArgs {
  const float *wave, *enc, *dec, *masks;
  float *est, *code;
  int64_t n, est_len;
  int n_src, frames, tiles, batch;
};

__device__ __forceinline__ void wg_sync(int wg) {       // named barrier of one warpgroup
  asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
}

// One persistent CTA per SM, four warpgroups.  Per tile of 128 frames:
//   all:        frames -> A1 (warpgroup 0), GEMM 1 into D1 (TMEM columns 0..255)
//   warpgroup w: work items g = w, w + 4, ... (item = 16 code columns j of source c):
//               D1 chunk -> registers, relu, times its thread-private staged mask row, split hi/lo
//               -> the warpgroup's A2 buffer -> its elected thread issues the 6 decoder MMAs into
//               the warpgroup's own accumulator D2[w][c]; the mask tile of the next item is in flight
//   epilogue:   warpgroup c sums D2[*][c], overlap-adds neighbouring frames, writes est_c
// Warpgroups only meet at two block barriers per tile; inside a tile they run on named barriers.
__global__ void __launch_bounds__(kFbThreads, 1) filterbank_kernel(const FbArgs a,
                                                                   const __grid_constant__ CUtensorMap mask_map) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int wg = threadIdx.x >> 7, m = threadIdx.x & 127, wq = (threadIdx.x >> 5) & 3;
  const int K = a.frames, C = a.n_src;
  // mbarriers: [0] gemm1 done, [1 + w] A2 buffer of warpgroup w free, [1 + kFbWG] all decoder MMAs done
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + kOffBar);
  // [8 + 2 w + s] mask slot s of warpgroup w full (TMA transaction barrier)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + kOffBar + 192);
  unsigned char *has = sm + kOffBar + 224;                         // has[w * 4 + c]: warpgroup w feeds source c
  const uint32_t bar0 = smem_u32(bars), sm0 = smem_u32(sm);
  unsigned char *wgs = sm + kOffWG + wg * kWGBytes;
  unsigned char *a2hi = wgs, *a2lo = wgs + kA2Bytes, *ring = wgs + 2 * kA2Bytes;
  float *up = reinterpret_cast<float *>(wgs + 2 * kA2Bytes + kFbStages * kMaskTile);
  const int G = (kFbN / kFbChunk) * C;

  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    for (int w = 0; w < kFbWG; ++w) mbar_init(bar0 + 8 * (1 + w), 1);
    mbar_init(bar0 + 8 * (1 + kFbWG), kFbWG);
    for (int i = 0; i < kFbWG * kFbStages; ++i) mbar_init(bar0 + 8 * (8 + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 16; ++i) has[i] = 0;
    for (int g = 0; g < G; ++g) has[(g % kFbWG) * 4 + g % C] = 1;
  }

  // ---- operands of GEMM 1 and the decoder weights, split hi / lo, canonical layout (once per CTA) ----
#pragma unroll 1
  for (int e0 = 0; e0 < kFbL * kFbN; e0 += kFbThreads * 4) {
    float we[4], wd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      we[u] = __ldg(a.enc + e0 + kFbThreads * u + threadIdx.x);
      wd[u] = __ldg(a.dec + e0 + kFbThreads * u + threadIdx.x);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // enc [L][N] row-major -> B1[n][k = l];  dec [N][L] row-major -> B2[n2 = l][k = n]
      const int e = e0 + kFbThreads * u + threadIdx.x;
      const int l = e / kFbN, nn = e % kFbN;
      float hi, lo;
      split_tf32(we[u], hi, lo);
      const int o1 = kmajor_off(nn, l, kFbN / 8);
      *reinterpret_cast<float *>(sm + kOffB1Hi + o1) = hi;
      *reinterpret_cast<float *>(sm + kOffB1Lo + o1) = lo;
      const int nf = e / kFbL, ll = e % kFbL;
      split_tf32(wd[u], hi, lo);
      const int o2 = kmajor_off(ll, nf, kFbL / 8);
      *reinterpret_cast<float *>(sm + kOffB2Hi + o2) = hi;
      *reinterpret_cast<float *>(sm + kOffB2Lo + o2) = lo;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
  uint32_t use = 0;                                    // how often this warpgroup's A2 buffer has been filled
  uint32_t fills = 0;                                  // mask tiles this warpgroup has consumed (slot + phase)
  const uint32_t ring_u32 = smem_u32(ring), full0 = bar0 + 8 * (8 + kFbStages * wg);
  // TMA of the mask tile of work item g of tile (b, k0) into ring slot `slot` (issued by one thread)
  auto issue_masks = [&](int b, int k0, int g, uint32_t slot) {
    const int j = g / C, c = g - j * C;
    mbar_expect_tx(full0 + 8 * slot, kMaskTile);
    tma_load_2d(ring_u32 + slot * kMaskTile, &mask_map, j * kFbChunk, (b * C + c) * K + k0, full0 + 8 * slot);
  };
  if (m == 0 && blockIdx.x < a.tiles * a.batch) {
    const int b = blockIdx.x / a.tiles, tile = blockIdx.x - b * a.tiles;
    issue_masks(b, tile * (kFbM - 1), wg, 0);
  }
  uint32_t round = 0;                                  // tiles done by this CTA (mbarrier phase)

  for (int t = blockIdx.x; t < a.tiles * a.batch; t += gridDim.x, ++round) {
    const int b = t / a.tiles, tile = t - b * a.tiles;
    const int k0 = tile * (kFbM - 1);                     // first frame of the tile (1-frame halo)
    const int frame = k0 + m;
    const bool row_ok = frame < K;
    const bool owner = row_ok && (m > 0 || tile == 0);    // row 0 of a later tile is the halo frame
    if (wg == 0) {
      const float *src = a.wave + static_cast<int64_t>(b) * a.n + static_cast<int64_t>(frame) * kFbHop;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                    // 4 k-chunks of 4 taps
        float4 x = row_ok ? __ldg(reinterpret_cast<const float4 *>(src) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
        split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        const int o = q * (kFbM / 8) * 128 + (m >> 3) * 128 + (m & 7) * 16;
        *reinterpret_cast<float4 *>(sm + kOffA1Hi + o) = hi;
        *reinterpret_cast<float4 *>(sm + kOffA1Lo + o) = lo;
      }
      fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();                                     // A1 ready; every read of D1 / D2 / `up` of the last tile is done
    tc_fence_after();

    // ---- GEMM 1: D1 = A1 * enc, 3xTF32 ----
    if (threadIdx.x == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kFbM, kFbN);
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {           // hi*hi, lo*hi, hi*lo
        const uint32_t aoff = sm0 + (pass == 1 ? kOffA1Lo : kOffA1Hi);
        const uint32_t boff = sm0 + (pass == 2 ? kOffB1Lo : kOffB1Hi);
#pragma unroll
        for (int ks = 0; ks < kFbL / 8; ++ks) {
          umma_tf32(tmem, umma_desc(aoff + ks * 2 * kLboA1, kLboA1, kSbo),
                    umma_desc(boff + ks * 2 * kLboB1, kLboB1, kSbo), idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar0);
    }
    mbar_wait(bar0, round & 1);
    tc_fence_after();

    // ---- this warpgroup's work items ----
    uint32_t touched = 0;                               // sources whose accumulator D2[wg][c] has been started
    int i = 0;
    for (int g = wg; g < G; g += kFbWG, ++i) {
      const int j = g / C, c = g - j * C;
      float d[kFbChunk];
      tmem_ld16(lane_addr + j * kFbChunk, d);
#pragma unroll
      for (int e = 0; e < kFbChunk; ++e) d[e] = fmaxf(d[e], 0.f);
      if (a.code && owner && c == 0) {
        float4 *dst = reinterpret_cast<float4 *>(a.code + (static_cast<int64_t>(b) * K + frame) * kFbN + j * kFbChunk);
#pragma unroll
        for (int q = 0; q < kFbChunk / 4; ++q) dst[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
      }
      // next item's mask tile (of this tile, or the first item of this CTA's next tile) into the other
      // slot: every thread of the warpgroup passed the named barrier of the item that last read it
      const uint32_t slot = fills % kFbStages;
      if (m == 0) {
        if (g + kFbWG < G) issue_masks(b, k0, g + kFbWG, (fills + 1) % kFbStages);
        else if (t + static_cast<int>(gridDim.x) < a.tiles * a.batch) {
          const int tn = t + gridDim.x, bn = tn / a.tiles;
          issue_masks(bn, (tn - bn * a.tiles) * (kFbM - 1), wg, (fills + 1) % kFbStages);
        }
      }
      mbar_wait(full0 + 8 * slot, (fills / kFbStages) & 1);
      ++fills;
      if (use > 0) mbar_wait(bar0 + 8 * (1 + wg), (use - 1) & 1);   // the MMAs that read the A2 buffer are done
      // SWIZZLE_64B: 16-byte chunk q of row m sits at chunk q ^ ((m >> 1) & 3)
      const unsigned char *mk_s = ring + slot * kMaskTile + m * 64;
      const int sw = (m >> 1) & 3;
#pragma unroll
      for (int q = 0; q < kFbChunk / 4; ++q) {
        const float4 mk = row_ok ? *reinterpret_cast<const float4 *>(mk_s + ((q ^ sw) << 4))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 hi, lo;
        split_tf32(d[4 * q] * mk.x, hi.x, lo.x); split_tf32(d[4 * q + 1] * mk.y, hi.y, lo.y);
        split_tf32(d[4 * q + 2] * mk.z, hi.z, lo.z); split_tf32(d[4 * q + 3] * mk.w, hi.w, lo.w);
        const int o = q * (kFbM / 8) * 128 + (m >> 3) * 128 + (m & 7) * 16;
        *reinterpret_cast<float4 *>(a2hi + o) = hi;
        *reinterpret_cast<float4 *>(a2lo + o) = lo;
      }
      ++use;
      fence_async_smem();
      tc_fence_before();
      wg_sync(wg);
      tc_fence_after();
      if (m == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(kFbM, kFbL);
        const uint32_t a2 = smem_u32(a2hi);
        const uint32_t dcol = tmem + kFbN + kFbL * (wg * C + c);
        uint32_t acc = (touched >> c) & 1u;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t aoff = a2 + (pass == 1 ? kA2Bytes : 0);
          const uint32_t boff = sm0 + (pass == 2 ? kOffB2Lo : kOffB2Hi) + j * (kFbChunk / 4) * kLboB2;
#pragma unroll
          for (int ks = 0; ks < kFbChunk / 8; ++ks) {
            umma_tf32(dcol, umma_desc(aoff + ks * 2 * kLboA2, kLboA2, kSbo),
                      umma_desc(boff + ks * 2 * kLboB2, kLboB2, kSbo), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(bar0 + 8 * (1 + wg));
      }
      touched |= 1u << c;
    }
    if (m == 0) umma_commit(bar0 + 8 * (1 + kFbWG));      // this warpgroup's decoder MMAs, all of them
    mbar_wait(bar0 + 8 * (1 + kFbWG), round & 1);
    tc_fence_after();

    // ---- sum of the warpgroups' D2_c -> overlap-add of neighbouring frames -> est ----
    for (int c = wg; c < C; c += kFbWG) {
      float y[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) y[e] = 0.f;
      for (int w = 0; w < kFbWG; ++w) {
        if (!has[w * 4 + c]) continue;
        float p[16];
        tmem_ld16(lane_addr + kFbN + kFbL * (w * C + c), p);
#pragma unroll
        for (int e = 0; e < 16; ++e) y[e] += p[e];
      }
      // hop-block h = frame gets y[frame][0:8] + y[frame-1][8:16]
#pragma unroll
      for (int e = 0; e < 8; ++e) up[m * 8 + e] = y[8 + e];
      wg_sync(wg);
      float *out = a.est + (static_cast<int64_t>(b) * C + c) * a.est_len;
      if (owner) {
        float4 lo4 = make_float4(y[0], y[1], y[2], y[3]), hi4 = make_float4(y[4], y[5], y[6], y[7]);
        if (m > 0) {
          const float *p = up + (m - 1) * 8;
          lo4.x += p[0]; lo4.y += p[1]; lo4.z += p[2]; lo4.w += p[3];
          hi4.x += p[4]; hi4.y += p[5]; hi4.z += p[6]; hi4.w += p[7];
        }
        float4 *dst = reinterpret_cast<float4 *>(out + static_cast<int64_t>(frame) * kFbHop);
        dst[0] = lo4;
        dst[1] = hi4;
        if (frame == K - 1) {                              // the tail hop-block of the utterance
          dst[2] = make_float4(y[8], y[9], y[10], y[11]);
          dst[3] = make_float4(y[12], y[13], y[14], y[15]);
        }
      }
      wg_sync(wg);                                        // `up` is reused by this warpgroup's next source
    }
    tc_fence_before();                                    // D1 / D2 are overwritten by the next tile
  }  // tile loop

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

// Tensor map of the masks seen as [rows = B * C * K][256] float32, box [128 rows][16 columns],
// SWIZZLE_64B.  The driver entry point is resolved at run time (no link against libcuda).
static int make_mask_map(CUtensorMap *map, const float *masks, uint64_t rows) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SEP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return SEP_ERR_CUDA;
    }
    encode = reinterpret_cast<encode_fn>(fn);
  }
  if (reinterpret_cast<uintptr_t>(masks) & 15) {
    set_error("sep_filterbank_separate_f32: masks must be 16-byte aligned");
    return SEP_ERR_INVALID;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kFbN), rows};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kFbN) * 4};
  const cuuint32_t box[2] = {kFbChunk, kFbM}, elem[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(masks), dims, strides, box,
                            elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return SEP_ERR_CUDA;
  }
  return SEP_OK;
}

}  // namespace sep

using namespace sep;

extern "C" int sep_filterbank_separate_f32(const float *wave, const float *enc, const float *dec,
                                           const float *masks, int batch, int n_src, int64_t n_samples,
                                           int taps, int n_filters, int stride, float *est, float *code,
                                           int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(wave && enc && dec && masks && est, "sep_filterbank_separate_f32: null argument");
  SEP_REQUIRE(batch >= 1 && n_src >= 1 && n_src <= SEP_MAX_SOURCES && n_samples >= taps,
              "sep_filterbank_separate_f32: bad shape");
  if (taps != kFbL || n_filters != kFbN || stride != kFbHop) {
    set_error("sep_filterbank_separate_f32: only L=%d, N=%d, stride=%d is built (got L=%d N=%d stride=%d)",
              kFbL, kFbN, kFbHop, taps, n_filters, stride);
    return SEP_ERR_UNSUPPORTED;
  }
  SEP_REQUIRE(n_samples % 4 == 0, "sep_filterbank_separate_f32: n_samples must be a multiple of 4");
  int rc = check_mem(mem);
  if (rc) return rc;
  const int64_t K = (n_samples - taps) / stride + 1;
  const int64_t est_len = (K - 1) * stride + taps;
  Scratch s(stream);
  FbArgs a{};
  const size_t n_mask = static_cast<size_t>(batch) * n_src * K * n_filters;
  const size_t n_est = static_cast<size_t>(batch) * n_src * est_len, n_code = static_cast<size_t>(batch) * K * n_filters;
  if ((rc = stage_in(s, wave, static_cast<size_t>(batch) * n_samples, mem, &a.wave))) return rc;
  if ((rc = stage_in(s, enc, static_cast<size_t>(taps) * n_filters, mem, &a.enc))) return rc;
  if ((rc = stage_in(s, dec, static_cast<size_t>(taps) * n_filters, mem, &a.dec))) return rc;
  if ((rc = stage_in(s, masks, n_mask, mem, &a.masks))) return rc;
  if ((rc = stage_out(s, est, n_est, mem, &a.est))) return rc;
  if ((rc = stage_out(s, code, n_code, mem, &a.code))) return rc;
  a.n = n_samples;
  a.est_len = est_len;
  a.n_src = n_src;
  a.batch = batch;
  a.frames = static_cast<int>(K);
  // tile t owns frames 127 t + 1 .. 127 t + 127 (and frame 0 for t = 0)
  a.tiles = static_cast<int>(std::max<int64_t>(1, (K - 1 + kFbM - 2) / (kFbM - 1)));
  SEP_CUDA(cudaFuncSetAttribute(filterbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFbSmem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // persistent: one CTA per SM (its 200+ KB of shared memory and 512 TMEM columns fill the SM)
  dim3 grid(static_cast<unsigned>(std::min<int64_t>(static_cast<int64_t>(a.tiles) * batch, sms)));
  CUtensorMap mask_map;
  if ((rc = make_mask_map(&mask_map, a.masks, static_cast<uint64_t>(batch) * n_src * K))) return rc;
  profile_begin(stream);
  filterbank_kernel<<<grid, kFbThreads, kFbSmem, stream>>>(a, mask_map);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, est, a.est, n_est, mem))) return rc;
  if ((rc = copy_back(s, code, a.code, n_code, mem))) return rc;
  return finish(s, mem);
}
