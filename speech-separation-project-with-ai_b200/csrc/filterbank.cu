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
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace sep {

constexpr int kFbM = 128;        // frames per tile (MMA M)
constexpr int kFbN = 256;        // filters
constexpr int kFbL = 16;         // taps
constexpr int kFbHop = 8;
constexpr int kFbChunk = 16;     // code columns per decoder work item
constexpr int kFbWG = 4;         // consumer warpgroups per CTA; warpgroup w works for source w mod C
constexpr int kFbThreads = 128 * kFbWG + 32 * kFbWG + 32;  // + one MMA-issue warp per warpgroup + a frame-staging warp
constexpr int kFbStages = 4;     // mask tiles in flight per warpgroup (TMA ring, one mbarrier pair per slot)

// ---- shared memory map (bytes) ----
constexpr int kA1Bytes = kFbM * kFbL * 4;            // 8 KB   [4 k-chunks][16 row groups][128 B]
constexpr int kB1Bytes = kFbN * kFbL * 4;            // 16 KB  [4][32][128 B]
constexpr int kB2Bytes = kFbL * kFbN * 4;            // 16 KB  [64 k-chunks][2][128 B]
constexpr int kA2Bytes = kFbM * kFbChunk * 4;        // 8 KB   [4 k-chunks][16][128 B]
constexpr int kMaskTile = kFbM * kFbChunk * 4;       // 8 KB   TMA box [128 rows][64 B], SWIZZLE_64B
constexpr int kOffA1Hi = 0, kOffA1Lo = kOffA1Hi + kA1Bytes;
constexpr int kOffB1Hi = kOffA1Lo + kA1Bytes, kOffB1Lo = kOffB1Hi + kB1Bytes;
constexpr int kOffB2Hi = kOffB1Lo + kB1Bytes, kOffB2Lo = kOffB2Hi + kB2Bytes;
constexpr int kOffWG = kOffB2Lo + kB2Bytes;          // per warpgroup: mask ring, `up`
constexpr int kWGBytes = kFbStages * kMaskTile + kFbM * 8 * 4;
// tensor memory columns: D1 [0, 256); D2 [256, 384): 32 columns per WARPGROUP (its partial sum over its code columns:
// hi*hi + lo*hi | hi*lo); masked code A2 of warpgroup w: two 8-column buffers at 384 + 32 w + 16 h (hi | lo)
constexpr int kTmD2 = kFbN, kTmA2 = kFbN + 128;
constexpr int kOffBar = kOffWG + kFbWG * kWGBytes;   // mbarriers + tmem address + has[][] table
constexpr int kFbSmem = kOffBar + 768;

constexpr uint32_t kLboA1 = 16 * 128, kLboB1 = 32 * 128, kLboB2 = 4 * 128, kSbo = 128;

struct FbArgs {
  const float *wave, *enc, *dec, *masks;
  float *est, *code;
  int64_t n, est_len;
  int n_src, frames, tiles, batch;
};

__device__ __forceinline__ void wg_sync(int wg) {       // named barrier of one warpgroup
  asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
}

// One persistent CTA per SM, warp-specialised:
//   warpgroups 0..3 (consumers): warpgroup w < n_active works for source w mod C on the 16-column code chunks
//       j = w / C, w / C + n_active / C, ...: D1 chunk -> registers (tcgen05.ld), relu, times the mask tile in shared
//       memory, split hi / lo -> its A2 buffer -> arrive on `a2_ready`; then the epilogue of the tile
//   warp 16 + w (MMA issue for warpgroup w; the warp runs converged, one elected lane issues): the decoder MMAs of
//       every half item of ITS warpgroup into that warpgroup's own accumulator D2[w], and the TMA refill of the mask
//       slot the half item came from (a2_ready of an item implies that all 128 threads have read its mask tile); warp
//       16 also issues GEMM 1 of every tile.  tcgen05.mma costs ~60-100 issue cycles per instruction from one warp
//       whatever its size (descriptor and address moves into uniform registers), and the decoder is 128 small MMAs per
//       tile: four issuing warps, none waiting for another warpgroup's hand-off (round 1/2 first version: two issuers
//       for two warpgroups each, 147.6 -> 118 us with converged issue, -> this layout)
//   warp 20 (frame staging): the next tile's frames -> A1
// Nobody meets at a block barrier inside the tile loop; all hand-offs are mbarriers.
enum : int {
  kBarG1 = 0,            // GEMM 1 of the tile done (commit)
  kBarD2 = 1,            // all decoder MMAs of the tile done (one commit per active issuer)
  kBarA1 = 2,            // frames of the tile staged (32 producer lanes)
  kBarD1Free = 3,        // every active consumer has read its last D1 chunk (128 n_active)
  kBarD2Free = 4,        // the epilogue has read D2 (128 per source)
  kBarA2Ready = 48,      // [2 w + h] half h (8 code columns) of an item written to A2 buffer h (128)
  kBarA2Free = 56,       // [2 w + h] the decoder MMAs that read A2 buffer h are done (commit)
  kBarFull = 16,         // [w * stages + s] mask tile landed (TMA transaction)
};

__global__ void __launch_bounds__(kFbThreads, 1) filterbank_kernel(const FbArgs a,
                                                                   const __grid_constant__ CUtensorMap mask_map) {
  extern __shared__ __align__(128) unsigned char sm[];
  // the warp index through a shuffle: ptxas then KNOWS that it is warp-uniform, so the role branches are uniform and the
  // addresses / descriptors of the MMA warps can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int wg = threadIdx.x >> 7, m = threadIdx.x & 127, wq = warp & 3;
  const int K = a.frames, C = a.n_src;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + kOffBar);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + kOffBar + 640);
  const uint32_t bar0 = smem_u32(bars), sm0 = smem_u32(sm);
#define SEP_BAR(i) (bar0 + 8u * static_cast<uint32_t>(i))
  // warpgroups at work: the largest multiple of C that fits (C = 3: three); each serves ONE source
  const int n_active = C >= kFbWG ? kFbWG : (kFbWG / C) * C, per_src = n_active / (C < kFbWG ? C : kFbWG);
  const int n_items = (kFbN / kFbChunk) / per_src;          // 16-column chunks per warpgroup and tile
  const int n_tiles = a.tiles * a.batch;

  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) {
    mbar_init(SEP_BAR(kBarG1), 1);
    mbar_init(SEP_BAR(kBarD2), n_active);
    mbar_init(SEP_BAR(kBarA1), 32);
    mbar_init(SEP_BAR(kBarD1Free), 128 * n_active);
    mbar_init(SEP_BAR(kBarD2Free), 128 * (C < kFbWG ? C : kFbWG));
    for (int w = 0; w < kFbWG; ++w) {
      for (int h = 0; h < 2; ++h) {
        mbar_init(SEP_BAR(kBarA2Ready + 2 * w + h), 128);
        mbar_init(SEP_BAR(kBarA2Free + 2 * w + h), 1);
      }
      for (int st = 0; st < kFbStages; ++st) mbar_init(SEP_BAR(kBarFull + w * kFbStages + st), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- encoder and decoder weights, split hi / lo, canonical layout (once per CTA) ----
#pragma unroll 1
  for (int e = threadIdx.x; e < kFbL * kFbN; e += kFbThreads) {
    // enc [L][N] row-major -> B1[n][k = l];  dec [N][L] row-major -> B2[n2 = l][k = n]
    const int l = e / kFbN, nn = e % kFbN;
    float hi, lo;
    split_tf32(__ldg(a.enc + e), hi, lo);
    const int o1 = kmajor_off(nn, l, kFbN / 8);
    *reinterpret_cast<float *>(sm + kOffB1Hi + o1) = hi;
    *reinterpret_cast<float *>(sm + kOffB1Lo + o1) = lo;
    const int nf = e / kFbL, ll = e % kFbL;
    // one decoder operand with N = 32 rows: rows 0..15 = hi, rows 16..31 = lo, so that A_hi * [hi | lo]
    // is ONE MMA (N = 32) and A_lo * hi a second one (N = 16, the first two row groups)
    split_tf32(__ldg(a.dec + e), hi, lo);
    *reinterpret_cast<float *>(sm + kOffB2Hi + kmajor_off(ll, nf, 2 * kFbL / 8)) = hi;
    *reinterpret_cast<float *>(sm + kOffB2Hi + kmajor_off(kFbL + ll, nf, 2 * kFbL / 8)) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 4 * kFbWG && warp < 5 * kFbWG) {
    // =========================== MMA warps (converged; one elected lane issues) ===========================
    const int w = warp - 4 * kFbWG;                        // the warpgroup this warp issues for
    if (w < n_active || w == 0) {
      const int c = w % C, j0 = w / C;                     // its source, its first chunk
      uint32_t served = 0;                                 // half items served, all tiles (phase of a2_ready)
      uint32_t pf = 0;                                     // mask tiles requested, all tiles
      int pf_t = blockIdx.x, pf_i = 0;                     // the next mask tile to request: tile, item
      auto request = [&]() {                               // one lane: arm the slot's barrier, start the TMA
        if (pf_t < n_tiles) {
          if (lane == 0) {
            const int b = pf_t / a.tiles, k0 = (pf_t - b * a.tiles) * (kFbM - 1);
            const uint32_t full = SEP_BAR(kBarFull + w * kFbStages + pf % kFbStages);
            mbar_expect_tx(full, kMaskTile);
            tma_load_2d(sm0 + kOffWG + w * kWGBytes + (pf % kFbStages) * kMaskTile, &mask_map,
                        (j0 + pf_i * per_src) * kFbChunk, (b * C + c) * K + k0, full);
          }
          __syncwarp();
          ++pf;
          if (++pf_i == n_items) { pf_i = 0; pf_t += gridDim.x; }
        }
      };
      if (w < n_active)
        for (int st = 0; st < kFbStages; ++st) request();
      uint32_t round = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++round) {
        if (w == 0) {
          mbar_wait_suspend(SEP_BAR(kBarA1), round & 1);                             // frames staged
          if (round > 0) mbar_wait_suspend(SEP_BAR(kBarD1Free), (round - 1) & 1);      // D1 of the last tile drained
          tc_fence_after();
          constexpr uint32_t idesc = umma_idesc_tf32(kFbM, kFbN);
          uint32_t acc = 0;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {         // hi*hi, lo*hi, hi*lo
            const uint32_t aoff = sm0 + (pass == 1 ? kOffA1Lo : kOffA1Hi);
            const uint32_t boff = sm0 + (pass == 2 ? kOffB1Lo : kOffB1Hi);
#pragma unroll
            for (int ks = 0; ks < kFbL / 8; ++ks) {
              umma_tf32_elect(tmem, umma_desc(aoff + ks * 2 * kLboA1, kLboA1, kSbo),
                              umma_desc(boff + ks * 2 * kLboB1, kLboB1, kSbo), idesc, acc);
              acc = 1;
            }
          }
          umma_commit_elect(SEP_BAR(kBarG1));
        }
        if (w >= n_active) continue;
        // this warpgroup's accumulator: read by the epilogue of the previous tile
        if (round > 0) {
          mbar_wait_suspend(SEP_BAR(kBarD2Free), (round - 1) & 1);
          tc_fence_after();
        }
        constexpr uint32_t idesc32 = umma_idesc_tf32(kFbM, 2 * kFbL), idesc16 = umma_idesc_tf32(kFbM, kFbL);
        const uint64_t bdesc0 = umma_desc(sm0 + kOffB2Hi, kLboB2, kSbo);
        const uint32_t dcol = tmem + kTmD2 + 2 * kFbL * w;
        for (int i = 0; i < n_items; ++i) {
          const int j = j0 + i * per_src;
#pragma unroll
          for (uint32_t h = 0; h < 2; ++h) {
            mbar_wait_suspend(SEP_BAR(kBarA2Ready + 2 * w + h), (served >> 1) & 1);
            tc_fence_after();
            // D2_w[:, 0:16] += A_hi * B_hi + A_lo * B_hi,  D2_w[:, 16:32] += A_hi * B_lo
            const uint32_t a2 = tmem + kTmA2 + 32 * w + 16 * h;
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(((j * (kFbChunk / 4) + 2 * h) * kLboB2) >> 4);
            umma_tf32_ts_elect(dcol, a2, bdesc, idesc32, (i > 0 || h > 0) ? 1u : 0u);
            umma_tf32_ts_elect(dcol, a2 + 8, bdesc, idesc16, 1u);
            umma_commit_elect(SEP_BAR(kBarA2Free + 2 * w + h));
            ++served;
            // a2_ready of half 0 means that all 128 threads have used the item's mask tile (their products depend
            // on it and were stored before they arrived): its slot takes the tile kFbStages items ahead
            if (h == 0) request();
          }
        }
        umma_commit_elect(SEP_BAR(kBarD2));
      }
    }
  } else if (warp == 5 * kFbWG) {
    // =========================== frame-staging warp: the next tile's frames -> A1 (hi / lo) ===========================
    auto stage_frames = [&](int t) {                     // all 32 lanes: 4 rows each
      const int b = t / a.tiles, k0 = (t - b * a.tiles) * (kFbM - 1);
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int row = lane + 32 * rr, frame = k0 + row;
        const float *src = a.wave + static_cast<int64_t>(b) * a.n + static_cast<int64_t>(frame) * kFbHop;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                    // 4 k-chunks of 4 taps
          const float4 x = frame < K ? __ldg(reinterpret_cast<const float4 *>(src) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 hi, lo;
          split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
          split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
          const int o = q * (kFbM / 8) * 128 + (row >> 3) * 128 + (row & 7) * 16;
          *reinterpret_cast<float4 *>(sm + kOffA1Hi + o) = hi;
          *reinterpret_cast<float4 *>(sm + kOffA1Lo + o) = lo;
        }
      }
      fence_async_smem();
      mbar_arrive(SEP_BAR(kBarA1));
    };
    if (blockIdx.x < n_tiles) stage_frames(blockIdx.x);
    uint32_t round = 0;
    for (int t = blockIdx.x; t + static_cast<int>(gridDim.x) < n_tiles; t += gridDim.x, ++round) {
      mbar_wait_suspend(SEP_BAR(kBarG1), round & 1);               // GEMM 1 of this tile has consumed A1
      stage_frames(t + gridDim.x);
    }
  } else {
    // =========================== consumer warpgroups ===========================
    unsigned char *wgs = sm + kOffWG + wg * kWGBytes;
    unsigned char *ring = wgs;
    float *up = reinterpret_cast<float *>(wgs + kFbStages * kMaskTile);
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    uint32_t use = 0;                                    // items this warpgroup has written to its A2 buffer
    uint32_t fills = 0;                                  // mask tiles this warpgroup has consumed
    uint32_t round = 0;
    const int sw = (m >> 1) & 3;                         // SWIZZLE_64B: chunk q of row m sits at q ^ ((m >> 1) & 3)
    const uint32_t a2_addr = lane_addr + kTmA2 + 32 * wg;        // this thread's row of the warpgroup's A2: buffer h at + 16 h (hi 8 columns, lo 8 columns)
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++round) {
      const int b = t / a.tiles, tile = t - b * a.tiles;
      const int k0 = tile * (kFbM - 1);                     // first frame of the tile (1-frame halo)
      const int frame = k0 + m;
      const bool row_ok = frame < K;
      const bool owner = row_ok && (m > 0 || tile == 0);    // row 0 of a later tile is the halo frame
      mbar_wait_suspend(SEP_BAR(kBarG1), round & 1);
      tc_fence_after();

      const int c = wg % C;                                 // this warpgroup's source
      for (int i = 0; i < n_items && wg < n_active; ++i) {  // its chunks j = wg / C, + per_src, ...
        const int j = wg / C + i * per_src;
        float d[kFbChunk];
        __syncwarp();                                       // tcgen05.ld is warp-collective (.sync.aligned)
        tmem_ld16(lane_addr + j * kFbChunk, d);
        if (i + 1 == n_items) {                             // that was this thread's last read of D1
          tc_fence_before();
          mbar_arrive(SEP_BAR(kBarD1Free));
        }
        // frames beyond the utterance are all-zero rows of A1 (code = relu(0) = 0), but the mask rows the TMA tile
        // covers there belong to the NEXT (b, c) plane (or are the tile's zero fill): a non-finite mask value
        // would turn 0 * inf into NaN, so those rows' products are forced to zero below
#pragma unroll
        for (int e = 0; e < kFbChunk; ++e) d[e] = fmaxf(d[e], 0.f);
        const uint32_t slot = fills % kFbStages;
        mbar_wait_suspend(SEP_BAR(kBarFull + wg * kFbStages + slot), (fills / kFbStages) & 1);
        const unsigned char *mk_s = ring + slot * kMaskTile + m * 64;
        float2 p[kFbChunk / 2];
#pragma unroll
        for (int q = 0; q < kFbChunk / 4; ++q) {
          const float4 mk = *reinterpret_cast<const float4 *>(mk_s + ((q ^ sw) << 4));
          p[2 * q] = __fmul2_rn(make_float2(d[4 * q], d[4 * q + 1]), make_float2(mk.x, mk.y));
          p[2 * q + 1] = __fmul2_rn(make_float2(d[4 * q + 2], d[4 * q + 3]), make_float2(mk.z, mk.w));
        }
        if (!row_ok) {
#pragma unroll
          for (int q = 0; q < kFbChunk / 2; ++q) p[q] = make_float2(0.f, 0.f);
        }
        ++fills;
        float hi[kFbChunk], lo[kFbChunk];
#pragma unroll
        for (int q = 0; q < kFbChunk / 2; ++q) {
          hi[2 * q] = __uint_as_float(__float_as_uint(p[q].x) & 0xFFFFE000u);
          hi[2 * q + 1] = __uint_as_float(__float_as_uint(p[q].y) & 0xFFFFE000u);
          const float2 l2 = __fadd2_rn(p[q], make_float2(-hi[2 * q], -hi[2 * q + 1]));
          lo[2 * q] = l2.x;
          lo[2 * q + 1] = l2.y;
        }
        // The TMA (another proxy) may overwrite the mask slot only after this thread's loads from it have COMPLETED.
        // The values stored to A2 below depend on all four loads, tcgen05.wait::st follows the stores and the arrive
        // on a2_ready follows that: the issuing warp, which refills the slot, sees a2_ready first.
        // two half items (8 code columns each) into the two A2 buffers: the wait is for the MMAs of the
        // previous ITEM's half, committed a whole item ago
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (use > 0) {
            mbar_wait_suspend(SEP_BAR(kBarA2Free + 2 * wg + h), (use - 1) & 1);
            tc_fence_after();
          }
          __syncwarp();
          tmem_st8(a2_addr + 16 * h, hi + 8 * h);
          tmem_st8(a2_addr + 16 * h + 8, lo + 8 * h);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(SEP_BAR(kBarA2Ready + 2 * wg + h));
        }
        ++use;
        if (a.code && owner && c == 0) {                    // optional dump of the unmasked code (tests)
          float4 *dst = reinterpret_cast<float4 *>(a.code + (static_cast<int64_t>(b) * K + frame) * kFbN + j * kFbChunk);
#pragma unroll
          for (int q = 0; q < kFbChunk / 4; ++q) dst[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
        }
      }

      // ---- sum of the warpgroups' D2_c -> overlap-add of neighbouring frames -> est ----
      if (wg < C) {
        mbar_wait_suspend(SEP_BAR(kBarD2), round & 1);
        tc_fence_after();
      }
      for (int cc = wg; cc < C; cc += kFbWG) {
        // est of source cc = the partial sums of the warpgroups that worked for it (w = cc, cc + C, ...), hi | lo halves
        float y[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) y[e] = 0.f;
        for (int w2 = cc; w2 < n_active; w2 += C) {
          float ya[16], yb[16];
          __syncwarp();
          tmem_ld16(lane_addr + kTmD2 + 2 * kFbL * w2, ya);
          tmem_ld16(lane_addr + kTmD2 + 2 * kFbL * w2 + kFbL, yb);
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] += ya[e] + yb[e];
        }
        if (cc + kFbWG >= C) {                              // this thread's last read of D2
          tc_fence_before();
          mbar_arrive(SEP_BAR(kBarD2Free));
        }
        // hop-block h = frame gets y[frame][0:8] + y[frame-1][8:16]
#pragma unroll
        for (int e = 0; e < 8; ++e) up[m * 8 + e] = y[8 + e];
        wg_sync(wg);
        float *out = a.est + (static_cast<int64_t>(b) * C + cc) * a.est_len;
        if (owner) {
          float4 lo4 = make_float4(y[0], y[1], y[2], y[3]), hi4 = make_float4(y[4], y[5], y[6], y[7]);
          if (m > 0) {
            const float *pu = up + (m - 1) * 8;
            lo4.x += pu[0]; lo4.y += pu[1]; lo4.z += pu[2]; lo4.w += pu[3];
            hi4.x += pu[4]; hi4.y += pu[5]; hi4.z += pu[6]; hi4.w += pu[7];
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
    }
  }
#undef SEP_BAR

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
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  // float4 loads of the frames and float4 stores of the estimates: a contiguous but offset device view
  // (buf[1:]) would fault with a misaligned address inside the kernel, so refuse it here
  SEP_REQUIRE((reinterpret_cast<uintptr_t>(a.wave) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.est) & 15) == 0 &&
                  (a.code == nullptr || (reinterpret_cast<uintptr_t>(a.code) & 15) == 0),
              "sep_filterbank_separate_f32: wave, est and code must be 16-byte aligned");
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
  profile_begin(stream, "filterbank_kernel (tcgen05 kind::tf32 x3, TMEM, TMA; N=256 L=16 stride=8, C=%d)", n_src);
  filterbank_kernel<<<grid, kFbThreads, kFbSmem, stream>>>(a, mask_map);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, est, a.est, n_est, mem))) return rc;
  if ((rc = copy_back(s, code, a.code, n_code, mem))) return rc;
  return finish(s, mem);
}
