// epilogue chunk as in epilogue_chunk.cu (mode mufu), plus per chunk a tcgen05.ld.x16 + tcgen05.wait::ld (LD = 1: issued one
// chunk ahead, as the kernel does; LD = 2: issued and waited back to back), to price the tensor-memory read in the loop
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace sep;
__device__ __forceinline__ float ex2p(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
template <int LD>
__global__ void k(float *out, int iters, long long *cyc, int mma_warp, int work_warps, float *gout) {
  extern __shared__ float sm[];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  __shared__ uint64_t mbar;
  __shared__ uint64_t dummy[2];
  float *nb = sm, *ob = sm + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) nb[i] = 0.01f * i;
  if (threadIdx.x == 0) { done = 0; mbar_init(smem_u32(&mbar), 1); mbar_init(smem_u32(&dummy[0]), 384); mbar_init(smem_u32(&dummy[1]), 96); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // B operand region for the MMA warp: 48 KB behind the output rows
  float *bop = ob + 8 * 32 * 129;
  for (int i = threadIdx.x; i < 12 * 1024; i += blockDim.x) bop[i] = 0.001f * (i & 63);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp >= work_warps + 2) {
    // four staging warps: 10 x LDG.128 of a row half (L2-resident), tf32 split, 20 x tcgen05.st.x8, wait::st, as the kernel's
    if (mma_warp & 4) {
      const int sidx = threadIdx.x - (work_warps + 2) * 32;
      const uint32_t a_addr = tmem + 300 + (static_cast<uint32_t>(sidx & ~31) << 16);
      const float4 *src = reinterpret_cast<const float4 *>(gout) + static_cast<size_t>(blockIdx.x) * (1u << 18);
      uint32_t n = 0;
      while (!done) {
        float4 pre[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) pre[q] = __ldg(src + ((n * 128 + sidx) * 10 + q) % (1u << 18));
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          float hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 v = pre[2 * g + q];
            split_tf32(v.x, hi[4 * q], lo[4 * q]); split_tf32(v.y, hi[4 * q + 1], lo[4 * q + 1]);
            split_tf32(v.z, hi[4 * q + 2], lo[4 * q + 2]); split_tf32(v.w, hi[4 * q + 3], lo[4 * q + 3]);
          }
          tmem_st8(a_addr + 8 * g, hi);
          tmem_st8(a_addr + 80 + 8 * g, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        ++n;
      }
      if (blockIdx.x == 0 && sidx == 0) cyc[3] = n;
    }
  } else if (warp == work_warps + 1) {
    // a store warp: bulk stores of 32-row quarters of the output image, as the kernel's store warp issues them
    if ((mma_warp & 2) && lane == 0) {
      uint32_t n = 0;
      char *gdst = reinterpret_cast<char *>(gout) + static_cast<size_t>(blockIdx.x) * (4u << 20);
      while (!done) {
        for (int q = 0; q < 4; ++q)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst + ((n & 31) * 4 + q) * 16512), "r"(smem_u32(ob) + q * 16512), "r"(16512u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        ++n;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (blockIdx.x == 0) cyc[2] = n;
    }
  } else if (warp == work_warps) {
    // the MMA warp: tcgen05.mma kind::tf32 M128 N144 K8, A in tensor memory, B in shared memory, until the workers are done
    if (mma_warp & 1) {
      const uint32_t idesc = umma_idesc_tf32(128, 144);
      const uint64_t bdesc = umma_desc(smem_u32(bop), 18 * 128, 128);
      uint32_t n = 0, phase = 0;
      while (!done) {
        for (int i = 0; i < 30; ++i) umma_tf32_ts_elect(tmem + 256, tmem + 420, bdesc, idesc, i > 0);
        umma_commit_elect(smem_u32(&mbar));
        mbar_wait(smem_u32(&mbar), phase);
        phase ^= 1;
        ++n;
      }
      if (lane == 0 && blockIdx.x == 0) cyc[1] = n;
    }
  } else {
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float *orow = ob + ((warp & 7) * 32 + lane) * 129;
  uint32_t r0[16], r1[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) r0[e] = r1[e] = __float_as_uint(0.001f * (lane + e));
  if (LD == 1) { ld16_issue(lane_addr, r0); ld16_wait(r0); }
  auto chunk = [&](uint32_t (&r)[16], int c0) {
    const float4 *nb4 = reinterpret_cast<const float4 *>(nb + c0);
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 b = nb4[q];
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[4 * q + e] = fmaf(__uint_as_float(r[4 * q + e]) * 1e-30f, -1.44269504f, bb[e]);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = ex2p(v[e]);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 1.f + v[e];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = rcpp(v[e]);
#pragma unroll
    for (int e = 0; e < 16; ++e) orow[c0 + e] = v[e];
  };
  long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    if ((mma_warp & 16) && (it % 6) == 0)                       // all epilogue warps start their tiles together
      asm volatile("bar.sync 1, %0;" :: "r"(work_warps * 32) : "memory");
    if ((mma_warp & 8) && (it % 3) == 0) {                      // the kernel's per-tile fences and arrives
      tc_fence_before();
      mbar_arrive(smem_u32(&dummy[0]));
      fence_async_smem();
      mbar_arrive(smem_u32(&dummy[1]));
      tc_fence_after();
    }
    if (LD == 1) ld16_issue(lane_addr + 16, r1);
    if (LD == 2) { ld16_issue(lane_addr, r0); ld16_wait(r0); }
    chunk(r0, 0);
    if (LD == 1) { ld16_wait(r1); ld16_issue(lane_addr + 32, r0); }
    if (LD == 2) { ld16_issue(lane_addr + 16, r1); ld16_wait(r1); }
    chunk(r1, 16);
    if (LD == 1) ld16_wait(r0);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(r0[0]) + orow[3];
  __syncwarp();
  if (warp == 0 && lane == 0) done = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u)); }
}
int main() {
  float *out, *gout; long long *cyc, h[4];
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 32); cudaMalloc(&gout, 148ull * (4u << 20));
  const int iters = 2000;
  const size_t smem = (256 + 8 * 32 * 129) * 4 + 48 * 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mma = 7; mma < 32; mma += 16)
  for (int ld = 1; ld < 2; ++ld)
    for (int warps = 8; warps <= 12; warps += 4) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(cyc, 0, 32);
        if (ld == 0) k<0><<<148, (warps + 6) * 32, smem>>>(out, iters, cyc, mma, warps, gout);
        if (ld == 1) k<1><<<148, (warps + 6) * 32, smem>>>(out, iters, cyc, mma, warps, gout);
        cudaError_t e = cudaDeviceSynchronize(); if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
      printf("%s, %s, %2d warps per SM: %.0f cycles per chunk and warp = %.0f per chunk and scheduler (MMA batches of 30: %lld, %.0f cycles per MMA; 66 KB tile stores: %lld; staging halves: %lld)\n",
             mma == 7 ? "MMA + store + staging warps" : "MMA + store + staging warps + a barrier of all epilogue warps every 3 chunks", ld == 0 ? "no tcgen05.ld" : "tcgen05.ld one chunk ahead", warps,
             (double)h[0] / iters, (double)h[0] / iters / (warps / 4.0), h[1], h[1] ? (double)h[0] / (30.0 * h[1]) : 0.0, h[2], h[3]);
    }
  return 0;
}
