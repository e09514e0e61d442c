// tcgen05.mma kind::tf32 throughput: M = 128, K = 8 per instruction, N varied, A from tensor memory or shared memory,
// one or two accumulators alternating; optional concurrent tcgen05.ld traffic from 4 other warps
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace sep;
__global__ void __launch_bounds__(256, 1) k(int N, int from_tmem, int n_acc, int n_mma, int ld_traffic, long long *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(sm)[i] = 0.001f * (i & 255);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 4); done = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp >= 4) {
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint32_t sm0 = smem_u32(sm);
    const uint32_t lbo_a = 16 * 128, lbo_b = (N / 8) * 128;
    const uint64_t adesc = umma_desc(sm0, lbo_a, 128), bdesc = umma_desc(sm0 + 8192, lbo_b, 128);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + (warp - 4) * 64 + (n_acc == 2 ? (i & 1) * 256 : 0);
      if (from_tmem) umma_tf32_ts_elect(d, tmem + 496, bdesc, idesc, i >= n_acc ? 1u : 0u);
      else umma_tf32_elect(d, adesc, bdesc, idesc, i >= n_acc ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (lane == 0) { done = 1; if (blockIdx.x == 0 && warp == 4) { out[0] = t1 - t0; out[1] = t2 - t0; } }
  } else if (ld_traffic && warp < 4) {
    // 4 warps reading accumulator columns in a loop (an epilogue's tcgen05.ld traffic)
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float d[16], s = 0.f;
    long long cnt = 0;
    while (!done) {
      tmem_ld16(lane_addr + 256 + 16 * (cnt & 7), d);
      s += d[0];
      ++cnt;
    }
    if (s == 123.f) out[5] = 1;
    if (blockIdx.x == 0 && lane == 0 && warp == 0) out[2] = cnt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u)); }
}
int main() {
  long long *out, h[3];
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 512;
  for (int ld = 0; ld < 2; ++ld)
    for (int from_tmem = 1; from_tmem >= 0; --from_tmem)
      for (int n_acc = 1; n_acc <= 2; ++n_acc) {
        const int Ns[] = {16, 32, 64};
        for (int N : Ns) {
          cudaMemset(out, 0, 64);
          for (int rep = 0; rep < 2; ++rep) k<<<148, 256, 64 * 1024>>>(N, from_tmem, n_acc, n_mma, ld, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e == cudaSuccess) e = cudaGetLastError();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
          printf("A from %s, %d accumulator(s), N=%3d, ld traffic %d: 4 issuers: issue %.1f clk/MMA, complete %.1f clk/MMA (floor N/2 = %d), ld16 per warp %lld\n",
                 from_tmem ? "tmem" : "smem", n_acc, N, ld, (double)h[0] / n_mma / 4, (double)h[1] / n_mma / 4, N / 2, h[2]);
        }
      }
  return 0;
}
