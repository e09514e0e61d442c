#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
// epilogue-like chunk: 4 LDS.128 bias, 16 x (FFMA, EX2, FADD, RCP, STS.32), per warp; W warps per SM; optional spinner warps
__device__ __forceinline__ float ex2p(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc, int work_warps) {
  extern __shared__ float sm[];
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  float *nb = sm, *ob = sm + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(1)); stop = 0; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) nb[i] = 0.01f * i;
  __syncthreads();
  if (warp >= work_warps) {   // spinner warps: wait on a barrier that never completes until stop
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    while (!stop) { mbar_try(b, 0); }
    return;
  }
  float *orow = ob + ((warp & 7) * 32 + lane) * 129;
  uint32_t r[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(0.001f * (lane + e));
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int c0 = (it & 3) * 16;
    const float4 *nb4 = reinterpret_cast<const float4 *>(nb + c0);
    float bb[16], v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) { float4 b = nb4[q]; bb[4*q]=b.x; bb[4*q+1]=b.y; bb[4*q+2]=b.z; bb[4*q+3]=b.w; }
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaf(__uint_as_float(r[e]), -1.44269504f, bb[e]);
    if (MODE >= 1) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = ex2p(v[e]);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 1.f + v[e];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = rcpp(v[e]);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) orow[c0 + e] = v[e];
#pragma unroll
    for (int e = 0; e < 16; ++e) r[e] ^= __float_as_uint(v[e]) & 0xff;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { if (blockIdx.x == 0) *cyc = t1 - t0; }
  __syncwarp();
  // let spinners go once all work warps are done (approx: warp 0 signals)
  if (warp == 0 && lane == 0) stop = 1;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(r[0]);
}
int main() {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const size_t smem = (256 + 8 * 32 * 129) * 4;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode)
    for (int work = 4; work <= 16; work *= 2)
      for (int spin = 0; spin <= 8; spin += 4) {
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) k<0><<<148, (work + spin) * 32, smem>>>(out, iters, cyc, work);
          else k<1><<<148, (work + spin) * 32, smem>>>(out, iters, cyc, work);
          cudaError_t e = cudaDeviceSynchronize(); if (e == cudaSuccess) e = cudaGetLastError();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%s work warps/SM %2d spinners %d: %.0f clk per chunk per warp\n", mode ? "mufu" : "nomufu", work, spin, (double)h / iters);
      }
  return 0;
}
