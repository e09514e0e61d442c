#include <cstdio>
#include <cuda_runtime.h>
// per-SMSP throughput of MUFU.EX2 / MUFU.RCP / FFMA chains with W warps per SM
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); v[i] += 1.f; asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i])); }
      if (MODE == 3) v[i] = fmaf(v[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double ops = (mode == 2 ? 2.0 : 1.0) * 16.0 * iters * (warps / 4.0);   // warp-instructions of that kind per SMSP
      printf("mode %d (%s) warps/SM %2d: %lld clk, %.2f clk per warp-instr per SMSP\n", mode,
             mode == 0 ? "ex2" : mode == 1 ? "rcp" : mode == 2 ? "ex2+add+rcp" : "ffma", warps, h, h / ops);
    }
  return 0;
}
