// fft.cuh -- shared-memory FFT building blocks (generic power-of-two path).
//
// A real FFT of n = 2M points is one M-point complex FFT of z[m] = x[2m] +
// i x[2m+1] plus a twiddled split (and the mirror image for the inverse), so a
// frame costs half a complex transform.  One warp owns one frame; stages are
// separated by __syncwarp only.  The index math is emulated on the CPU against
// numpy.fft in tools/emulate_fft.py.
#pragma once

#include "common.cuh"

namespace sep {

// M-point complex FFT by one warp, Stockham autosort (no bit reversal), radix-4
// stages and one trailing radix-2 stage when log2(M) is odd.  Ping-pongs
// between `a` (input) and `b`; returns the buffer holding the result.
// tw[k] = exp(-2 pi i k / M).  INV = true conjugates the twiddles
// (unnormalised inverse).  M >= 4.
template <bool INV>
__device__ __forceinline__ float2 *warp_fft(float2 *a, float2 *b, const float2 *__restrict__ tw,
                                            int M, int lane) {
  float2 *in = a, *out = b;
  int Ns = 1;
  const int q4 = M >> 2;
  while (Ns * 4 <= M) {
    const int tstep = M / (Ns * 4);
    for (int j = lane; j < q4; j += 32) {
      const int k = j & (Ns - 1);
      float2 v0 = in[j], v1 = in[j + q4], v2 = in[j + 2 * q4], v3 = in[j + 3 * q4];
      if (Ns > 1) {
        float2 w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        v1 = cmul(v1, w1);
        v2 = cmul(v2, w2);
        v3 = cmul(v3, w3);
      }
      const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
      const float2 d = csub(v1, v3);
      // forward: a3 = -i d = (d.y, -d.x); inverse: a3 = +i d = (-d.y, d.x)
      const float2 a3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
      const int dst = ((j - k) << 2) + k;
      out[dst] = cadd(a0, a2);
      out[dst + Ns] = cadd(a1, a3);
      out[dst + 2 * Ns] = csub(a0, a2);
      out[dst + 3 * Ns] = csub(a1, a3);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
    Ns <<= 2;
  }
  if (Ns < M) {  // Ns * 2 == M
    const int q2 = M >> 1;
    for (int j = lane; j < q2; j += 32) {
      float2 w = tw[j];
      if (INV) w.y = -w.y;
      const float2 v0 = in[j], v1 = cmul(in[j + q2], w);
      out[j] = cadd(v0, v1);
      out[j + Ns] = csub(v0, v1);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
  }
  return in;
}

// Loads one windowed frame as z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1].
// `x` points at the frame start inside a shared wave tile; win_half holds
// 0.5 * analysis window so that the split below needs no extra scaling.
__device__ __forceinline__ void load_frame_packed(float2 *z, const float *x,
                                                  const float *__restrict__ win_half, int M,
                                                  int lane) {
  for (int m = lane; m < M; m += 32) {
    const float2 w = reinterpret_cast<const float2 *>(win_half)[m];
    z[m] = make_float2(x[2 * m] * w.x, x[2 * m + 1] * w.y);
  }
  __syncwarp();
}

// Split of the packed spectrum: X[k] = (Z[k] + conj Z[M-k]) - i e^{-2 pi i k/n} (Z[k] - conj Z[M-k])
// for one k in [0, M]  (Z[M] == Z[0]; the 1/2 lives in the window).
__device__ __forceinline__ float2 real_split(const float2 *Z, const float2 *__restrict__ tw_full,
                                             int M, int k) {
  const float2 zk = Z[k == M ? 0 : k];
  float2 zc = Z[k == 0 ? 0 : M - k];
  zc.y = -zc.y;
  const float2 e = cadd(zk, zc), o = csub(zk, zc);
  const float2 t = cmul(o, tw_full[k]);  // times e^{-2 pi i k / n}
  // -i * t = (t.y, -t.x)
  return make_float2(e.x + t.y, e.y - t.x);
}

// Inverse of real_split for one k in [0, M): Z[k] = (Y[k] + conj Y[M-k]) + i e^{+2 pi i k/n} (Y[k] - conj Y[M-k]).
// The caller must have zeroed the imaginary parts of Y[0] and Y[M] (numpy's
// irfft ignores them).  The 1/n normalisation lives in the synthesis window.
__device__ __forceinline__ float2 real_merge(float2 yk, float2 ymk, float2 tw_k) {
  ymk.y = -ymk.y;
  const float2 e = cadd(yk, ymk), o = csub(yk, ymk);
  tw_k.y = -tw_k.y;  // conj -> e^{+2 pi i k / n}
  const float2 t = cmul(o, tw_k);
  // +i * t = (-t.y, t.x)
  return make_float2(e.x - t.y, e.y + t.x);
}

}  // namespace sep
