// fft256w.cuh -- 256-point complex FFT on a whole warp, 8 points per lane: half the registers per lane
// of the half-warp transform in fft256.cuh, so twice as many warps fit on an SM.
//
// 256 = 8 x 4 x 8, natural order in AND out: lane q holds element q + 32 j in register j, before and
// after.  With n = q0 + 8 q1 + 32 m (lane q = q0 + 8 q1) and k = k0 + 8 k1 + 32 k2 (lane l = k0 + 8 k1):
//   1   8-point FFT in registers over m                       -> k0
//       twiddle W256^(q k0)
//   X1  exchange through shared memory: lane (k0, h) = k0 + 8 h gets the values of q0 = 2h, 2h + 1, all q1
//   2   two 4-point FFTs over q1                               -> k1
//       twiddle W32^(q0 k1)
//   X2  exchange: lane l = k0 + 8 k1 gets q0 = 0..7
//   3   8-point FFT over q0                                    -> k2
// Input and output layouts are the same, so the inverse is the same routine with conjugated twiddles
// (unnormalised).  Exchange rows: X1 pitch 34 float2 (STS.64 by 32 consecutive lanes, LDS.128 with the
// 8 lanes of a quarter-warp 16 bytes apart mod 128), X2 pitch 10 float2 (STS.64 with q0 = 2h + e in column
// h + 4e, LDS.128 with quarter-warp lanes 80 bytes apart: all 8 bank groups) -- conflict-free both ways.
// Because lane q holds bins q + 32 j, the conjugate partner Z[256 - k] of register j lives on lane
// (32 - q) % 32 in register 7 - j (lane 0: its own register (8 - j) % 8): one shuffle per bin separates
// the half spectra of two real frames riding in one complex transform, and mask rows are read as
// 128-byte runs.  The index algebra is emulated on the CPU in tools/emulate_wfft.py.
#pragma once

#include "common.cuh"
#include "fft256.cuh"

namespace sep {

constexpr int kWxA = 34;                          // exchange 1 row pitch, float2
constexpr int kWxB = 10;                          // exchange 2 row pitch, float2
constexpr int kWxOne = 8 * kWxA;                   // exchange 1 region (272 float2)
constexpr int kWxFloat2 = kWxOne + 32 * kWxB;     // exchange buffer of one warp: X1 rows, then X2 rows
constexpr int kWTw1 = 10;                         // tw1 row pitch: tw1[q * 10 + k0] = W256^(q k0)
                                                  // tw2[h * 8 + 2 k1 + e] = W32^((2h + e) k1)

// In-register 8-point FFT, natural order in and out.
template <bool INV>
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  float2 a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = cadd(v[j], v[j + 4]); b[j] = csub(v[j], v[j + 4]); }
  b[1] = tw16<INV, 2>(b[1]);                       // W8^1
  b[3] = tw16<INV, 6>(b[3]);                       // W8^3
  // W8^2 = -+i on b[2] is folded into the 4-point butterfly below
  // a: X0 X2 X4 X6
  fft4<INV>(a[0], a[1], a[2], a[3]);
  // b with b2' = -+i b2:  t0 = b0 + b2', t1 = b0 - b2', t2 = b1 + b3, t3 = b1 - b3
  const float2 t0 = INV ? cadd_pi(b[0], b[2]) : cadd_mi(b[0], b[2]);
  const float2 t1 = INV ? cadd_mi(b[0], b[2]) : cadd_pi(b[0], b[2]);
  const float2 t2 = cadd(b[1], b[3]), t3 = csub(b[1], b[3]);
  v[0] = a[0]; v[2] = a[1]; v[4] = a[2]; v[6] = a[3];
  v[1] = cadd(t0, t2);
  v[5] = csub(t0, t2);
  v[3] = INV ? cadd_pi(t1, t3) : cadd_mi(t1, t3);
  v[7] = INV ? cadd_mi(t1, t3) : cadd_pi(t1, t3);
}

// The three phases of the transform, separated by the two exchanges (warp barriers between them).
// t1: this lane's row of tw1 (8 float2, 16-byte aligned); t2: the row of tw2 of this lane's h = q >> 3;
// ex: an exchange buffer of kWxFloat2 float2 (X1 rows, then X2 rows).  Two exchange regions: the reads of
// one are ordered before the next writes of the same region by the barrier of the OTHER exchange, so a
// transform needs two warp barriers, not four.
// twiddle rows of a lane: tw1[1..7] = W256^(q k0), tw2[2..7] = W32^(q0 k1) (entries 0 / 0, 1 are ones)
struct WTwiddles {
  float2 a[8], b[8];
};
__device__ __forceinline__ void wfft_load_tw1(WTwiddles &t, const float2 *t1) {
  const float4 *t14 = reinterpret_cast<const float4 *>(t1);
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const float4 w = t14[k / 2];
    t.a[k] = make_float2(w.x, w.y);
    t.a[k + 1] = make_float2(w.z, w.w);
  }
}
__device__ __forceinline__ void wfft_load_tw2(WTwiddles &t, const float2 *t2) {
  const float4 *t24 = reinterpret_cast<const float4 *>(t2);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    const float4 w = t24[i / 2];
    t.b[i] = make_float2(w.x, w.y);
    t.b[i + 1] = make_float2(w.z, w.w);
  }
}

template <bool INV>
__device__ __forceinline__ void wfft_phase1(float2 (&v)[8], const WTwiddles &t, float2 *ex, int q) {
  fft8<INV>(v);
#pragma unroll
  for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], make_float2(t.a[k].x, INV ? -t.a[k].y : t.a[k].y));
#pragma unroll
  for (int k = 0; k < 8; ++k) ex[k * kWxA + q] = v[k];
}

template <bool INV>
__device__ __forceinline__ void wfft_phase2(float2 (&v)[8], const WTwiddles &t, float2 *ex, int q) {
  const int k0 = q & 7, h = q >> 3;
  const float4 *ra = reinterpret_cast<const float4 *>(ex + k0 * kWxA + 2 * h);
#pragma unroll
  for (int q1 = 0; q1 < 4; ++q1) {
    const float4 r = ra[4 * q1];
    v[2 * q1] = make_float2(r.x, r.y);
    v[2 * q1 + 1] = make_float2(r.z, r.w);
  }
  fft4<INV>(v[0], v[2], v[4], v[6]);
  fft4<INV>(v[1], v[3], v[5], v[7]);
#pragma unroll
  for (int i = 2; i < 8; ++i) v[i] = cmul(v[i], make_float2(t.b[i].x, INV ? -t.b[i].y : t.b[i].y));
  // X2: row k0 + 8 k1, column h + 4 e holds q0 = 2 h + e (64-bit stores: no register quads to assemble; the 16
  // lanes of a half-warp -- k0 = 0..7, two values of h -- hit 16 distinct 8-byte banks)
  float2 *wb = ex + kWxOne + k0 * kWxB + h;
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    wb[8 * kWxB * k1] = v[2 * k1];
    wb[8 * kWxB * k1 + 4] = v[2 * k1 + 1];
  }
}

template <bool INV>
__device__ __forceinline__ void wfft_phase3(float2 (&v)[8], const float2 *ex, int q) {
  const float4 *rb = reinterpret_cast<const float4 *>(ex + kWxOne + q * kWxB);
  const float4 a = rb[0], b = rb[1], c = rb[2], d = rb[3];     // columns (0,1) (2,3) (4,5) (6,7) = q0 (0,2) (4,6) (1,3) (5,7)
  v[0] = make_float2(a.x, a.y); v[2] = make_float2(a.z, a.w);
  v[4] = make_float2(b.x, b.y); v[6] = make_float2(b.z, b.w);
  v[1] = make_float2(c.x, c.y); v[3] = make_float2(c.z, c.w);
  v[5] = make_float2(d.x, d.y); v[7] = make_float2(d.z, d.w);
  fft8<INV>(v);
}

// v[j] = x[q + 32 j] -> v[j] = X[q + 32 j].
template <bool INV>
__device__ __forceinline__ void wfft256(float2 (&v)[8], const float2 *t1, const float2 *t2, float2 *ex, int q) {
  WTwiddles t;
  wfft_load_tw1(t, t1);
  wfft_phase1<INV>(v, t, ex, q);
  wfft_load_tw2(t, t2);
  __syncwarp();
  wfft_phase2<INV>(v, t, ex, q);
  __syncwarp();
  wfft_phase3<INV>(v, ex, q);
}

// Two independent transforms in lockstep (exchange buffers ex and ex + kWxFloat2): twice the independent
// instructions between the same two barriers, so the shared-memory round trips of one hide behind the
// butterflies of the other; the twiddle rows are read once for both.
template <bool INV>
__device__ __forceinline__ void wfft256x2(float2 (&va)[8], float2 (&vb)[8], const float2 *t1, const float2 *t2,
                                          float2 *ex, int q) {
  WTwiddles t;
  wfft_load_tw1(t, t1);
  wfft_phase1<INV>(va, t, ex, q);
  wfft_phase1<INV>(vb, t, ex + kWxFloat2, q);
  wfft_load_tw2(t, t2);
  __syncwarp();
  wfft_phase2<INV>(va, t, ex, q);
  wfft_phase2<INV>(vb, t, ex + kWxFloat2, q);
  __syncwarp();
  wfft_phase3<INV>(va, ex, q);
  wfft_phase3<INV>(vb, ex + kWxFloat2, q);
}

// After a forward transform of z = a + i b (a, b real, the 1/2 of the split folded into the window):
// XR = (Re A, Re B), XI = (Im A, Im B) at this lane's bin q + 32 r, r = 0..4 (r = 4, bin 128, is
// meaningful on lane 0 only).
__device__ __forceinline__ void wsplit_planar(const float2 (&v)[8], int q, int r, float2 &XR, float2 &XI) {
  const int src = (32 - q) & 31;
  float2 zp = v[4];                      // r == 4: bin 128 is its own partner (and lives on lane 0 only)
  if (r < 4) {
    float2 got;
    got.x = __shfl_sync(0xffffffffu, v[7 - r].x, src);
    got.y = __shfl_sync(0xffffffffu, v[7 - r].y, src);
    const float2 own = v[(8 - r) & 7];
    zp = (q == 0) ? own : got;
  }
  const float2 z = v[r];
  XR = __fadd2_rn(z, zp);                                                  // (zx + z'x, zy + z'y)
  XI = __fadd2_rn(make_float2(z.y, -z.x), make_float2(-zp.y, zp.x));       // (zy - z'y, z'x - zx)
}

// Input of the inverse transform whose real part is the time frame of spectrum P and whose imaginary
// part is the time frame of spectrum Q, given at this lane's bins q + 32 r (r = 0..4) as
//   L[r] = P[k] + i Q[k],   Mi[r] = conj(P[k]) + i conj(Q[k])   (the value at the mirror bin 256 - k).
__device__ __forceinline__ void wmerge_pair(float2 (&v)[8], int q, const float2 (&L)[5], const float2 (&Mi)[5]) {
  const int src = (32 - q) & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) v[r] = L[r];
#pragma unroll
  for (int j = 4; j < 8; ++j) {
    float2 got;
    got.x = __shfl_sync(0xffffffffu, Mi[7 - j].x, src);
    got.y = __shfl_sync(0xffffffffu, Mi[7 - j].y, src);
    const float2 own = (j == 4) ? L[4] : Mi[8 - j];
    v[j] = (q == 0) ? own : got;
  }
}

}  // namespace sep
