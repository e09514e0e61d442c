// fft256.cuh -- register-resident 256-point complex FFT on a half-warp.
//
// 256 = 16 x 16.  Sixteen lanes own one transform; lane p holds x[p + 16 m],
// m = 0..15, in registers.  Pass 1 is a 16-point FFT over m inside each lane,
// then a twiddle by W256^(p k1), ONE transpose through shared memory (pitch 17
// float2: conflict-free both ways), and pass 2 is a 16-point FFT over p.  The
// result stays in registers with lane k1 holding X[k1 + 16 k2] in v[k2].
// The inverse is the mirror image (same per-lane twiddles, conjugated): input
// v[k2] = X[lane + 16 k2], output v[m] = x[lane + 16 m].
//
// Two real sequences ride in one complex transform (re = first, im = second);
// `split_pair` separates their half spectra with one shuffle per bin and
// `merge_pair` builds the Hermitian-extended input of the inverse transform.
// The whole index algebra is emulated on the CPU in tools/emulate_fft.py.
#pragma once

#include "common.cuh"

namespace sep {

constexpr int kXchPitch = 17;                    // float2 per transpose row
constexpr int kXchFloat2 = 16 * kXchPitch;       // per half-warp

template <bool INV>
__device__ __forceinline__ void fft4(float2 &x0, float2 &x1, float2 &x2, float2 &x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
  // forward: X1 = t1 - i t3, X3 = t1 + i t3 ; inverse: swapped.  8 packed adds in all.
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = INV ? cadd_pi(t1, t3) : cadd_mi(t1, t3);
  x3 = INV ? cadd_mi(t1, t3) : cadd_pi(t1, t3);
}

// v *= W16^e (forward) or its conjugate (inverse), e in {1, 2, 3, 4, 6, 9}
template <bool INV, int E>
__device__ __forceinline__ float2 tw16(float2 v) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
  // W16^e = (wr, wi) forward
  constexpr float wr = (E == 1) ? c1 : (E == 2) ? r : (E == 3) ? s1 : (E == 4) ? 0.f : (E == 6) ? -r : -c1;
  constexpr float wi0 = (E == 1) ? -s1 : (E == 2) ? -r : (E == 3) ? -c1 : (E == 4) ? -1.f : (E == 6) ? -r : s1;
  constexpr float wi = INV ? -wi0 : wi0;
  if (E == 4) return INV ? cadd_pi(make_float2(0.f, 0.f), v) : cadd_mi(make_float2(0.f, 0.f), v);
  return cmul(v, make_float2(wr, wi));
}

// In-register 16-point FFT, natural order in and out (n = 4a + b, k = c + 4d).
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) fft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);   // v[4c + b] = y[b][c]
  v[4 * 1 + 1] = tw16<INV, 1>(v[4 * 1 + 1]);
  v[4 * 1 + 2] = tw16<INV, 2>(v[4 * 1 + 2]);
  v[4 * 1 + 3] = tw16<INV, 3>(v[4 * 1 + 3]);
  v[4 * 2 + 1] = tw16<INV, 2>(v[4 * 2 + 1]);
  v[4 * 2 + 2] = tw16<INV, 4>(v[4 * 2 + 2]);
  v[4 * 2 + 3] = tw16<INV, 6>(v[4 * 2 + 3]);
  v[4 * 3 + 1] = tw16<INV, 3>(v[4 * 3 + 1]);
  v[4 * 3 + 2] = tw16<INV, 6>(v[4 * 3 + 2]);
  v[4 * 3 + 3] = tw16<INV, 9>(v[4 * 3 + 3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) fft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // v[4c + d] = X[c + 4d]
  float2 o[16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = 0; d < 4; ++d) o[c + 4 * d] = v[4 * c + d];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// 256-point complex FFT on 16 lanes.  tw points at this lane's column of the shared
// twiddle table: tw[16 * j] = exp(-2 pi i * l16 * j / 256) (lanes read consecutive
// float2: conflict-free, and both half-warps read the same words: broadcast).
// xch: this half-warp's private transpose buffer (kXchFloat2 float2).
template <bool INV>
__device__ __forceinline__ void fft256(float2 (&v)[16], const float2 *tw, float2 *xch, int l16) {
  fft16<INV>(v);
#pragma unroll
  for (int j = 1; j < 16; ++j) {
    float2 w = tw[16 * j];
    if (INV) w.y = -w.y;
    v[j] = cmul(v[j], w);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) xch[j * kXchPitch + l16] = v[j];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = xch[l16 * kXchPitch + q];
  fft16<INV>(v);
}

// After a forward transform of z = a + i b (a, b real): half spectra of a and b
// at this lane's bins l16 + 16 r, r = 0..8 (r = 8, bin 128, is meaningful on
// lane 0 only).  The 1/2 of the split is expected to be folded into the window.
__device__ __forceinline__ void split_pair(const float2 (&v)[16], int l16, float2 (&A)[9],
                                           float2 (&B)[9]) {
  const int src = (16 - l16) & 15;
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    // partner Z[256 - k]: lane (16 - l16) % 16, register 15 - r; on lane 0 it is
    // this lane's own register (16 - r) % 16
    float2 got;
    got.x = __shfl_sync(0xffffffffu, v[r < 8 ? 15 - r : 7].x, src, 16);
    got.y = __shfl_sync(0xffffffffu, v[r < 8 ? 15 - r : 7].y, src, 16);
    const float2 own = v[(16 - r) & 15];
    const float2 zp = (l16 == 0) ? own : got;
    const float2 z = v[r];
    A[r] = cadd_conj(z, zp);                                   // Z + conj Z'
    const float2 d = __fadd2_rn(z, make_float2(-zp.x, zp.y));  // Z - conj Z'
    B[r] = make_float2(d.y, -d.x);                             // -i (Z - conj Z')
  }
}

// Input of the inverse transform whose real part is the time frame of spectrum
// P and whose imaginary part is the time frame of spectrum Q, both given at
// this lane's bins l16 + 16 r (r = 0..8) as local products:
//   L[r] = P[k] + i Q[k]            (k <= 128)
//   Mi[r] = conj(P[k]) + i conj(Q[k])   (value needed at the mirror bin 256 - k)
__device__ __forceinline__ void merge_pair(float2 (&v)[16], int l16, const float2 (&L)[9],
                                           const float2 (&Mi)[9]) {
  const int src = (16 - l16) & 15;
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = L[r];
#pragma unroll
  for (int k2 = 8; k2 < 16; ++k2) {
    // bin l16 + 16 k2 >= 128: mirror of the bin held by lane `src`, register 15 - k2
    float2 got;
    got.x = __shfl_sync(0xffffffffu, Mi[15 - k2].x, src, 16);
    got.y = __shfl_sync(0xffffffffu, Mi[15 - k2].y, src, 16);
    // lane 0: bin 16 k2 mirrors its own bin 16 (16 - k2); k2 == 8 is bin 128 itself
    const float2 own = (k2 == 8) ? L[8] : Mi[(16 - k2) & 7];
    v[k2] = (l16 == 0) ? own : got;
  }
}


// ---- variant with 128-bit shared-memory reads (fused_strip.cu) ----
// Transpose rows and the per-lane twiddle rows have a pitch of 18 float2: a lane's 16
// values are contiguous and 16-byte aligned (8 LDS.128 instead of 16 LDS.64), and the 8
// lanes of a quarter-warp land on 8 distinct 16-byte bank groups (36 l mod 32 = 4 l).
constexpr int kXchPitchV = 18;
constexpr int kXchFloat2V = 16 * kXchPitchV;     // per half-warp

// twl: this lane's twiddle row, twl[j] = exp(-2 pi i * l16 * j / 256), j = 0..15.
template <bool INV>
__device__ __forceinline__ void fft256v(float2 (&v)[16], const float2 *twl, float2 *xch, int l16) {
  fft16<INV>(v);
  const float4 *tw4 = reinterpret_cast<const float4 *>(twl);
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float4 w = tw4[j / 2];
    if (j > 0) v[j] = cmul(v[j], make_float2(w.x, INV ? -w.y : w.y));
    v[j + 1] = cmul(v[j + 1], make_float2(w.z, INV ? -w.w : w.w));
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) xch[j * kXchPitchV + l16] = v[j];
  __syncwarp();
  const float4 *row = reinterpret_cast<const float4 *>(xch + l16 * kXchPitchV);
#pragma unroll
  for (int q = 0; q < 16; q += 2) {
    const float4 t = row[q / 2];
    v[q] = make_float2(t.x, t.y);
    v[q + 1] = make_float2(t.z, t.w);
  }
  fft16<INV>(v);
}

// Planar split: after a forward transform of z = a + i b, XR[r] = (Re A, Re B) and
// XI[r] = (Im A, Im B) at this lane's bins l16 + 16 r -- the same two packed adds as
// split_pair, but frames a and b side by side so that everything downstream (|X|, labels,
// squared differences, mask multiply) runs on packed pairs.
__device__ __forceinline__ void split_planar(const float2 (&v)[16], int l16, int r, float2 &XR, float2 &XI) {
  const int src = (16 - l16) & 15;
  float2 got;
  got.x = __shfl_sync(0xffffffffu, v[r < 8 ? 15 - r : 7].x, src, 16);
  got.y = __shfl_sync(0xffffffffu, v[r < 8 ? 15 - r : 7].y, src, 16);
  const float2 own = v[(16 - r) & 15];
  const float2 zp = (l16 == 0) ? own : got;
  const float2 z = v[r];
  XR = __fadd2_rn(z, zp);                                                  // (zx + z'x, zy + z'y)
  XI = __fadd2_rn(make_float2(z.y, -z.x), make_float2(-zp.y, zp.x));       // (zy - z'y, z'x - zx)
}

}  // namespace sep
