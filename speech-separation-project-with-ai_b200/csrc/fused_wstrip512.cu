// fused_wstrip512.cu -- the fused hot path for size = 512, shift = 128 (BASELINE config 4: 3-speaker uPIT,
// 6 permutations) on the whole-warp transform of fft256w.cuh: ONE frame per WARP and iteration.
//
// The 512-point REAL transform of a frame is one 256-point COMPLEX transform of z[n] = x[2n] + i x[2n + 1]
// (lane q holds z[q + 32 j] = the sample pairs (2q + 64 j, 2q + 64 j + 1): 64-bit shared / global accesses,
// lane-contiguous) followed by the real-transform butterfly on bin pairs (k, 256 - k):
//   forward   E = Z[k] + conj Z[256-k],  F = Z[k] - conj Z[256-k],  T = i W512^k F      (the 1/2 rides in the window)
//             X[k] = E - T,   X[256-k] = conj(E + T)
//   inverse   Ey = Y[k] + conj Y[256-k], Fy = Y[k] - conj Y[256-k], U = i conj(W512^k) Fy
//             Z'[k] = Ey + U, Z'[256-k] = conj(Ey - U);   y[2n] + i y[2n+1] = IDFT256(Z')  (1/size in the synthesis window)
// Lane q owns the bin pairs (kA, kB) = (q + 32 r, 256 - q - 32 r), r = 0..3 -- its own Z[kA] and ONE shuffled partner
// Z[256 - kA] (lane (32 - q) % 32, register 7 - r) give both bins, so a transform needs 8 shuffles, not 16 -- and lane 0
// additionally bin 128 (slot 4).  Spectra are kept PLANAR per pair, (Re X[kA], Re X[kB]) / (Im ...), so |X|, the PSA
// labels and the squared differences run as packed FP32 on both bins at once, exactly as the 256-point kernel
// (fused_wstrip.cu) does on its frame pairs (the mixture also keeps the complex form X[kA], conj X[kB], which is what
// the inverse butterflies want: every butterfly is six packed instructions, no scalar arithmetic); mask rows (257 floats, 4-byte aligned) are read straight
// into registers as two lane-contiguous 128-byte runs per source (ascending kA, descending kB): no shared-memory
// staging and none of the 4-byte cp.async traffic of the half-warp kernel (fused_strip512.cu).
//
// What else differs from the half-warp kernel: transforms run in LOCKSTEP pairs (mixture + reference 0, references
// 1 + 2, estimates 0 + 1; wfft256x2); the waveforms live in a 5-slot ring of hop blocks per signal, so every sample
// is staged ONCE (one 16-byte cp.async per lane, signal and frame) instead of once per frame that covers it; the
// overlap-add (4 frames per hop block) is a 6-register-pair carry per source in the lane that owns the samples -- no
// shared memory, no hand-over between half-warps; 8 points per lane leave room for 2 CTAs x 4 warps per SM.
// Strip planning, partial rows and the in-kernel finalisation are those of the other strip kernels.
// Reference lines: see fused.cu.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fft256w.cuh"
#include "fused.cuh"
#include "strip_common.cuh"

namespace sep {

constexpr int kW5Pitch = 10;                       // window / twiddle rows [32 lanes][10] float2: conflict-free LDS.128
constexpr int kW5U = 5;                            // i W512^kA rows [32 lanes][5] float2: conflict-free LDS.64

template <int C, bool SCORE, int W>
struct WStrip512Geom {
  static constexpr int SHIFT = 128, H = 3, RING = 5;
  static constexpr int NSIG = SCORE ? 1 + C : 1;
  static constexpr int STAGE_FLOATS = NSIG * RING * SHIFT;
  static constexpr int NACC = SCORE ? C * C + 2 * C : 0;
  static constexpr int NPIT = SCORE ? C * C : 0;
  static constexpr int NEX = C >= 2 || SCORE ? 2 : 1;
  static constexpr int WARP_BYTES = 4 * STAGE_FLOATS + 8 * kWxFloat2 * NEX + 8 * 32 * NACC + 8 * 32 * NPIT;
  static constexpr int TABLE_BYTES = 8 * (3 * 32 * kW5Pitch + 32 + 32 * kW5U);
  static constexpr size_t smem() { return TABLE_BYTES + static_cast<size_t>(W) * WARP_BYTES; }
  static constexpr int NV = FusedVals<C>::NV;
};

// Forward real-transform butterfly of slot r: this lane's Z[kA] (register r) and the partner Z[256 - kA].
// U = i W512^kA.  Everything is packed FP32 with operand swizzles / per-half negation (free on sm_100):
//   E = A + conj P, F = A - conj P, T = U F  ->  planar  XR = (Re X[kA], Re X[kB]) = (E.x - T.x, E.x + T.x),
//   XI = (Im X[kA], Im X[kB]) = (E.y - T.y, -E.y - T.y);  complex  XA = X[kA] = E - T,  XBc = conj X[kB] = E + T.
__device__ __forceinline__ void w5_partner(const float2 (&v)[8], int q, int r, float2 &A, float2 &P) {
  P = v[4];                                        // slot 4: bin 128 is its own partner (lane 0 only)
  if (r < 4) {
    const int src = (32 - q) & 31;
    float2 got;
    got.x = __shfl_sync(0xffffffffu, v[7 - r].x, src);
    got.y = __shfl_sync(0xffffffffu, v[7 - r].y, src);
    const float2 own = v[(8 - r) & 7];
    P = (q == 0) ? own : got;
  }
  A = v[r < 4 ? r : 4];
}
__device__ __forceinline__ void w5_split(const float2 (&v)[8], int q, int r, float2 U, float2 &XR, float2 &XI) {
  float2 A, P;
  w5_partner(v, q, r, A, P);
  const float2 E = __fadd2_rn(A, make_float2(P.x, -P.y));
  const float2 F = __fadd2_rn(A, make_float2(-P.x, P.y));
  const float2 Tt = cmul(F, U);
  XR = __fadd2_rn(make_float2(E.x, E.x), make_float2(-Tt.x, Tt.x));
  XI = __fadd2_rn(make_float2(E.y, -E.y), make_float2(-Tt.y, -Tt.y));
}
// the mixture also keeps the complex form for the inverse butterflies
__device__ __forceinline__ void w5_split_both(const float2 (&v)[8], int q, int r, float2 U, float2 &XR, float2 &XI,
                                              float2 &XA, float2 &XBc) {
  float2 A, P;
  w5_partner(v, q, r, A, P);
  const float2 E = __fadd2_rn(A, make_float2(P.x, -P.y));
  const float2 F = __fadd2_rn(A, make_float2(-P.x, P.y));
  const float2 Tt = cmul(F, U);
  XR = __fadd2_rn(make_float2(E.x, E.x), make_float2(-Tt.x, Tt.x));
  XI = __fadd2_rn(make_float2(E.y, -E.y), make_float2(-Tt.y, -Tt.y));
  XA = __fadd2_rn(E, make_float2(-Tt.x, -Tt.y));
  XBc = __fadd2_rn(E, Tt);
}

// Inverse butterfly of slot r: Y[kA] = mA X[kA], conj Y[kB] = mB conj X[kB];  Ey = Y[kA] + conj Y[kB],
// Fy = Y[kA] - conj Y[kB], V = i conj(W) Fy = (-U.x, U.y) Fy;  Z'[kA] = Ey + V,  Z'[kB] = conj(Ey - V).
__device__ __forceinline__ void w5_unsplit(float2 XA, float2 XBc, float2 mk, float2 U, float2 &ZA, float2 &ZB) {
  const float2 YA = __fmul2_rn(XA, make_float2(mk.x, mk.x));
  const float2 YB = __fmul2_rn(XBc, make_float2(mk.y, mk.y));
  const float2 Ey = __fadd2_rn(YA, YB);
  const float2 Fy = __fadd2_rn(YA, make_float2(-YB.x, -YB.y));
  const float2 V = cmul(Fy, make_float2(-U.x, U.y));
  ZA = __fadd2_rn(Ey, V);
  ZB = __fadd2_rn(make_float2(Ey.x, -Ey.y), make_float2(-V.x, V.y));
}

template <int C, bool SCORE, int W, int CPS>
__global__ void __launch_bounds__(W * 32, CPS) wstrip512_kernel(const FusedArgs a) {
  using G = WStrip512Geom<C, SCORE, W>;
  constexpr int SHIFT = G::SHIFT, H = G::H, NSIG = G::NSIG, RING = G::RING, NV = G::NV, BINS = 257;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *win2 = reinterpret_cast<float2 *>(smem_raw);              // [32][10] 0.5 * analysis, pair (2q + 64 j, + 1) at [q][j]
  float2 *syn2 = win2 + 32 * kW5Pitch;                              // [32][10] synthesis
  float2 *tw1 = syn2 + 32 * kW5Pitch;                               // [32][10] W256^(q k0)
  float2 *tw2 = tw1 + 32 * kW5Pitch;                                // [4][8]   W32^((2h + e) k1)
  float2 *twu = tw2 + 32;                                           // [32][5]  i W512^(q + 32 r), r < 4; slot 4: bin 128
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + G::TABLE_BYTES + warp * G::WARP_BYTES;
  float *stage = reinterpret_cast<float *>(wbase);                  // [NSIG][RING][128]
  float2 *ex = reinterpret_cast<float2 *>(stage + G::STAGE_FLOATS);
  double *acc = reinterpret_cast<double *>(ex + kWxFloat2 * G::NEX) + lane;           // [NACC][32]
  float2 *pit2 = reinterpret_cast<float2 *>(acc - lane + 32 * G::NACC) + lane;        // [C * C][32]

  {
    // W512^m for any m from the plan's table of m = 0..256
    auto wpow = [&](int m) {
      const float2 w = a.tw_full[m & 255];
      return (m & 256) ? make_float2(-w.x, -w.y) : w;
    };
    for (int i = threadIdx.x; i < 256; i += W * 32) {
      const int q = i & 31, j = i >> 5;
      win2[q * kW5Pitch + j] = make_float2(a.win_half[2 * q + 64 * j], a.win_half[2 * q + 64 * j + 1]);
      syn2[q * kW5Pitch + j] = make_float2(a.syn[2 * q + 64 * j], a.syn[2 * q + 64 * j + 1]);
      tw1[(i >> 3) * kW5Pitch + (i & 7)] = wpow(2 * (((i >> 3) * (i & 7)) & 255));      // W256^e = W512^(2e)
    }
    if (threadIdx.x < 32) {
      const int h = threadIdx.x >> 3, k1 = (threadIdx.x & 7) >> 1, e = threadIdx.x & 1;
      tw2[threadIdx.x] = wpow(2 * ((8 * (2 * h + e) * k1) & 255));
    }
    for (int i = threadIdx.x; i < 32 * kW5U; i += W * 32) {
      const int q = i / kW5U, r = i - q * kW5U;
      const float2 w = wpow(r < 4 ? q + 32 * r : 128);
      twu[i] = make_float2(-w.y, w.x);                               // i W
    }
  }
  __syncthreads();                                                  // the only block-wide barrier

  const float2 *t1 = tw1 + lane * kW5Pitch;
  const float2 *t2 = tw2 + (lane >> 3) * 8;
  const float4 *winp = reinterpret_cast<const float4 *>(win2 + lane * kW5Pitch);
  const float4 *synp = reinterpret_cast<const float4 *>(syn2 + lane * kW5Pitch);
  const float2 *up = twu + lane * kW5U;
  const int T = a.T, S = a.tiles, I = a.strip_iters;
  const int total = a.batch * S;
  const float bw4 = lane == 0 ? 1.f : 0.f;                          // bin 128 lives on lane 0 only

  for (int strip = warp * gridDim.x + blockIdx.x; strip < total; strip += gridDim.x * W) {
    const int b = strip / S, s = strip - b * S;
    const int q0 = I / S, rem = I - q0 * S;
    const int n_it = q0 + (s < rem ? 1 : 0);
    const int a0 = s * q0 + min(s, rem) - H * s;                    // first frame transformed
    const int own_frame0 = s == 0 ? 0 : a0 + H;                     // frames counted by this strip (PIT)
    const int own_block0 = a0 + H;                                  // hop blocks written by this strip
    const float *mix_row = a.mix + static_cast<int64_t>(b) * a.n;
    const float *ref_row = SCORE ? a.refs + static_cast<int64_t>(b) * C * a.n : nullptr;
    const float *mask_b = a.masks + static_cast<int64_t>(b) * C * T * BINS;
    const int mask_q = T * BINS;
    const int n32 = static_cast<int>(a.n);                           // per-utterance sample indices fit 32 bits (checked by the host)
    const int n_valid = (SCORE && a.valid) ? min(a.valid[b], n32) : n32;
    int len_i = T;
    if (SCORE && a.lengths) len_i = static_cast<int>(a.lengths[b]);

    // hop block h (samples [128 h - pad, 128 h - pad + 128)) of every signal -> ring slot `slot` (zeros outside [0, n))
    auto issue_block = [&](int h, int slot) {
      const int g0 = h * SHIFT - a.pad;
#pragma unroll
      for (int sg = 0; sg < NSIG; ++sg) {
        const float *row = sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n;
        float *dst = stage + (sg * RING + slot) * SHIFT;
        if (a.vec_ok) {
          const int g = g0 + 4 * lane;
          const int bytes = g < 0 ? 0 : max(0, min(16, (n32 - g) * 4));
          cp_async16_zfill(dst + 4 * lane, bytes > 0 ? row + g : row, bytes);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int g = g0 + lane + 32 * i;
            const bool ok = g >= 0 && g < n32;
            cp_async4_zfill(dst + lane + 32 * i, ok ? row + g : row, ok ? 4 : 0);
          }
        }
      }
    };

    float2 mab[C][5];                                               // (mask at kA, mask at kB) per slot
    auto load_masks = [&](int ta) {
      // rows beyond T - 1 multiply all-zero spectra: any valid row will do (no predicates).  Two lane bases per
      // row (ascending kA run, descending kB run) and compile-time offsets: no per-load address arithmetic.
      const float *pa = mask_b + min(ta, T - 1) * BINS + lane;
      const float *pb = pa + (256 - 2 * lane);
#pragma unroll
      for (int i = 0; i < C; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) mab[i][r] = make_float2(__ldg(pa + 32 * r), __ldg(pb - 32 * r));
        const float m128 = __ldg(pa - lane + 128);
        mab[i][4] = make_float2(m128, m128);
        pa += mask_q;
        pb += mask_q;
      }
    };

    __syncwarp();                                                   // previous strip's reads are done
#pragma unroll
    for (int k = 0; k < 4; ++k) issue_block(a0 + k, k);
    cp_async_commit();
    load_masks(a0);

    float2 carry[C][6];                                             // unfinished overlap-add sums (three hop blocks)
#pragma unroll
    for (int i = 0; i < C; ++i) {
#pragma unroll
      for (int k = 0; k < 6; ++k) carry[i][k] = make_float2(0.f, 0.f);
    }
#pragma unroll
    float2 pitr[G::NPIT > 0 ? G::NPIT : 1];                          // PIT pair sums of this lane (registers: 8 warps per SM leave room)
#pragma unroll
    for (int i = 0; i < G::NPIT; ++i) pitr[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < G::NACC; ++i) acc[32 * i] = 0.0;
    double accr[SCORE ? 2 * C : 1];                                  // |e_i|^2, |r_i|^2 of this lane: registers
#pragma unroll
    for (int i = 0; i < (SCORE ? 2 * C : 1); ++i) accr[i] = 0.0;

    int slot0 = 0;                                                  // ring slot of hop block ta
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int ta = a0 + it;
      cp_async_wait<0>();                                           // hop block ta + 3 (requested an iteration ago) has landed
      __syncwarp();
      {
        // the free slot (it held hop block ta - 1) takes block ta + 4 while this frame is transformed
        const int free_slot = slot0 == 0 ? RING - 1 : slot0 - 1;
        if (it + 1 < n_it) issue_block(ta + 4, free_slot);
        cp_async_commit();
      }
      // this lane's sample pairs (2 lane + 64 j, + 1): hop block j / 2 of the frame, offset 2 lane + 64 (j & 1)
      int so[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sl = slot0 + k;
        so[k] = (sl >= RING ? sl - RING : sl) * SHIFT + 2 * lane;
      }
      auto windowed = [&](float2 (&vv)[8], int sg) {
        const float *sp = stage + sg * RING * SHIFT;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 w = winp[j / 2];
          const float2 x0 = *reinterpret_cast<const float2 *>(sp + so[j >> 1]);
          const float2 x1 = *reinterpret_cast<const float2 *>(sp + so[j >> 1] + 64);
          vv[j] = __fmul2_rn(x0, make_float2(w.x, w.y));
          vv[j + 1] = __fmul2_rn(x1, make_float2(w.z, w.w));
        }
      };

      float2 v[8];
      float2 XR[5], XI[5];                                          // mixture spectra, planar over (kA, kB) (labels)
      float2 XA[5], XBc[5];                                         // ... and complex: X[kA], conj X[kB] (inverse butterflies)
      float2 inv[5], mag[5];                                        // 1/|X|, gated |X|
      float pmin = 1.f;
      const float own_w = (ta >= own_frame0 && ta < T) ? 1.f : 0.f;
      const float gate = ta < len_i ? 1.f : 0.f;

      // label = |S| cos(angle X - angle S) = Re(S conj X) / |X| ; angle(0) = 0 -> Re S
      auto labels = [&](const float2 (&vv)[8], int j, bool anyzero) {
        float2 pj[C];
#pragma unroll
        for (int i = 0; i < C; ++i) pj[i] = make_float2(0.f, 0.f);
        auto bins = [&](auto zero_tag) {
          constexpr bool ZERO = decltype(zero_tag)::value;
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            float2 SR, SI;
            w5_split(vv, lane, r, up[r], SR, SI);
            float2 l = __fmul2_rn(__ffma2_rn(SR, XR[r], __fmul2_rn(SI, XI[r])), inv[r]);
            if (ZERO) {
              const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
              if (!(p.x > 0.f)) l.x = SR.x;
              if (!(p.y > 0.f)) l.y = SR.y;
            }
#pragma unroll
            for (int i = 0; i < C; ++i) {
              float2 d = __ffma2_rn(mab[i][r], mag[r], make_float2(-l.x, -l.y));
              if (r == 4) d = __fmul2_rn(d, make_float2(bw4, 0.f));
              pj[i] = __ffma2_rn(d, d, pj[i]);
            }
          }
        };
        if (!anyzero) bins(std::false_type{}); else bins(std::true_type{});
        const float2 own2 = make_float2(own_w, own_w);
#pragma unroll
        for (int i = 0; i < C; ++i) {                               // column j of the pair sums
          pitr[i * C + j] = __ffma2_rn(pj[i], own2, pitr[i * C + j]);
        }
      };
      auto mixture_spectra = [&]() {
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          if (SCORE) w5_split_both(v, lane, r, up[r], XR[r], XI[r], XA[r], XBc[r]);
          else {
            float2 A, P;
            w5_partner(v, lane, r, A, P);
            const float2 E = __fadd2_rn(A, make_float2(P.x, -P.y));
            const float2 Tt = cmul(__fadd2_rn(A, make_float2(-P.x, P.y)), up[r]);
            XA[r] = __fadd2_rn(E, make_float2(-Tt.x, -Tt.y));
            XBc[r] = __fadd2_rn(E, Tt);
          }
        }
        if (SCORE) {
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
            pmin = fminf(pmin, fminf(p.x, (r == 4) ? 1.f : p.y));
            // |X| = 0 -> a finite 1/|X| (and |X| * 1/|X| = 0); the exact-zero label is redone in `labels`
            inv[r] = make_float2(rsqrt_fast(fmaxf(p.x, 1e-36f)), rsqrt_fast(fmaxf(p.y, 1e-36f)));
            mag[r] = __fmul2_rn(__fmul2_rn(p, inv[r]), make_float2(gate, gate));
          }
        }
      };

      // ---- forward: mixture (+ reference 0 in lockstep), then the remaining references ----
      windowed(v, 0);
      if constexpr (SCORE) {
        float2 vb[8];
        windowed(vb, 1);
        wfft256x2<false>(v, vb, t1, t2, ex, lane);
        mixture_spectra();
        const bool anyzero = __any_sync(0xffffffffu, !(pmin > 0.f));
        labels(vb, 0, anyzero);
        if constexpr (C == 2) {
          windowed(vb, 2);
          wfft256<false>(vb, t1, t2, ex, lane);
          labels(vb, 1, anyzero);
        } else if constexpr (C >= 3) {
          float2 vc[8];
          windowed(vb, 2);
          windowed(vc, 3);
          wfft256x2<false>(vb, vc, t1, t2, ex, lane);
          labels(vb, 1, anyzero);
          labels(vc, 2, anyzero);
        }
      } else {
        wfft256<false>(v, t1, t2, ex, lane);
        mixture_spectra();
      }

      // ---- masked spectra -> time frame -> in-register overlap-add -> HBM ----
      const int gb = ta * SHIFT - a.pad;                            // first sample of hop block ta
      const bool owned = ta >= own_block0;
      const bool plain = owned && gb + SHIFT <= n_valid && a.vec_ok;
      // Hermitian butterfly of estimate qi: Z'[q + 32 j], j = 0..7
      auto merged = [&](float2 (&vv)[8], auto qc) {
        constexpr int qi = decltype(qc)::value;
        float2 ZB[4], Z4 = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          float2 za, zb;
          w5_unsplit(XA[r], XBc[r], mab[qi][r], up[r], za, zb);
          if (r < 4) { vv[r] = za; ZB[r] = zb; } else Z4 = za;
        }
        const int src = (32 - lane) & 31;
#pragma unroll
        for (int j = 4; j < 8; ++j) {
          float2 got;
          got.x = __shfl_sync(0xffffffffu, ZB[7 - j].x, src);
          got.y = __shfl_sync(0xffffffffu, ZB[7 - j].y, src);
          const float2 own = (j == 4) ? Z4 : ZB[8 - j];
          vv[j] = (lane == 0) ? own : got;
        }
      };
      // time frame of estimate qi -> overlap-add -> hop block ta -> HBM; the finished block is kept for the Gram pass
      float2 yo[C][2];
      auto finish = [&](const float2 (&vv)[8], auto qc) {
        constexpr int qi = decltype(qc)::value;
        float2 y[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 w = synp[j / 2];
          y[j] = __ffma2_rn(vv[j], make_float2(w.x, w.y), j < 6 ? carry[qi][j] : make_float2(0.f, 0.f));
          y[j + 1] = __ffma2_rn(vv[j + 1], make_float2(w.z, w.w), j + 1 < 6 ? carry[qi][j + 1] : make_float2(0.f, 0.f));
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) carry[qi][k] = y[k + 2];
        yo[qi][0] = y[0];
        yo[qi][1] = y[1];
        float *out = a.est ? a.est + (static_cast<int64_t>(b) * C + qi) * a.n + gb + 2 * lane : nullptr;
        if (out == nullptr || !owned) return;
        if (plain) {
#pragma unroll
          for (int m = 0; m < 2; ++m) *reinterpret_cast<float2 *>(out + 64 * m) = y[m];
        } else {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const int g = gb + 2 * lane + 64 * m;
            if (g >= 0 && g < n32) out[64 * m] = y[m].x;
            if (g + 1 >= 0 && g + 1 < n32) out[64 * m + 1] = y[m].y;
          }
        }
      };
      // Gram statistics <e_i, r_j>, |e_i|^2, |r_j|^2 of hop block ta for ALL estimates at once: every reference sample
      // is read and converted once (float64: products of float32 are exact)
      auto gram_all = [&]() {
        if (!SCORE || !owned) return;
        const float *rf = stage + RING * SHIFT + so[0];             // reference 0 at hop block ta, this lane's pairs
        double e[C][4], r[C][4];
        if (plain) {                                                // warp-uniform: whole hop block inside [0, n_valid)
#pragma unroll
          for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int j = 0; j < C; ++j) {
              const float2 r2 = *reinterpret_cast<const float2 *>(rf + j * RING * SHIFT + 64 * m);
              r[j][2 * m] = static_cast<double>(r2.x);
              r[j][2 * m + 1] = static_cast<double>(r2.y);
              e[j][2 * m] = static_cast<double>(yo[j][m].x);
              e[j][2 * m + 1] = static_cast<double>(yo[j][m].y);
            }
          }
        } else {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const int g = gb + 2 * lane + 64 * m;
            const bool ok0 = g >= 0 && g < n_valid, ok1 = g + 1 >= 0 && g + 1 < n_valid;
#pragma unroll
            for (int j = 0; j < C; ++j) {
              const float2 r2 = *reinterpret_cast<const float2 *>(rf + j * RING * SHIFT + 64 * m);
              r[j][2 * m] = ok0 ? static_cast<double>(r2.x) : 0.0;
              r[j][2 * m + 1] = ok1 ? static_cast<double>(r2.y) : 0.0;
              e[j][2 * m] = ok0 ? static_cast<double>(yo[j][m].x) : 0.0;
              e[j][2 * m + 1] = ok1 ? static_cast<double>(yo[j][m].y) : 0.0;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
          double ee = 0.0, rr = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { ee = fma(e[i][k], e[i][k], ee); rr = fma(r[i][k], r[i][k], rr); }
          accr[i] += ee;
          accr[C + i] += rr;
#pragma unroll
          for (int j = 0; j < C; ++j) {
            double gq = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) gq = fma(e[i][k], r[j][k], gq);
            acc[32 * (i * C + j)] += gq;
          }
        }
      };
      using Q0 = std::integral_constant<int, 0>;
      using Q1 = std::integral_constant<int, C >= 2 ? 1 : 0>;
      using Q2 = std::integral_constant<int, C >= 3 ? 2 : 0>;
      if constexpr (C == 1) {
        merged(v, Q0{});
        if (it + 1 < n_it) load_masks(ta + 1);
        wfft256<true>(v, t1, t2, ex, lane);
        finish(v, Q0{});
        gram_all();
      } else {
        float2 vb[8];
        merged(v, Q0{});
        merged(vb, Q1{});
        if constexpr (C == 2) {
          if (it + 1 < n_it) load_masks(ta + 1);                    // next frame's masks, a transform ahead
          wfft256x2<true>(v, vb, t1, t2, ex, lane);
          finish(v, Q0{});
          finish(vb, Q1{});
          gram_all();
        } else {
          wfft256x2<true>(v, vb, t1, t2, ex, lane);
          finish(v, Q0{});
          finish(vb, Q1{});
          merged(v, Q2{});
          if (it + 1 < n_it) load_masks(ta + 1);
          wfft256<true>(v, t1, t2, ex, lane);
          finish(v, Q2{});
          gram_all();
        }
      }
      slot0 = slot0 + 1 == RING ? 0 : slot0 + 1;
    }
    cp_async_wait<0>();

    if (SCORE) {
      double vals[NV];
#pragma unroll
      for (int i = 0; i < C * C; ++i) {
        const float2 pp = pitr[i];
        vals[i] = static_cast<double>(pp.x) + static_cast<double>(pp.y);
      }
#pragma unroll
      for (int i = 0; i < C * C; ++i) vals[C * C + i] = acc[32 * i];
#pragma unroll
      for (int i = 0; i < 2 * C; ++i) vals[2 * C * C + i] = accr[i];
#pragma unroll
      for (int i = 0; i < NV; ++i) vals[i] = warp_sum(vals[i]);
      if (lane == 0) {
        double *dst = a.partials + static_cast<int64_t>(strip) * NV;
#pragma unroll
        for (int i = 0; i < NV; ++i) dst[i] = vals[i];
      }
      if (a.counters != nullptr) finalize_by_warp<C>(a, b, lane);
    }
  }
}

template <int C, bool SCORE>
static int launch_wstrip512(const sep_plan *p, FusedArgs a, int batch, double *d_scores, double *d_sums,
                            Scratch &s, cudaStream_t stream) {
  constexpr int W = 4, CPS = 2;                                     // two 4-warp CTAs per SM (<= 113 KB of shared memory each)
  using G = WStrip512Geom<C, SCORE, W>;
  const int sms = p->sm_count > 0 ? p->sm_count : 148;
  // strips planned for ONE wave of 4-warp CTAs at small batches (see fused_wstrip.cu): fewer, longer strips
  // (three halo frames each); in a stream of independent steps the other launches fill the SMs
  static const int plan_env = getenv("SEPCORE_WSTRIP_WARPS") ? atoi(getenv("SEPCORE_WSTRIP_WARPS")) : 0;
  const int plan_warps = plan_env > 0 ? plan_env : (batch >= 2 * sms ? 2 * sms * W : sms * W);
  pick_strips(a.T, G::H, 1, batch, plan_warps, &a.tiles, &a.strip_iters);
  int rc;
  double *partials = nullptr;
  int *counters = nullptr;
  static const bool single_launch = !(getenv("SEPCORE_SINGLE_LAUNCH") && atoi(getenv("SEPCORE_SINGLE_LAUNCH")) == 0);
  if (SCORE) {
    if ((rc = s.alloc(&counters, static_cast<size_t>(batch) + 1))) return rc;
    if ((rc = s.alloc(&partials, static_cast<size_t>(batch) * a.tiles * G::NV))) return rc;
    if (!single_launch) counters = nullptr;
    else if ((rc = reset_counters(counters, batch, s, stream))) return rc;
  }
  a.partials = partials;
  a.counters = counters;
  a.scores = d_scores;
  a.sums = d_sums;
  const auto aligned = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  a.vec_ok = (a.n % 4 == 0) && aligned(a.mix) && (!a.refs || aligned(a.refs)) && (!a.est || aligned(a.est)) ? 1 : 0;
  const size_t smem = G::smem();
  SEP_CUDA(cudaFuncSetAttribute(wstrip512_kernel<C, SCORE, W, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t total = static_cast<int64_t>(batch) * a.tiles;
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(sms) * CPS, (total + W - 1) / W));
  profile_begin(stream, "wstrip512_kernel<C=%d,SCORE=%d,W=%d,CTAS_PER_SM=%d>", C, int(SCORE), W, CPS);
  wstrip512_kernel<C, SCORE, W, CPS><<<grid, W * 32, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (SCORE && counters == nullptr) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

int fused_wstrip512_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                        double *d_sums, Scratch &s, cudaStream_t stream, bool *handled) {
  *handled = false;
  if (p->size != 512 || p->shift != 128 || C > 3 || a.T < 4) return SEP_OK;
  if (a.n > (int64_t(1) << 30)) return SEP_OK;                      // 32-bit sample indices inside the kernel
  if (getenv("SEPCORE_FORCE_GENERIC") || getenv("SEPCORE_FORCE_TILES") || getenv("SEPCORE_FORCE_HALFWARP")) return SEP_OK;
  *handled = true;
  const bool score = a.refs != nullptr;
  switch (C) {
    case 1: return score ? launch_wstrip512<1, true>(p, a, batch, d_scores, d_sums, s, stream)
                         : launch_wstrip512<1, false>(p, a, batch, d_scores, d_sums, s, stream);
    case 2: return score ? launch_wstrip512<2, true>(p, a, batch, d_scores, d_sums, s, stream)
                         : launch_wstrip512<2, false>(p, a, batch, d_scores, d_sums, s, stream);
    default: return score ? launch_wstrip512<3, true>(p, a, batch, d_scores, d_sums, s, stream)
                          : launch_wstrip512<3, false>(p, a, batch, d_scores, d_sums, s, stream);
  }
}

}  // namespace sep
