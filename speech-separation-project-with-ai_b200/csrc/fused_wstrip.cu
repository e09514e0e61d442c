// fused_wstrip.cu -- the hot path for size = 256 (shift 128 or 64) and two sources, second generation:
// warp-autonomous sliding strips with ONE frame pair per WARP (fft256w.cuh: 8 points per lane).
//
// Why: the half-warp strips of fused_strip.cu need ~250 registers per thread, so an SM holds 8 warps
// (two per scheduler) and issues 0.43 instructions per cycle -- the kernel is bound by instruction
// latency, not by bytes (profiles/r1_ncu_strip_kernels.md).  Spreading a frame pair over 32 lanes halves
// every per-lane array (transform, spectra, masks, carries): 168 registers, 12 warps per SM.  What else
// changes with the layout:
//   * lane q holds samples q + 32 m and bins q + 32 j: span reads, mask rows (128-byte runs) and
//     estimate stores are all lane-contiguous; no skewed shared layouts are needed;
//   * the overlap-add needs no shared memory at all: the two frames of a pair overlap inside the lane
//     (frame t+1 sample q + 32 m is frame t sample q + 32 (m + SHIFT/32)), and the tail that the NEXT
//     pair needs is the same lane's, so it is carried in registers from one iteration to the next;
//   * with 8 points per lane a transform has half the independent instructions between its exchanges, so
//     the two reference transforms -- and the two inverse transforms -- of an iteration run in LOCKSTEP
//     (wfft256x2): the shared-memory round trips of one hide behind the butterflies of the other;
//   * the float64 Gram accumulators and the PIT pair sums live in per-lane shared-memory slots (registers
//     are the scarce resource; one read-modify-write per iteration);
//   * a strip recomputes R - 1 halo frames at its start, as before (atomic-free: every output sample
//     is written by exactly one lane, once).
// Staging (zero-filling cp.async, every span single-buffered and requested as soon as its buffer is
// free), the PIT / Gram arithmetic, the partial rows and the finalisation are those of fused_strip.cu,
// so results agree with the half-warp kernel to float32 round-off
// (tests/test_gpu_parity.py::test_fused_two_strip_kernels_agree).
// Measured (profiles/r1_ncu_wstrip.md): 21.9 us per cfg2 step in the replayed loop against 24.8 us for the
// half-warp strips (separate finalisation kernels), 180 against 206 us at batch 512; with one source the half-warp
// strips stay faster.
// Reference lines: see fused.cu.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fft256w.cuh"
#include "fused.cuh"
#include "strip_common.cuh"

namespace sep {

constexpr int kWWinPitch = 12;                     // window rows [32 lanes][12]: conflict-free LDS.128

template <int C, int R, bool SCORE, int W, bool DUAL>
struct WStripGeom {
  static constexpr int SHIFT = 256 / R, H = R - 1;
  static constexpr int NSIG = SCORE ? 1 + C : 1;
  static constexpr int D = SHIFT / 32;             // register distance between the frames of a pair
  static constexpr int NY = 8 + D;                 // pair-summed values per lane
  static constexpr int NOUT = 2 * D;               // finished values per lane and iteration (two hop blocks)
  static constexpr int NCARRY = NY - NOUT;         // values carried to the next iteration (registers)
  static constexpr int SPAN = SHIFT + 256;         // samples under the two frames of an iteration
  static constexpr int STAGE_FLOATS = NSIG * SPAN;
  static constexpr int NACC = SCORE ? C * C + 2 * C : 0;   // Gram statistics: per-lane float64 accumulators in shared memory
  static constexpr int NPIT = SCORE ? C * C : 0;           // PIT pair sums, per lane (frame ta part, frame tb part)
  static constexpr int NEX = (C == 2 && DUAL) ? 2 : 1;    // exchange buffers (two transforms in lockstep)
  static constexpr int WARP_BYTES = 4 * STAGE_FLOATS + 8 * kWxFloat2 * NEX + 8 * 32 * NACC + 8 * 32 * NPIT;
  static constexpr int TABLE_BYTES = 4 * 2 * 32 * kWWinPitch + 8 * 32 * kWTw1 + 8 * 32;
  static constexpr size_t smem() { return TABLE_BYTES + static_cast<size_t>(W) * WARP_BYTES; }
  static constexpr int NV = FusedVals<C>::NV;
};

// Built and tested for C = 2, DUAL = true (dispatch_wstrip); the C = 1 / one-transform-at-a-time branches are what the
// measurements in profiles/r1_ncu_wstrip.md compared against and are not instantiated in the library.
template <int C, int R, bool SCORE, int W, int CPS, bool DUAL>
__global__ void __launch_bounds__(W * 32, CPS) wstrip256_kernel(const FusedArgs a) {
  using G = WStripGeom<C, R, SCORE, W, DUAL>;
  constexpr int SHIFT = G::SHIFT, H = G::H, NSIG = G::NSIG, D = G::D, NY = G::NY, NOUT = G::NOUT;
  constexpr int NCARRY = G::NCARRY, SPAN = G::SPAN, NV = G::NV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *win = reinterpret_cast<float *>(smem_raw);               // [32][12] 0.5 * analysis, tap q + 32 m at [q][m]
  float *syn = win + 32 * kWWinPitch;                             // [32][12] synthesis
  float2 *tw1 = reinterpret_cast<float2 *>(syn + 32 * kWWinPitch);  // [32][10] W256^(q k0)
  float2 *tw2 = tw1 + 32 * kWTw1;                                 // [4][8]   W32^((2h + e) k1)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + G::TABLE_BYTES + warp * G::WARP_BYTES;
  float *stage = reinterpret_cast<float *>(wbase);                // [NSIG][SPAN]
  float2 *ex = reinterpret_cast<float2 *>(stage + G::STAGE_FLOATS);
  // Gram statistics <e_i, r_j>, |e_i|^2, |r_j|^2 of this lane: [NACC][32] float64 (registers are the scarce resource)
  double *acc = reinterpret_cast<double *>(ex + kWxFloat2 * G::NEX) + (threadIdx.x & 31);
  float2 *pit2 = reinterpret_cast<float2 *>(acc - (threadIdx.x & 31) + 32 * G::NACC) + (threadIdx.x & 31);   // [C * C][32]

  {
    // W256^e from the plan's table of e = 0..128
    auto wpow = [&](int e) {
      const float2 w = a.tw_full[e & 127];
      return (e & 128) ? make_float2(-w.x, -w.y) : w;
    };
    for (int i = threadIdx.x; i < 256; i += W * 32) {
      win[(i & 31) * kWWinPitch + (i >> 5)] = a.win_half[i];
      syn[(i & 31) * kWWinPitch + (i >> 5)] = a.syn[i];
      tw1[(i >> 3) * kWTw1 + (i & 7)] = wpow(((i >> 3) * (i & 7)) & 255);
    }
    if (threadIdx.x < 32) {
      const int h = threadIdx.x >> 3, k1 = (threadIdx.x & 7) >> 1, e = threadIdx.x & 1;
      tw2[threadIdx.x] = wpow((8 * (2 * h + e) * k1) & 255);
    }
  }
  __syncthreads();                                                // the only block-wide barrier

  const float2 *t1 = tw1 + lane * kWTw1;
  const float2 *t2 = tw2 + (lane >> 3) * 8;
  const float4 *winp = reinterpret_cast<const float4 *>(win + kWWinPitch * lane);
  const float4 *synp = reinterpret_cast<const float4 *>(syn + kWWinPitch * lane);
  const int T = a.T, S = a.tiles, I = a.strip_iters;
  const int total = a.batch * S;
  const float bw4 = lane == 0 ? 1.f : 0.f;                        // bin 128 lives on lane 0 only

  for (int strip = warp * gridDim.x + blockIdx.x; strip < total; strip += gridDim.x * W) {
    const int b = strip / S, s = strip - b * S;
    const int q0 = I / S, rem = I - q0 * S;
    const int n_it = q0 + (s < rem ? 1 : 0);
    const int a0 = 2 * (s * q0 + min(s, rem)) - H * s;            // first frame transformed
    const int own_frame0 = s == 0 ? 0 : a0 + H;                   // frames counted by this strip (PIT)
    const int own_block0 = a0 + H;                                // hop blocks written by this strip
    const float *mix_row = a.mix + static_cast<int64_t>(b) * a.n;
    const float *ref_row = SCORE ? a.refs + static_cast<int64_t>(b) * C * a.n : nullptr;
    const float *mask_b = a.masks + static_cast<int64_t>(b) * C * T * 129;
    const int mask_q = T * 129;
    const int n32 = static_cast<int>(a.n);                         // per-utterance sample indices fit 32 bits (checked by the host)
    const int n_valid = (SCORE && a.valid) ? min(a.valid[b], n32) : n32;
    int len_i = T;
    if (SCORE && a.lengths) len_i = static_cast<int>(a.lengths[b]);

    // signals sig_lo..sig_hi of iteration `it` -> shared memory (zero-filled outside [0, n)); one commit group
    auto issue_span = [&](int it, int sig_lo, int sig_hi) {
      const int g0 = (a0 + 2 * it) * SHIFT - a.pad;
      if (a.vec_ok && g0 >= 0 && g0 + SPAN <= n32) {                // interior span: plain 16-byte copies
#pragma unroll
        for (int sg = 0; sg < NSIG; ++sg) {
          if (sg < sig_lo || sg >= sig_hi) continue;
          float *dst = stage + sg * SPAN + 4 * lane;
          const float *src = (sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n) + g0 + 4 * lane;
#pragma unroll
          for (int c0 = 0; c0 < SPAN / 4; c0 += 32)
            if (c0 + 32 <= SPAN / 4 || lane < SPAN / 4 - c0) cp_async16(dst + 4 * c0, src + 4 * c0);
        }
      } else {
#pragma unroll 1
        for (int sg = sig_lo; sg < sig_hi; ++sg) {
          const float *row = sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n;
          float *dst = stage + sg * SPAN;
          if (a.vec_ok) {
            for (int c = lane; c < SPAN / 4; c += 32) {
              const int g = g0 + 4 * c;
              const int bytes = g < 0 ? 0 : max(0, min(16, (n32 - g) * 4));
              cp_async16_zfill(dst + 4 * c, bytes > 0 ? row + g : row, bytes);
            }
          } else {
            for (int i = lane; i < SPAN; i += 32) {
              const int g = g0 + i;
              const bool ok = g >= 0 && g < n32;
              cp_async4_zfill(dst + i, ok ? row + g : row, ok ? 4 : 0);
            }
          }
        }
      }
      cp_async_commit();
    };

    float2 mab[C][5];                                             // (mask of frame ta, mask of frame tb)
    auto load_masks = [&](int ta) {
      // rows beyond T - 1 multiply all-zero spectra: any valid row will do (no predicates)
      const float *pa = mask_b + min(ta, T - 1) * 129 + lane;
      const float *pb = mask_b + min(ta + 1, T - 1) * 129 + lane;
#pragma unroll
      for (int i = 0; i < C; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) mab[i][r] = make_float2(__ldg(pa + 32 * r), __ldg(pb + 32 * r));
        mab[i][4] = make_float2(__ldg(pa - lane + 128), __ldg(pb - lane + 128));
        pa += mask_q;
        pb += mask_q;
      }
    };

    __syncwarp();                                                 // previous strip's reads are done
    issue_span(0, 0, 1);
    load_masks(a0);

    float carry[C][NCARRY];                                       // unfinished overlap-add sums
#pragma unroll
    for (int i = 0; i < C; ++i) {
#pragma unroll
      for (int k = 0; k < NCARRY; ++k) carry[i][k] = 0.f;
    }
#pragma unroll
    float2 pitr[G::NPIT > 0 ? G::NPIT : 1];                       // PIT pair sums of this lane
#pragma unroll
    for (int i = 0; i < G::NPIT; ++i) pitr[i] = make_float2(0.f, 0.f);
    // |e_0|^2, |e_1|^2, |r_0|^2, |r_1|^2 of this lane in registers at shift 64 (measured: 39.65 -> 39.04 us; at shift 128
    // the same change costs 20.67 -> 20.83 us, so there they stay in shared memory)
    constexpr bool ACCR = R == 4;
    double accr[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < G::NACC; ++i) acc[32 * i] = 0.0;

#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int ta = a0 + 2 * it, tb = ta + 1;                      // this iteration's frame pair
      __syncwarp();                                                 // reads of the buffers refilled below are done
      if (NSIG > 1) issue_span(it, 1, NSIG);                        // references of this iteration
      if (NSIG > 1) cp_async_wait<1>(); else cp_async_wait<0>();    // this iteration's mixture (requested an iteration ago) has landed
      __syncwarp();
      const float *st = stage + lane;                               // signal sg, sample lane + 32 j at st[sg * SPAN + 32 j]

      float2 v[8];
      float2 XR[5], XI[5];                                          // mixture spectra, planar: (X_ta, X_tb)
      float2 inv[5], mag[5];                                        // 1/|X|, gated |X|
      float pmin = 1.f;
      const float2 own2 = make_float2((ta >= own_frame0 && ta < T) ? 1.f : 0.f,
                                      (tb >= own_frame0 && tb < T) ? 1.f : 0.f);
      const float2 gate2 = make_float2(ta < len_i ? 1.f : 0.f, tb < len_i ? 1.f : 0.f);

      // ---- mixture: windowed frame pair -> spectra ----
      {
        bool silent = false;                                        // a mixture frame of the pair is all zeros
        float2 live2 = make_float2(1.f, 1.f);
        {
          float xs[NY];
#pragma unroll
          for (int j = 0; j < NY; ++j) xs[j] = st[32 * j];
          const float4 w0 = winp[0], w1 = winp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int m = 0; m < 8; ++m) v[m] = make_float2(xs[m] * w[m], xs[m + D] * w[m]);   // scalar: no (w, w) splats
          if (SCORE) {
            // Two frames ride in one complex transform, so an all-zero frame next to a non-zero one comes out
            // of the split as round-off, not as exact zeros -- and the reference's angle(0) = 0 label rule
            // (label = Re S) must see exact zeros.  Digital silence is detected in the time domain.
            unsigned oa = 0u, ob = 0u;
#pragma unroll
            for (int m = 0; m < 8; ++m) { oa |= __float_as_uint(xs[m]); ob |= __float_as_uint(xs[m + D]); }
            const bool la = __any_sync(0xffffffffu, (oa << 1) != 0u), lb = __any_sync(0xffffffffu, (ob << 1) != 0u);
            silent = !(la && lb);
            live2 = make_float2(la ? 1.f : 0.f, lb ? 1.f : 0.f);
          }
        }
        __syncwarp();                                               // the mixture span is in registers: refill it
        if (it + 1 < n_it) issue_span(it + 1, 0, 1); else cp_async_commit();
        wfft256<false>(v, t1, t2, ex, lane);
#pragma unroll
        for (int r = 0; r < 5; ++r) wsplit_planar(v, lane, r, XR[r], XI[r]);
        if (SCORE && silent) {                                      // warp-uniform, rare
#pragma unroll
          for (int r = 0; r < 5; ++r) { XR[r] = __fmul2_rn(XR[r], live2); XI[r] = __fmul2_rn(XI[r], live2); }
        }
        if (SCORE) {
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
            pmin = fminf(pmin, fminf(p.x, p.y));
            // |X| = 0 -> a finite 1/|X| (and |X| * 1/|X| = 0); the exact-zero label is redone below
            inv[r] = make_float2(rsqrt_fast(fmaxf(p.x, 1e-36f)), rsqrt_fast(fmaxf(p.y, 1e-36f)));
            mag[r] = __fmul2_rn(__fmul2_rn(p, inv[r]), gate2);
          }
        }
      }

      // ---- references: spectra -> PSA labels -> PIT pair sums ----
      if constexpr (SCORE) {
        cp_async_wait<1>();                                         // the references have landed
        __syncwarp();
        const bool anyzero = __any_sync(0xffffffffu, !(pmin > 0.f));
        auto windowed = [&](float2 (&vv)[8], const float *sp) {
          float xs[NY];
#pragma unroll
          for (int j = 0; j < NY; ++j) xs[j] = sp[32 * j];
          const float4 w0 = winp[0], w1 = winp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int m = 0; m < 8; ++m) vv[m] = make_float2(xs[m] * w[m], xs[m + D] * w[m]);
        };
        // label = |S| cos(angle X - angle S) = Re(S conj X) / |X| ; angle(0) = 0 -> Re S
        auto labels = [&](const float2 (&vv)[8], int j) {
          float2 pj[C];
#pragma unroll
          for (int i = 0; i < C; ++i) pj[i] = make_float2(0.f, 0.f);
          // one straight-line bin loop per case (ordinary / some mixture bin exactly zero)
          auto bins = [&](auto zero_tag) {
            constexpr bool ZERO = decltype(zero_tag)::value;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
              float2 SR, SI;
              wsplit_planar(vv, lane, r, SR, SI);
              float2 l = __fmul2_rn(__ffma2_rn(SR, XR[r], __fmul2_rn(SI, XI[r])), inv[r]);
              if (ZERO) {
                const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
                if (!(p.x > 0.f)) l.x = SR.x;
                if (!(p.y > 0.f)) l.y = SR.y;
              }
#pragma unroll
              for (int i = 0; i < C; ++i) {
                float2 d = __ffma2_rn(mab[i][r], mag[r], make_float2(-l.x, -l.y));
                if (r == 4) d = __fmul2_rn(d, make_float2(bw4, bw4));
                pj[i] = __ffma2_rn(d, d, pj[i]);
              }
            }
          };
          if (!anyzero) bins(std::false_type{}); else bins(std::true_type{});
#pragma unroll
          for (int i = 0; i < C; ++i) {                             // column j of the pair sums
            pitr[i * C + j] = __ffma2_rn(pj[i], own2, pitr[i * C + j]);
          }
        };
        if constexpr (C == 2 && DUAL) {
          float2 vb[8];
          windowed(v, st + SPAN);
          windowed(vb, st + 2 * SPAN);
          wfft256x2<false>(v, vb, t1, t2, ex, lane);
          labels(v, 0);
          labels(vb, 1);
        } else {
#pragma unroll 1
          for (int j = 0; j < C; ++j) {
            windowed(v, st + (1 + j) * SPAN);
            wfft256<false>(v, t1, t2, ex, lane);
            labels(v, j);
          }
        }
      }

      // ---- masked spectra -> time frames -> in-register overlap-add -> HBM ----
      const int gb = ta * SHIFT - a.pad + lane;                     // sample index of y[0]
      // leading values that belong to the previous strip (or to the fade-in padding)
      const int skip = min(NOUT, max(0, own_block0 - ta) * (NOUT / 2));
      const int room = n32 - gb, room_v = n_valid - gb;
      const int lim = room <= 0 ? 0 : min(NOUT, (room + 31) >> 5);
      const int lim_v = room_v <= 0 ? 0 : min(NOUT, (room_v + 31) >> 5);
      const bool plain = !__any_sync(0xffffffffu, !(skip == 0 && lim_v == NOUT));
      // spectrum of estimate q for both frames -> Hermitian-extended input of the inverse transform
      auto masked = [&](float2 (&vv)[8], auto qc) {
        constexpr int q = decltype(qc)::value;
        float2 L[5], Mi[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const float2 mk = mab[q][r];
          // (P, Q) = (m_a X_a, m_b X_b):  mR = (Re P, Re Q), mI = (Im P, Im Q)
          const float2 mR = __fmul2_rn(XR[r], mk), mI = __fmul2_rn(XI[r], mk);
          L[r] = __fadd2_rn(mR, make_float2(-mI.y, mI.x));           // P + i Q
          Mi[r] = __fadd2_rn(mR, make_float2(mI.y, -mI.x));          // conj P + i conj Q (for the mirror bin)
        }
        wmerge_pair(vv, lane, L, Mi);
      };
      // time frames of estimate q -> overlap-add -> HBM, Gram statistics
      auto finish = [&](const float2 (&vv)[8], auto qc) {
        constexpr int q = decltype(qc)::value;
        float y[NY];
        {
          // synthesis window, the in-lane overlap of the pair and the carried tail of the previous pair (this
          // lane's own: registers) as scalar multiply-adds
          const float4 w0 = synp[0], w1 = synp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int mm = 0; mm < NY; ++mm) {
            float t = mm < NCARRY ? carry[q][mm] : 0.f;
            if (mm >= D) t = fmaf(vv[mm - D].y, w[mm - D], t);
            if (mm < 8) t = fmaf(vv[mm].x, w[mm], t);
            y[mm] = t;
          }
        }
#pragma unroll
        for (int k = 0; k < NCARRY; ++k) carry[q][k] = y[NOUT + k];
        float *out = a.est ? a.est + (static_cast<int64_t>(b) * C + q) * a.n + gb : nullptr;
        double gq[C], eq = 0.0, rq[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { gq[j] = 0.0; rq[j] = 0.0; }
        if (plain) {
          if (out) {
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) out[32 * mm] = y[mm];
          }
          if (SCORE) {
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) {
              const double e = static_cast<double>(y[mm]);
              eq = fma(e, e, eq);
#pragma unroll
              for (int j = 0; j < C; ++j) {
                const double r = static_cast<double>(st[(1 + j) * SPAN + 32 * mm]);
                gq[j] = fma(e, r, gq[j]);
                if (q == 0) rq[j] = fma(r, r, rq[j]);
              }
            }
          }
        } else {
#pragma unroll
          for (int mm = 0; mm < NOUT; ++mm) {
            const bool okw = mm >= skip && mm < lim;
            if (out && okw) out[32 * mm] = y[mm];
            if (SCORE) {
              const bool okv = mm >= skip && mm < lim_v;
              const double e = okv ? static_cast<double>(y[mm]) : 0.0;
              eq = fma(e, e, eq);
#pragma unroll
              for (int j = 0; j < C; ++j) {
                const double r = okv ? static_cast<double>(st[(1 + j) * SPAN + 32 * mm]) : 0.0;
                gq[j] = fma(e, r, gq[j]);
                if (q == 0) rq[j] = fma(r, r, rq[j]);
              }
            }
          }
        }
        if (SCORE) {
          acc[32 * (C * C + q)] += eq;
#pragma unroll
          for (int j = 0; j < C; ++j) acc[32 * (q * C + j)] += gq[j];
          if (q == 0) {
#pragma unroll
            for (int j = 0; j < C; ++j) acc[32 * (C * C + C + j)] += rq[j];
          }
        }
      };
      // both estimates at once (two sources): the reference samples are read and converted once for the four
      // Gram products instead of once per estimate
      auto finish2 = [&](const float2 (&va)[8], const float2 (&vb)[8]) {
        float y0[NY], y1[NY];
        {
          const float4 w0 = synp[0], w1 = synp[1];
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int mm = 0; mm < NY; ++mm) {
            float t0 = mm < NCARRY ? carry[0][mm] : 0.f, t1 = mm < NCARRY ? carry[C - 1][mm] : 0.f;
            if (mm >= D) { t0 = fmaf(va[mm - D].y, w[mm - D], t0); t1 = fmaf(vb[mm - D].y, w[mm - D], t1); }
            if (mm < 8) { t0 = fmaf(va[mm].x, w[mm], t0); t1 = fmaf(vb[mm].x, w[mm], t1); }
            y0[mm] = t0;
            y1[mm] = t1;
          }
        }
#pragma unroll
        for (int k = 0; k < NCARRY; ++k) { carry[0][k] = y0[NOUT + k]; carry[C - 1][k] = y1[NOUT + k]; }
        float *out0 = a.est ? a.est + static_cast<int64_t>(b) * C * a.n + gb : nullptr;
        float *out1 = a.est ? out0 + a.n : nullptr;
        // one straight-line copy per case: whole hop blocks inside the utterance / blocks on an edge
        auto emit = [&](auto plain_tag) {
          constexpr bool PLAIN = decltype(plain_tag)::value;
          if (out0) {
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) {
              if (PLAIN || (mm >= skip && mm < lim)) { out0[32 * mm] = y0[mm]; out1[32 * mm] = y1[mm]; }
            }
          }
          if (SCORE) {
            double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0, e0 = 0.0, e1 = 0.0, r0 = 0.0, r1 = 0.0;
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) {
              const bool okv = PLAIN || (mm >= skip && mm < lim_v);
              const double a0 = okv ? static_cast<double>(y0[mm]) : 0.0, a1 = okv ? static_cast<double>(y1[mm]) : 0.0;
              const double b0 = okv ? static_cast<double>(st[SPAN + 32 * mm]) : 0.0;
              const double b1 = okv ? static_cast<double>(st[2 * SPAN + 32 * mm]) : 0.0;
              e0 = fma(a0, a0, e0); e1 = fma(a1, a1, e1);
              g00 = fma(a0, b0, g00); g01 = fma(a0, b1, g01);
              g10 = fma(a1, b0, g10); g11 = fma(a1, b1, g11);
              r0 = fma(b0, b0, r0); r1 = fma(b1, b1, r1);
            }
            acc[32 * 0] += g00; acc[32 * 1] += g01; acc[32 * 2] += g10; acc[32 * 3] += g11;
            if (ACCR) { accr[0] += e0; accr[1] += e1; accr[2] += r0; accr[3] += r1; }
            else { acc[32 * 4] += e0; acc[32 * 5] += e1; acc[32 * 6] += r0; acc[32 * 7] += r1; }
          }
        };
        if (plain) emit(std::true_type{}); else emit(std::false_type{});
      };
      using Q0 = std::integral_constant<int, 0>;
      using Q1 = std::integral_constant<int, C - 1>;
      if constexpr (C == 2 && DUAL) {
        float2 vb[8];
        masked(v, Q0{});
        masked(vb, Q1{});
        if (it + 1 < n_it) load_masks(ta + 2);                      // next iteration's masks, a transform ahead
        wfft256x2<true>(v, vb, t1, t2, ex, lane);
        finish2(v, vb);
      } else {
        masked(v, Q0{});
        if (C == 1 && it + 1 < n_it) load_masks(ta + 2);
        wfft256<true>(v, t1, t2, ex, lane);
        finish(v, Q0{});
        if constexpr (C == 2) {
          masked(v, Q1{});
          if (it + 1 < n_it) load_masks(ta + 2);
          wfft256<true>(v, t1, t2, ex, lane);
          finish(v, Q1{});
        }
      }
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
      for (int i = 0; i < G::NACC; ++i) vals[C * C + i] = acc[32 * i];
      if constexpr (C == 2 && DUAL && ACCR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) vals[C * C + 4 + i] += accr[i];
      }
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

template <int C, int R, bool SCORE, int W, int CPS, bool DUAL>
static int launch_wstrip_w(const sep_plan *p, FusedArgs a, int batch, double *d_scores, double *d_sums,
                           Scratch &s, cudaStream_t stream) {
  using G = WStripGeom<C, R, SCORE, W, DUAL>;
  const int sms = p->sm_count > 0 ? p->sm_count : 148;
  // Strips are planned for ONE wave of W-warp CTAs (sms * W warps), not for every resident warp: in a stream of
  // independent steps (several launches in flight) fewer, longer strips mean fewer recomputed halo frames and
  // strip prologues, and the other launches fill the rest of the SM (measured at cfg2: 23.9 us per step against
  // 24.8 / 25.5 us when planned for two / three waves; a launch alone takes 50 instead of 40 us).  Large batches
  // get two waves.  SEPCORE_WSTRIP_WARPS overrides the number of warps planned for.
  static const int plan_env = getenv("SEPCORE_WSTRIP_WARPS") ? atoi(getenv("SEPCORE_WSTRIP_WARPS")) : 0;
  const int plan_warps = plan_env > 0 ? plan_env : (batch >= 2 * sms ? 2 * sms * W : sms * W);
  pick_strips(a.T, G::H, 2, batch, plan_warps, &a.tiles, &a.strip_iters);
  int rc;
  double *partials = nullptr;
  int *counters = nullptr;
  // the last strip (tile) of an utterance finalises it inside the kernel: one launch per step instead of three
  // (cfg2: 23.9 -> 22.9 us per step); SEPCORE_SINGLE_LAUNCH=0 brings the separate finalisation kernels back
  static const bool single_launch = !(getenv("SEPCORE_SINGLE_LAUNCH") && atoi(getenv("SEPCORE_SINGLE_LAUNCH")) == 0);
  if (SCORE) {
    // counters first: with a caller workspace they sit at its start, which the caller
    // zero-filled once and every launch leaves at zero
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
  a.vec_ok = (a.n % 4 == 0) && aligned(a.mix) && (!a.refs || aligned(a.refs)) ? 1 : 0;
  const size_t smem = G::smem();
  SEP_CUDA(cudaFuncSetAttribute(wstrip256_kernel<C, R, SCORE, W, CPS, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t total = static_cast<int64_t>(batch) * a.tiles;
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(sms) * CPS, (total + W - 1) / W));
  profile_begin(stream, "wstrip256_kernel<C=%d,R=%d,SCORE=%d,W=%d,CTAS_PER_SM=%d>", C, R, int(SCORE), W, CPS);
  wstrip256_kernel<C, R, SCORE, W, CPS, DUAL><<<grid, W * 32, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (SCORE && counters == nullptr) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

template <int R>
static int dispatch_wstrip(const sep_plan *p, const FusedArgs &a, int batch, double *d_scores,
                           double *d_sums, Scratch &s, cudaStream_t stream) {
  // three 4-warp CTAs per SM: 12 warps, 168 registers per thread, no spills (four CTAs / 128 registers were
  // measured slower: the compiler rematerialises addresses and shuffles register pairs)
  return a.refs != nullptr ? launch_wstrip_w<2, R, true, 4, 3, true>(p, a, batch, d_scores, d_sums, s, stream)
                           : launch_wstrip_w<2, R, false, 4, 3, true>(p, a, batch, d_scores, d_sums, s, stream);
}

int fused_wstrip_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                     double *d_sums, Scratch &s, cudaStream_t stream, bool *handled) {
  *handled = false;
  // two sources only: with one source there is no second transform to run in lockstep and the half-warp strips
  // of fused_strip.cu are faster (14.5 against 16.0 us per step at 64 x 4 s)
  if (p->size != 256 || (p->hops != 2 && p->hops != 4) || C != 2 || a.T < 4) return SEP_OK;
  if (a.n > (int64_t(1) << 30)) return SEP_OK;                    // 32-bit sample indices inside the kernel
  if (getenv("SEPCORE_FORCE_GENERIC") || getenv("SEPCORE_FORCE_TILES") || getenv("SEPCORE_FORCE_HALFWARP")) return SEP_OK;
  *handled = true;
  return p->hops == 2 ? dispatch_wstrip<2>(p, a, batch, d_scores, d_sums, s, stream)
                      : dispatch_wstrip<4>(p, a, batch, d_scores, d_sums, s, stream);
}

}  // namespace sep
