// fused_strip.cu -- the hot path for size = 256 (shift 128 or 64), 1..2 sources:
// warp-autonomous sliding strips, one persistent CTA per SM.
//
// Why: the tile kernel (fused_fast.cu) is bound by instruction issue, not by bytes
// (profiles/r1_ncu_fused256_v2.md): every 16-frame tile stages through CTA-wide
// barriers, round-trips the overlap-add through a shared accumulator, recomputes a
// halo and pays a block reduction, and 1088 short-lived CTAs cost ~15 us before they
// do anything (profiles/r1_phase_costs.md).  Here
//   * a WARP owns a strip of consecutive frames of one utterance and slides over it
//     four frames at a time (two half-warps x one frame pair each, fft256.cuh);
//     nothing but __syncwarp separates its steps, so the warps of an SM drift apart
//     and the load / transform / store phases of different warps overlap;
//   * the overlap-add happens in registers: the two frames of a pair overlap inside
//     the lane that holds them (frame t+1 sample p+16m is frame t sample
//     p+16(m+16/R)), and only the tail that the NEXT frame pair needs crosses to the
//     neighbouring half-warp / the next iteration through a 2-deep shared slot;
//     finished hop blocks go from registers straight to HBM (64 B per half-warp);
//   * waveform spans are double-buffered with zero-filling cp.async one iteration
//     ahead, masks are requested one iteration ahead into the registers they free;
//   * PIT pair sums and the Gram statistics stay in registers for the whole strip:
//     one warp reduction and one partial row per strip (not per tile);
//   * tables are loaded once per SM; the grid is one CTA per SM.
// A strip recomputes R-1 halo frames at its start (atomic-free: every output sample
// is written by exactly one lane, once).  Reference lines: see fused.cu.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fft256.cuh"
#include "fused.cuh"
#include "strip_common.cuh"

namespace sep {

// Warps per CTA (= per SM) is the template parameter W; 8 warps leave each thread 255 registers.

template <int C, int R, bool SCORE, int W>
struct StripGeom {
  static constexpr int SHIFT = 256 / R, H = R - 1;
  static constexpr int NSIG = SCORE ? 1 + C : 1;
  static constexpr int SPAN = 3 * SHIFT + 256;     // samples under the four frames of an iteration
  static constexpr int D = 16 / R;                 // register distance between the frames of a pair
  static constexpr int NY = 16 + D;                // pair-summed values per lane
  static constexpr int NOUT = 32 / R;              // finished values per lane and iteration (two hop blocks)
  static constexpr int NSLOT = NY - NOUT;          // values handed to later frame pairs
  // Shared layout of a staged span: sample i lives at i + 16 * (i / (2 * SHIFT)).  The two
  // half-warps read 2 * SHIFT samples apart (a multiple of 32 banks); the skew moves them
  // 16 banks apart.
  __host__ __device__ static constexpr int sk(int i) { return i + 16 * (i / (2 * SHIFT)); }
  static constexpr int SPAN_SK = (sk(SPAN - 1) + 1 + 3) & ~3;
  static constexpr int UNIT_STRIDE = C * NSLOT * 16 + 16;      // == 16 (mod 32): half-warps on disjoint banks
  // every span is single-buffered but requested as early as its buffer is free: the next mixture span right
  // after this one has been read into registers, the reference spans at the top of their iteration (needed
  // only after the mixture transform)
  static constexpr int STAGE_FLOATS = NSIG * SPAN_SK;
  // tail slots: [unit 0] and [unit 1][parity] for R = 2 (only unit 1's tail crosses an iteration);
  // [parity][unit] for R = 4 (a frame pair also needs its own tail of the previous iteration)
  static constexpr int NSLOTBUF = (R == 2) ? 3 : 4;
  static constexpr int WARP_BYTES = 4 * STAGE_FLOATS + 8 * 2 * kXchFloat2V + 4 * NSLOTBUF * UNIT_STRIDE;
  static constexpr int TABLE_BYTES = 4 * 2 * kWT + 8 * 16 * kXchPitchV;
  static constexpr size_t smem() { return TABLE_BYTES + static_cast<size_t>(W) * WARP_BYTES; }
  static constexpr int NV = FusedVals<C>::NV;
};

template <int C, int R, bool SCORE, int W, int CPS>
__global__ void __launch_bounds__(W * 32, CPS) strip256_kernel(const FusedArgs a) {
  using G = StripGeom<C, R, SCORE, W>;
  constexpr int SHIFT = G::SHIFT, H = G::H, NSIG = G::NSIG, D = G::D, NY = G::NY, NOUT = G::NOUT;
  constexpr int NSLOT = G::NSLOT, SPAN = G::SPAN, SSK = G::SPAN_SK, US = G::UNIT_STRIDE, NV = G::NV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *win = reinterpret_cast<float *>(smem_raw);               // [16][18] 0.5 * analysis
  float *syn = win + kWT;                                         // [16][18] synthesis
  float2 *twt = reinterpret_cast<float2 *>(syn + kWT);            // [16 lanes][18] W256^(lane j)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, unit = lane >> 4, l16 = lane & 15;
  unsigned char *wbase = smem_raw + G::TABLE_BYTES + warp * G::WARP_BYTES;
  float *stage = reinterpret_cast<float *>(wbase);                // [2][NSIG][SSK], skewed
  float2 *xch = reinterpret_cast<float2 *>(stage + G::STAGE_FLOATS) + unit * kXchFloat2V;
  float *slot = reinterpret_cast<float *>(reinterpret_cast<float2 *>(stage + G::STAGE_FLOATS) + 2 * kXchFloat2V);

  if (threadIdx.x < kWT / 4) {
    cp_async16(win + 4 * threadIdx.x, a.win_t + 4 * threadIdx.x);
    cp_async16(syn + 4 * threadIdx.x, a.syn_t + 4 * threadIdx.x);
  }
  for (int i = threadIdx.x; i < 256; i += W * 32) twt[(i >> 4) * kXchPitchV + (i & 15)] = a.tw16[i];
  cp_async_wait_all();
  __syncthreads();                                                // the only block-wide barrier

  const float2 *twl = twt + l16 * kXchPitchV;
  const float2 *winp = reinterpret_cast<const float2 *>(win + 18 * l16);
  const float2 *synp = reinterpret_cast<const float2 *>(syn + 18 * l16);
  const int T = a.T, S = a.tiles, I = a.strip_iters;
  const int total = a.batch * S;
  const int ubase = unit * (2 * SHIFT + 16) + l16;                // skewed position of this lane's first sample
#define SEP_POS(mm) (16 * (mm) + 16 * ((mm) / NOUT))
  const int col8 = l16 == 0 ? 128 : 0;                            // bin 128 lives on lane 0 only
  const float bw8 = l16 == 0 ? 1.f : 0.f;

  for (int strip = warp * gridDim.x + blockIdx.x; strip < total; strip += gridDim.x * W) {
    const int b = strip / S, s = strip - b * S;
    const int q0 = I / S, rem = I - q0 * S;
    const int n_it = q0 + (s < rem ? 1 : 0);
    const int a0 = 4 * (s * q0 + min(s, rem)) - H * s;            // first frame transformed
    const int own_frame0 = s == 0 ? 0 : a0 + H;                   // frames counted by this strip (PIT)
    const int own_block0 = a0 + H;                                // hop blocks written by this strip
    const float *mix_row = a.mix + static_cast<int64_t>(b) * a.n;
    const float *ref_row = SCORE ? a.refs + static_cast<int64_t>(b) * C * a.n : nullptr;
    const float *mask_b = a.masks + static_cast<int64_t>(b) * C * T * 129;
    const int64_t mask_q = static_cast<int64_t>(T) * 129;
    const int64_t n_valid = (SCORE && a.valid) ? min(static_cast<int64_t>(a.valid[b]), a.n) : a.n;
    int len_i = T;
    if (SCORE && a.lengths) len_i = static_cast<int>(a.lengths[b]);

    // sig_lo..sig_hi of iteration `it` -> shared memory (zero-filled outside [0, n)); one commit group
    auto issue_span = [&](int it, int sig_lo, int sig_hi) {
      const int64_t g0 = static_cast<int64_t>(a0 + 4 * it) * SHIFT - a.pad;
      if (a.vec_ok && g0 >= 0 && g0 + SPAN <= a.n) {                // interior span: plain 16-byte copies
#pragma unroll
        for (int sg = 0; sg < NSIG; ++sg) {
          if (sg < sig_lo || sg >= sig_hi) continue;
          float *dst = stage + sg * SSK + 4 * lane;
          const float *src = (sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n) + g0 + 4 * lane;
#pragma unroll
          for (int c0 = 0; c0 < SPAN / 4; c0 += 32)
            if (c0 + 32 <= SPAN / 4 || lane < SPAN / 4 - c0)
              cp_async16(dst + 4 * c0 + 16 * ((4 * c0) / (2 * SHIFT)), src + 4 * c0);
        }
      } else {
#pragma unroll 1
        for (int sg = sig_lo; sg < sig_hi; ++sg) {
          const float *row = sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n;
          float *dst = stage + sg * SSK;
          if (a.vec_ok) {
            for (int c = lane; c < SPAN / 4; c += 32) {
              const int64_t g = g0 + 4 * c;
              const int64_t left = (a.n - g) * 4;
              const int bytes = g < 0 ? 0 : static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(16), left)));
              cp_async16_zfill(dst + G::sk(4 * c), bytes > 0 ? row + g : row, bytes);
            }
          } else {
            for (int i = lane; i < SPAN; i += 32) {
              const int64_t g = g0 + i;
              const bool ok = g >= 0 && g < a.n;
              cp_async4_zfill(dst + G::sk(i), ok ? row + g : row, ok ? 4 : 0);
            }
          }
        }
      }
      cp_async_commit();
    };

    float2 mab[C][9];                                             // (mask of frame ta, mask of frame tb)
    auto load_masks = [&](int ta) {
      // rows beyond T - 1 multiply all-zero spectra: any valid row will do (no predicates)
      const float *pa = mask_b + static_cast<int64_t>(min(ta, T - 1)) * 129 + l16;
      const float *pb = mask_b + static_cast<int64_t>(min(ta + 1, T - 1)) * 129 + l16;
#pragma unroll
      for (int q = 0; q < C; ++q) {
#pragma unroll
        for (int r = 0; r < 8; ++r) mab[q][r] = make_float2(__ldg(pa + 16 * r), __ldg(pb + 16 * r));
        mab[q][8] = make_float2(__ldg(pa - l16 + col8), __ldg(pb - l16 + col8));
        pa += mask_q;
        pb += mask_q;
      }
    };

    __syncwarp();                                                 // previous strip's reads are done
    for (int i = lane; i < G::NSLOTBUF * US; i += 32) slot[i] = 0.f;
    issue_span(0, 0, 1);
    load_masks(a0 + 2 * unit);

    float2 pit2[C * C];                                           // (frame ta part, frame tb part)
    double gram[C * C], ee[C], er[C];
#pragma unroll
    for (int i = 0; i < C * C; ++i) { pit2[i] = make_float2(0.f, 0.f); gram[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < C; ++i) { ee[i] = 0.0; er[i] = 0.0; }

#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int par = it & 1;
      const int ta = a0 + 4 * it + 2 * unit, tb = ta + 1;           // this half-warp's frame pair
      __syncwarp();                                                 // reads of the buffer refilled below are done
      if (NSIG > 1) issue_span(it, 1, NSIG);                        // references of this iteration
      if (NSIG > 1) cp_async_wait<1>(); else cp_async_wait<0>();    // this iteration's mixture (requested an iteration ago) has landed
      __syncwarp();
      const float *st = stage + ubase;                              // signal sg at st + sig_off(sg)

      float2 v[16];
      float2 XR[9], XI[9];                                          // mixture spectra, planar: (X_ta, X_tb)
      float2 inv[9], mag[9];                                        // 1/|X|, gated |X|
      float pmin = 1.f;
      const float2 own2 = make_float2((ta >= own_frame0 && ta < T) ? 1.f : 0.f,
                                      (tb >= own_frame0 && tb < T) ? 1.f : 0.f);
      const float2 gate2 = make_float2(ta < len_i ? 1.f : 0.f, tb < len_i ? 1.f : 0.f);
#pragma unroll 1
      for (int sg = 0; sg < NSIG; ++sg) {
        const float *sp = st + sg * SSK;
        if (sg == 1) {                                              // the references have landed
          cp_async_wait<1>();
          __syncwarp();
        }
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          const float2 w = winp[m / 2];
          v[m] = cscale(make_float2(sp[SEP_POS(m)], sp[SEP_POS(m + D)]), w.x);
          v[m + 1] = cscale(make_float2(sp[SEP_POS(m + 1)], sp[SEP_POS(m + 1 + D)]), w.y);
        }
        if (sg == 0) {                                              // the mixture span is in registers: refill it
          __syncwarp();
          if (it + 1 < n_it) issue_span(it + 1, 0, 1); else cp_async_commit();
        }
        fft256v<false>(v, twl, xch, l16);
        if (sg == 0) {
#pragma unroll
          for (int r = 0; r < 9; ++r) {
            split_planar(v, l16, r, XR[r], XI[r]);
            if (SCORE) {
              const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
              pmin = fminf(pmin, fminf(p.x, p.y));
              // |X| = 0 -> a finite 1/|X| (and |X| * 1/|X| = 0); the exact-zero label is redone below
              inv[r] = make_float2(rsqrt_fast(fmaxf(p.x, 1e-36f)), rsqrt_fast(fmaxf(p.y, 1e-36f)));
              mag[r] = __fmul2_rn(__fmul2_rn(p, inv[r]), gate2);
            }
          }
        } else {
          // label = |S| cos(angle X - angle S) = Re(S conj X) / |X| ; angle(0) = 0 -> Re S
          float2 pj[C];
#pragma unroll
          for (int i = 0; i < C; ++i) pj[i] = make_float2(0.f, 0.f);
          const bool anyzero = __any_sync(0xffffffffu, !(pmin > 0.f));
          // one straight-line bin loop per case (ordinary / some mixture bin exactly zero)
          auto bins = [&](auto zero_tag) {
            constexpr bool ZERO = decltype(zero_tag)::value;
#pragma unroll
            for (int r = 0; r < 9; ++r) {
              float2 SR, SI;
              split_planar(v, l16, r, SR, SI);
              float2 l = __fmul2_rn(__ffma2_rn(SR, XR[r], __fmul2_rn(SI, XI[r])), inv[r]);
              if (ZERO) {
                const float2 p = __ffma2_rn(XR[r], XR[r], __fmul2_rn(XI[r], XI[r]));
                if (!(p.x > 0.f)) l.x = SR.x;
                if (!(p.y > 0.f)) l.y = SR.y;
              }
#pragma unroll
              for (int i = 0; i < C; ++i) {
                float2 d = __ffma2_rn(mab[i][r], mag[r], make_float2(-l.x, -l.y));
                if (r == 8) d = __fmul2_rn(d, make_float2(bw8, bw8));
                pj[i] = __ffma2_rn(d, d, pj[i]);
              }
            }
          };
          if (!anyzero) bins(std::false_type{}); else bins(std::true_type{});
          // fold into column j = sg - 1 without indexing registers by a runtime value
#pragma unroll
          for (int jj = 0; jj < C; ++jj) {
            const float2 wj = (jj == sg - 1) ? own2 : make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < C; ++i) pit2[i * C + jj] = __ffma2_rn(pj[i], wj, pit2[i * C + jj]);
          }
        }
      }

      // ---- masked spectra -> time frames -> in-register overlap-add -> HBM ----
      const int64_t gb = static_cast<int64_t>(ta) * SHIFT - a.pad + l16;      // sample index of y[0]
      // leading values that belong to the previous strip (or to the fade-in padding)
      const int skip = min(NOUT, max(0, own_block0 - ta) * (NOUT / 2));
      const int64_t room = a.n - gb, room_v = n_valid - gb;
      const int lim = room <= 0 ? 0 : static_cast<int>(min(static_cast<int64_t>(NOUT), (room + 15) >> 4));
      const int lim_v = room_v <= 0 ? 0 : static_cast<int>(min(static_cast<int64_t>(NOUT), (room_v + 15) >> 4));
      // one path per warp: if either half-warp has an edge, both take the predicated path (instead of the
      // two paths running one after the other)
      const bool plain = !__any_sync(0xffffffffu, !(skip == 0 && lim_v == NOUT));
#pragma unroll
      for (int q = 0; q < C; ++q) {
        {
          float2 L[9], Mi[9];
#pragma unroll
          for (int r = 0; r < 9; ++r) {
            float2 mk = mab[0][r];
#pragma unroll
            for (int qq = 1; qq < C; ++qq)
              if (q == qq) mk = mab[qq][r];
            // (P, Q) = (m_a X_a, m_b X_b):  mR = (Re P, Re Q), mI = (Im P, Im Q)
            const float2 mR = __fmul2_rn(XR[r], mk), mI = __fmul2_rn(XI[r], mk);
            L[r] = __fadd2_rn(mR, make_float2(-mI.y, mI.x));         // P + i Q
            Mi[r] = __fadd2_rn(mR, make_float2(mI.y, -mI.x));        // conj P + i conj Q (for the mirror bin)
          }
          merge_pair(v, l16, L, Mi);
        }
        if (q == C - 1 && it + 1 < n_it) load_masks(ta + 4);       // next iteration's masks, a transform ahead
        fft256v<true>(v, twl, xch, l16);
        float y[NY];
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          const float2 w = synp[m / 2];
          v[m] = cscale(v[m], w.x);
          v[m + 1] = cscale(v[m + 1], w.y);
        }
#pragma unroll
        for (int mm = 0; mm < NY; ++mm) {
          if (mm < D) y[mm] = v[mm].x;
          else if (mm < 16) y[mm] = v[mm].x + v[mm - D].y;
          else y[mm] = v[mm - D].y;
        }
        // hand the tail to the frame pairs that follow; take what the previous ones left
        // slot buffers: R = 2: [unit 1 even], [unit 0], [unit 1 odd] (neighbours 16 banks apart);  R = 4: [parity][unit]
        const int s_mine = R == 2 ? (unit ? 2 * par : 1) : par * 2 + unit;
        const int s_prev = R == 2 ? (unit ? 1 : 2 * (par ^ 1)) : (unit ? par : par ^ 1) * 2 + (unit ^ 1);
        float *mine = slot + s_mine * US + q * NSLOT * 16 + l16;
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) mine[16 * k] = y[NOUT + k];
        __syncwarp();
        const float *prev = slot + s_prev * US + q * NSLOT * 16 + l16;
#pragma unroll
        for (int mm = 0; mm < (NSLOT < NOUT ? NSLOT : NOUT); ++mm) y[mm] += prev[16 * mm];
        if (NSLOT > NOUT) {
          const float *pp = slot + ((par ^ 1) * 2 + unit) * US + q * NSLOT * 16 + l16;
#pragma unroll
          for (int mm = 0; mm < NSLOT - NOUT; ++mm) y[mm] += pp[16 * (NOUT + mm)];
        }
        float *out = a.est ? a.est + (static_cast<int64_t>(b) * C + q) * a.n + gb : nullptr;
        double gq[C], eq = 0.0, rq[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { gq[j] = 0.0; rq[j] = 0.0; }
        if (plain) {
          if (out) {
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) out[16 * mm] = y[mm];
          }
          if (SCORE) {
#pragma unroll
            for (int mm = 0; mm < NOUT; ++mm) {
              const double e = static_cast<double>(y[mm]);
              eq = fma(e, e, eq);
#pragma unroll
              for (int j = 0; j < C; ++j) {
                const double r = static_cast<double>(st[(1 + j) * SSK + SEP_POS(mm)]);
                gq[j] = fma(e, r, gq[j]);
                if (q == 0) rq[j] = fma(r, r, rq[j]);
              }
            }
          }
        } else {
#pragma unroll
          for (int mm = 0; mm < NOUT; ++mm) {
            const bool okw = mm >= skip && mm < lim;
            if (out && okw) out[16 * mm] = y[mm];
            if (SCORE) {
              const bool okv = mm >= skip && mm < lim_v;
              const double e = okv ? static_cast<double>(y[mm]) : 0.0;
              eq = fma(e, e, eq);
#pragma unroll
              for (int j = 0; j < C; ++j) {
                const double r = okv ? static_cast<double>(st[(1 + j) * SSK + SEP_POS(mm)]) : 0.0;
                gq[j] = fma(e, r, gq[j]);
                if (q == 0) rq[j] = fma(r, r, rq[j]);
              }
            }
          }
        }
        if (SCORE) {
#pragma unroll
          for (int qq = 0; qq < C; ++qq)
            if (qq == q) {
              ee[qq] += eq;
#pragma unroll
              for (int j = 0; j < C; ++j) gram[qq * C + j] += gq[j];
            }
          if (q == 0) {
#pragma unroll
            for (int j = 0; j < C; ++j) er[j] += rq[j];
          }
        }
      }
    }
    cp_async_wait<0>();

    if (SCORE) {
      double vals[NV];
#pragma unroll
      for (int i = 0; i < C * C; ++i) {
        vals[i] = static_cast<double>(pit2[i].x) + static_cast<double>(pit2[i].y);
        vals[C * C + i] = gram[i];
      }
#pragma unroll
      for (int i = 0; i < C; ++i) { vals[2 * C * C + i] = ee[i]; vals[2 * C * C + C + i] = er[i]; }
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
#undef SEP_POS
}

template <int C, int R, bool SCORE, int W, int CTAS_PER_SM>
static int launch_strip_w(const sep_plan *p, FusedArgs a, int batch, double *d_scores, double *d_sums,
                          Scratch &s, cudaStream_t stream) {
  using G = StripGeom<C, R, SCORE, W>;
  const int sms = p->sm_count > 0 ? p->sm_count : 148;
  static const int plan_env = getenv("SEPCORE_STRIP_WARPS") ? atoi(getenv("SEPCORE_STRIP_WARPS")) : 0;
  // planned for one wave of W-warp CTAs at small batches: fewer, longer strips; in a stream of independent steps the
  // other launches fill the SMs (one source, 64 x 4 s: 14.3 -> 13.2 us per step; a launch alone 29 -> 34 us); see
  // fused_wstrip.cu
  pick_strips(a.T, G::H, 4, batch, plan_env > 0 ? plan_env : (batch >= 2 * sms ? sms * W * CTAS_PER_SM : sms * W),
              &a.tiles, &a.strip_iters);
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
  SEP_CUDA(cudaFuncSetAttribute(strip256_kernel<C, R, SCORE, W, CTAS_PER_SM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t total = static_cast<int64_t>(batch) * a.tiles;
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(sms) * CTAS_PER_SM, (total + W - 1) / W));
  profile_begin(stream, "strip256_kernel<C=%d,R=%d,SCORE=%d,W=%d>", C, R, int(SCORE), W);
  strip256_kernel<C, R, SCORE, W, CTAS_PER_SM><<<grid, W * 32, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (SCORE && counters == nullptr) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

template <int C, int R, bool SCORE>
static int launch_strip(const sep_plan *p, const FusedArgs &a, int batch, double *d_scores, double *d_sums,
                        Scratch &s, cudaStream_t stream) {
  // 12 warps (168 registers) were measured slower: the spills miss the small L1 left beside 200 KB of shared memory
  // two 4-warp CTAs per SM (104 KB each) instead of one 8-warp CTA: CTAs of the NEXT launch (another stream
  // of the replayed graph) move in as soon as four warps are done, not eight
  // (three CTAs per SM -- 168 registers, 200 bytes of spills per thread -- were measured 40 % slower; 10 or 11
  // warps per SM do not help either: warps are allocated in fours, so ptxas is held to 168 registers all the same)
  return launch_strip_w<C, R, SCORE, 4, 2>(p, a, batch, d_scores, d_sums, s, stream);
}

template <int C>
static int dispatch_strip(const sep_plan *p, const FusedArgs &a, int batch, double *d_scores,
                          double *d_sums, Scratch &s, cudaStream_t stream) {
  const bool score = a.refs != nullptr;
  if (p->hops == 2)
    return score ? launch_strip<C, 2, true>(p, a, batch, d_scores, d_sums, s, stream)
                 : launch_strip<C, 2, false>(p, a, batch, d_scores, d_sums, s, stream);
  return score ? launch_strip<C, 4, true>(p, a, batch, d_scores, d_sums, s, stream)
               : launch_strip<C, 4, false>(p, a, batch, d_scores, d_sums, s, stream);
}

int fused_strip_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                    double *d_sums, Scratch &s, cudaStream_t stream, bool *handled) {
  *handled = false;
  if (p->size != 256 || (p->hops != 2 && p->hops != 4) || C > 2 || a.T < 4) return SEP_OK;
  if (getenv("SEPCORE_FORCE_GENERIC") || getenv("SEPCORE_FORCE_TILES")) return SEP_OK;
  *handled = true;
  return C == 1 ? dispatch_strip<1>(p, a, batch, d_scores, d_sums, s, stream)
                : dispatch_strip<2>(p, a, batch, d_scores, d_sums, s, stream);
}

}  // namespace sep
