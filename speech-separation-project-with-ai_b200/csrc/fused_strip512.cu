// fused_strip512.cu -- the fused hot path for size = 512, shift = 128 (BASELINE config 4:
// 3-speaker uPIT, 6 permutations), 1..3 sources: warp-autonomous sliding strips like
// fused_strip.cu, with the 512-point REAL transform of a frame done as ONE 256-point
// complex transform on a half-warp (fft256.cuh) of z[n] = x[2n] + i x[2n+1]:
//   forward   X[k] = E[k] - G[k],  E = Z[k] + conj Z[256-k],  G = i W512^k (Z[k] - conj Z[256-k])
//             (the 1/2 rides in the window table);  X[256] = E[0] + G[0] on lane 0
//   inverse   Z'[k] = (Y[k] + conj Y[256-k]) + i conj(W512^k) (Y[k] - conj Y[256-k]),
//             y[2n] + i y[2n+1] = IDFT256(Z')   (unnormalised; 1/size rides in the synthesis window)
// A lane holds bins l16 + 16 r (r = 0..15; lane 0 also bin 256) and sample pairs
// (2p + 32m, 2p + 32m + 1): every shared/global access of a half-warp is one contiguous
// 128-byte run of float2.
//
// A warp slides over its strip two frames at a time (one frame per half-warp).  The
// overlap-add (4 frames per hop block) is a running accumulator of the three unfinished
// hop blocks, handed from frame to frame through shared memory; a finished block goes from
// registers to HBM.  All staging is single-buffered but requested as soon as a buffer is
// free: the next mixture span right after this one was read, the next masks after the last
// mask multiply, the next references after the Gram statistics.
// Reference lines: see fused.cu.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "fft256.cuh"
#include "fused.cuh"
#include "strip_common.cuh"

namespace sep {

template <int C, bool SCORE, int W>
struct Strip512Geom {
  static constexpr int SHIFT = 128, H = 3, FPI = 2, BINS = 257;
  static constexpr int NSIG = SCORE ? 1 + C : 1;
  static constexpr int SPAN = SHIFT * (FPI - 1) + 512;          // samples under the two frames of an iteration
  static constexpr int STAGE_FLOATS = NSIG * SPAN;
  static constexpr int MASK_ROW = 260;                          // 257 used
  // frames of an iteration 16 banks apart
  static constexpr int MASK_FRAME = ((C * MASK_ROW + 15) / 32) * 32 + 16;
  static constexpr int MASK_FLOATS = FPI * MASK_FRAME;
  static constexpr int ACC_FLOAT2 = C * 3 * 64;                 // three unfinished hop blocks per source
  static constexpr int WARP_BYTES = 4 * (STAGE_FLOATS + MASK_FLOATS) + 8 * ACC_FLOAT2 + 8 * 2 * kXchFloat2V;
  static constexpr int TABLE_BYTES = 8 * 4 * 16 * kXchPitchV;   // win2, syn2, tw256 rows, tw512 rows
  static constexpr size_t smem() { return TABLE_BYTES + static_cast<size_t>(W) * WARP_BYTES; }
  static constexpr int NV = FusedVals<C>::NV;
};

template <int C, bool SCORE, int W>
__global__ void __launch_bounds__(W * 32, 8 / W) strip512_kernel(const FusedArgs a) {
  using G = Strip512Geom<C, SCORE, W>;
  constexpr int SHIFT = G::SHIFT, H = G::H, NSIG = G::NSIG, SPAN = G::SPAN, NV = G::NV;
  constexpr int MROW = G::MASK_ROW, MFRAME = G::MASK_FRAME, BINS = G::BINS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *win2 = reinterpret_cast<float2 *>(smem_raw);            // [16][18]
  float2 *syn2 = win2 + 16 * kXchPitchV;
  float2 *tw256 = syn2 + 16 * kXchPitchV;                         // [16 lanes][18] W256^(lane j)
  float2 *tw512 = tw256 + 16 * kXchPitchV;                        // [16 lanes][18] W512^(lane + 16 r)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, unit = lane >> 4, l16 = lane & 15;
  unsigned char *wbase = smem_raw + G::TABLE_BYTES + warp * G::WARP_BYTES;
  float *stage = reinterpret_cast<float *>(wbase);                // [NSIG][SPAN]
  float *msk = stage + G::STAGE_FLOATS;                           // [2 frames][C][MROW]
  float2 *accb = reinterpret_cast<float2 *>(msk + G::MASK_FLOATS);   // [C][3][64] running overlap-add
  float2 *xch = accb + G::ACC_FLOAT2 + unit * kXchFloat2V;

  for (int i = threadIdx.x; i < 16 * kXchPitchV; i += W * 32) {
    win2[i] = a.win2_t[i];
    syn2[i] = a.syn2_t[i];
    tw512[i] = a.tw512_t[i];
  }
  for (int i = threadIdx.x; i < 256; i += W * 32) tw256[(i >> 4) * kXchPitchV + (i & 15)] = a.tw16[i];
  __syncthreads();                                                // the only block-wide barrier

  const float2 *twl = tw256 + l16 * kXchPitchV;
  const float4 *tw5 = reinterpret_cast<const float4 *>(tw512 + l16 * kXchPitchV);
  const float4 *winp = reinterpret_cast<const float4 *>(win2 + l16 * kXchPitchV);
  const float4 *synp = reinterpret_cast<const float4 *>(syn2 + l16 * kXchPitchV);
  const int T = a.T, S = a.tiles, I = a.strip_iters;
  const int total = a.batch * S;
  const int src_lane = (16 - l16) & 15;
  const float bw16 = l16 == 0 ? 1.f : 0.f;                        // bin 256 lives on lane 0 only

  for (int strip = warp * gridDim.x + blockIdx.x; strip < total; strip += gridDim.x * W) {
    const int b = strip / S, s = strip - b * S;
    const int q0 = I / S, rem = I - q0 * S;
    const int n_it = q0 + (s < rem ? 1 : 0);
    const int a0 = 2 * (s * q0 + min(s, rem)) - H * s;            // first frame transformed
    const int own_frame0 = s == 0 ? 0 : a0 + H;                   // frames counted by this strip (PIT)
    const int own_block0 = a0 + H;                                // hop blocks written by this strip
    const float *mix_row = a.mix + static_cast<int64_t>(b) * a.n;
    const float *ref_row = SCORE ? a.refs + static_cast<int64_t>(b) * C * a.n : nullptr;
    const float *mask_b = a.masks + static_cast<int64_t>(b) * C * T * BINS;
    const int64_t n_valid = (SCORE && a.valid) ? min(static_cast<int64_t>(a.valid[b]), a.n) : a.n;
    int len_i = T;
    if (SCORE && a.lengths) len_i = static_cast<int>(a.lengths[b]);

    // signals [sig_lo, sig_hi) of iteration `it` -> shared memory (zeros outside [0, n)); one commit group
    auto issue_span = [&](int it, int sig_lo, int sig_hi) {
      if (it < n_it) {
        const int64_t g0 = static_cast<int64_t>(a0 + 2 * it) * SHIFT - a.pad;
        const bool inside = a.vec_ok && g0 >= 0 && g0 + SPAN <= a.n;
#pragma unroll 1
        for (int sg = sig_lo; sg < sig_hi; ++sg) {
          const float *row = sg == 0 ? mix_row : ref_row + static_cast<int64_t>(sg - 1) * a.n;
          float *dst = stage + sg * SPAN;
          if (inside) {
#pragma unroll
            for (int c0 = 0; c0 < SPAN / 4; c0 += 32) cp_async16(dst + 4 * (c0 + lane), row + g0 + 4 * (c0 + lane));
          } else if (a.vec_ok) {
            for (int c = lane; c < SPAN / 4; c += 32) {
              const int64_t g = g0 + 4 * c;
              const int64_t left = (a.n - g) * 4;
              const int bytes = g < 0 ? 0 : static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(16), left)));
              cp_async16_zfill(dst + 4 * c, bytes > 0 ? row + g : row, bytes);
            }
          } else {
            for (int i = lane; i < SPAN; i += 32) {
              const int64_t g = g0 + i;
              const bool ok = g >= 0 && g < a.n;
              cp_async4_zfill(dst + i, ok ? row + g : row, ok ? 4 : 0);
            }
          }
        }
      }
      cp_async_commit();
    };
    // mask rows of the two frames of iteration `it` (rows beyond T - 1 multiply all-zero spectra: clamp)
    auto issue_masks = [&](int it) {
      if (it < n_it) {
#pragma unroll 1
        for (int f = 0; f < 2; ++f) {
          const int t = min(a0 + 2 * it + f, T - 1);
#pragma unroll 1
          for (int q = 0; q < C; ++q) {
            const float *row = mask_b + (static_cast<int64_t>(q) * T + t) * BINS;
            float *dst = msk + f * MFRAME + q * MROW;
#pragma unroll
            for (int i0 = 0; i0 < 288; i0 += 32)
              if (i0 + lane < BINS) cp_async4(dst + i0 + lane, row + i0 + lane);
          }
        }
      }
      cp_async_commit();
    };

    __syncwarp();                                                 // previous strip's reads are done
    for (int i = lane; i < G::ACC_FLOAT2; i += 32) accb[i] = make_float2(0.f, 0.f);
    issue_span(0, 0, 1);
    issue_masks(0);
    issue_span(0, 1, NSIG);

    float pit[C * C];
    double gram[C * C], ee[C], er[C];
#pragma unroll
    for (int i = 0; i < C * C; ++i) { pit[i] = 0.f; gram[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < C; ++i) { ee[i] = 0.0; er[i] = 0.0; }

#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
      const int ta = a0 + 2 * it + unit;                            // this half-warp's frame
      // outstanding groups, oldest first: mixture(it), masks(it), references(it)
      cp_async_wait<2>();
      __syncwarp();
      const float2 *st2 = reinterpret_cast<const float2 *>(stage) + unit * (SHIFT / 2) + l16;
      const float *mrow = msk + unit * MFRAME;

      float2 v[16];
      float2 X[17];                                                 // bins l16 + 16 r; X[16] = bin 256 (lane 0)
      float inv[17], mag[17];
      float pmin = 1.f;
      const float own_w = (ta >= own_frame0 && ta < T) ? 1.f : 0.f;
      const float gate = ta < len_i ? 1.f : 0.f;
#pragma unroll 1
      for (int sg = 0; sg < NSIG; ++sg) {
        if (sg == 1) {                                              // masks and references have landed
          cp_async_wait<1>();
          __syncwarp();
        }
        const float2 *sp = st2 + sg * (SPAN / 2);
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          const float4 w = winp[m / 2];
          v[m] = __fmul2_rn(sp[16 * m], make_float2(w.x, w.y));
          v[m + 1] = __fmul2_rn(sp[16 * (m + 1)], make_float2(w.z, w.w));
        }
        if (sg == 0) {
          __syncwarp();                                             // both half-warps have read the mixture span
          issue_span(it + 1, 0, 1);
        }
        fft256v<false>(v, twl, xch, l16);
        float pj[C];
#pragma unroll
        for (int i = 0; i < C; ++i) pj[i] = 0.f;
        const bool anyzero = sg > 0 && __any_sync(0xffffffffu, !(pmin > 0.f));
        // one straight-line body per case (mixture / labels / labels with exactly-zero mixture bins): the case is
        // decided once per transform, not per bin
        auto bins = [&](auto mix_tag, auto zero_tag) {
          constexpr bool MIX = decltype(mix_tag)::value, ZERO = decltype(zero_tag)::value;
#pragma unroll
          for (int r = 0; r < 16; r += 2) {
            const float4 w4 = tw5[r / 2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int rr = r + h;
              const float2 wk = h == 0 ? make_float2(w4.x, w4.y) : make_float2(w4.z, w4.w);
              float2 got;
              got.x = __shfl_sync(0xffffffffu, v[15 - rr].x, src_lane, 16);
              got.y = __shfl_sync(0xffffffffu, v[15 - rr].y, src_lane, 16);
              const float2 own = v[(16 - rr) & 15];
              const float2 zp = l16 == 0 ? own : got;
              const float2 z = v[rr];
              const float2 E = cadd_conj(z, zp);                                        // Z + conj Z'
              const float2 O = __fadd2_rn(z, make_float2(-zp.x, zp.y));                 // Z - conj Z'
              const float2 Gk = cmul(O, make_float2(-wk.y, wk.x));                      // i W^k O
              const float2 Sx = csub(E, Gk);
              if (MIX) {
                X[rr] = Sx;
                if (rr == 0) X[16] = cadd(E, Gk);
                if (SCORE) {
                  const float p = fmaf(Sx.x, Sx.x, Sx.y * Sx.y);
                  pmin = fminf(pmin, p);
                  inv[rr] = rsqrt_fast(fmaxf(p, 1e-36f));
                  mag[rr] = p * inv[rr] * gate;
                  if (rr == 0) {
                    const float p16 = fmaf(X[16].x, X[16].x, X[16].y * X[16].y);
                    pmin = fminf(pmin, l16 == 0 ? p16 : 1.f);
                    inv[16] = rsqrt_fast(fmaxf(p16, 1e-36f));
                    mag[16] = p16 * inv[16] * gate;
                  }
                }
              } else {
                // label = |S| cos(angle X - angle S) = Re(S conj X) / |X| ; angle(0) = 0 -> Re S
                float l = fmaf(Sx.x, X[rr].x, Sx.y * X[rr].y) * inv[rr];
                if (ZERO && !(fmaf(X[rr].x, X[rr].x, X[rr].y * X[rr].y) > 0.f)) l = Sx.x;
#pragma unroll
                for (int i = 0; i < C; ++i) {
                  const float d = fmaf(mrow[i * MROW + l16 + 16 * rr], mag[rr], -l);
                  pj[i] = fmaf(d, d, pj[i]);
                }
                if (rr == 0) {
                  const float2 S16 = cadd(E, Gk);
                  float l16v = fmaf(S16.x, X[16].x, S16.y * X[16].y) * inv[16];
                  if (ZERO && !(fmaf(X[16].x, X[16].x, X[16].y * X[16].y) > 0.f)) l16v = S16.x;
#pragma unroll
                  for (int i = 0; i < C; ++i) {
                    const float d = fmaf(mrow[i * MROW + 256], mag[16], -l16v) * bw16;
                    pj[i] = fmaf(d, d, pj[i]);
                  }
                }
              }
            }
          }
        };
        if (sg == 0) bins(std::true_type{}, std::false_type{});
        else if (!anyzero) bins(std::false_type{}, std::false_type{});
        else bins(std::false_type{}, std::true_type{});
        if (sg > 0) {
          // fold into column j = sg - 1 without indexing registers by a runtime value
#pragma unroll
          for (int jj = 0; jj < C; ++jj) {
            const float wj = (jj == sg - 1) ? own_w : 0.f;
#pragma unroll
            for (int i = 0; i < C; ++i) pit[i * C + jj] = fmaf(pj[i], wj, pit[i * C + jj]);
          }
        }
      }

      if (NSIG == 1) {                                              // no references: the masks have landed
        cp_async_wait<1>();
        __syncwarp();
      }
      // ---- masked spectra -> time frame -> running overlap-add -> HBM ----
      const int64_t gb = static_cast<int64_t>(ta) * SHIFT - a.pad;          // first sample of hop block ta
      const bool owned = ta >= own_block0;
      const bool plain = owned && gb + SHIFT <= n_valid && a.vec_ok;
#pragma unroll 1
      for (int q = 0; q < C; ++q) {
        {
          float2 Y[17];
          const float *mq = mrow + q * MROW + l16;
#pragma unroll
          for (int r = 0; r < 16; ++r) Y[r] = cscale(X[r], mq[16 * r]);
          Y[16] = cscale(X[16], mq[256 - l16]);
          if (q == C - 1) {                                         // the mask rows are free again
            __syncwarp();
            issue_masks(it + 1);
          }
#pragma unroll
          for (int r = 0; r < 16; r += 2) {
            const float4 w4 = tw5[r / 2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int rr = r + h;
              const float2 wk = h == 0 ? make_float2(w4.x, w4.y) : make_float2(w4.z, w4.w);
              float2 got;
              got.x = __shfl_sync(0xffffffffu, Y[15 - rr].x, src_lane, 16);
              got.y = __shfl_sync(0xffffffffu, Y[15 - rr].y, src_lane, 16);
              const float2 own = Y[16 - rr];
              const float2 yp = l16 == 0 ? own : got;
              const float2 ye = cadd_conj(Y[rr], yp);                                    // Y + conj Y'
              const float2 tt = __fadd2_rn(Y[rr], make_float2(-yp.x, yp.y));             // Y - conj Y'
              const float2 yo = cmul_conj(tt, wk);                                       // conj(W^k) (Y - conj Y')
              v[rr] = cadd_pi(ye, yo);                                                   // Ye + i Yo
            }
          }
        }
        fft256v<true>(v, twl, xch, l16);
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          const float4 w = synp[m / 2];
          v[m] = __fmul2_rn(v[m], make_float2(w.x, w.y));
          v[m + 1] = __fmul2_rn(v[m + 1], make_float2(w.z, w.w));
        }
        // running overlap-add: hop block ta is finished by this frame; frames take turns
        float2 *Aq = accb + q * 3 * 64 + l16;
        float2 o[4];
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
          if (unit == step) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              o[m] = __fadd2_rn(Aq[16 * m], v[m]);
              Aq[16 * m] = __fadd2_rn(Aq[64 + 16 * m], v[4 + m]);
              Aq[64 + 16 * m] = __fadd2_rn(Aq[128 + 16 * m], v[8 + m]);
              Aq[128 + 16 * m] = v[12 + m];
            }
          }
          __syncwarp();
        }
        double gq[C], eq = 0.0, rq[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { gq[j] = 0.0; rq[j] = 0.0; }
        float *out = a.est ? a.est + (static_cast<int64_t>(b) * C + q) * a.n + gb + 2 * l16 : nullptr;
        const float2 *rf = st2 + (SPAN / 2);                       // reference 0 at this frame's first hop block
        if (plain) {
          if (out) {
#pragma unroll
            for (int m = 0; m < 4; ++m) *reinterpret_cast<float2 *>(out + 32 * m) = o[m];
          }
          if (SCORE) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const double e0 = static_cast<double>(o[m].x), e1 = static_cast<double>(o[m].y);
              eq = fma(e0, e0, eq);
              eq = fma(e1, e1, eq);
#pragma unroll
              for (int j = 0; j < C; ++j) {
                const float2 r2 = rf[j * (SPAN / 2) + 16 * m];
                const double r0 = static_cast<double>(r2.x), r1 = static_cast<double>(r2.y);
                gq[j] = fma(e0, r0, gq[j]);
                gq[j] = fma(e1, r1, gq[j]);
                if (q == 0) { rq[j] = fma(r0, r0, rq[j]); rq[j] = fma(r1, r1, rq[j]); }
              }
            }
          }
        } else if (owned) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int64_t g = gb + 2 * l16 + 32 * m + h;
              const float ev = h == 0 ? o[m].x : o[m].y;
              if (out && g < a.n) out[32 * m + h] = ev;
              if (SCORE && g < n_valid) {
                const double e = static_cast<double>(ev);
                eq = fma(e, e, eq);
#pragma unroll
                for (int j = 0; j < C; ++j) {
                  const float2 r2 = rf[j * (SPAN / 2) + 16 * m];
                  const double r = static_cast<double>(h == 0 ? r2.x : r2.y);
                  gq[j] = fma(e, r, gq[j]);
                  if (q == 0) rq[j] = fma(r, r, rq[j]);
                }
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
      __syncwarp();                                                 // the reference spans are free again
      issue_span(it + 1, 1, NSIG);
    }
    cp_async_wait<0>();

    if (SCORE) {
      double vals[NV];
#pragma unroll
      for (int i = 0; i < C * C; ++i) { vals[i] = static_cast<double>(pit[i]); vals[C * C + i] = gram[i]; }
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
}

template <int C, bool SCORE>
static int launch_strip512(const sep_plan *p, FusedArgs a, int batch, double *d_scores, double *d_sums,
                           Scratch &s, cudaStream_t stream) {
  // two 4-warp CTAs per SM (<= 112 KB each): CTAs of the next launch move in as soon as four warps are done
  constexpr int W = 4, CTAS_PER_SM = 8 / W;
  using G = Strip512Geom<C, SCORE, W>;
  const int sms = p->sm_count > 0 ? p->sm_count : 148;
  static const int plan_env = getenv("SEPCORE_STRIP_WARPS") ? atoi(getenv("SEPCORE_STRIP_WARPS")) : 0;
  // planned for half of the resident warps at small batches: fewer, longer strips (three halo frames each) --
  // in a stream of independent steps the other launches fill the SMs (cfg4: 172 -> 164 us per step; a launch
  // alone 197 -> 263 us); see fused_wstrip.cu
  pick_strips(a.T, G::H, G::FPI, batch, plan_env > 0 ? plan_env : (batch >= 2 * sms ? sms * 8 : sms * 4), &a.tiles,
              &a.strip_iters);
  int rc;
  double *partials = nullptr;
  int *counters = nullptr;
  // the last strip (tile) of an utterance finalises it inside the kernel: one launch per step instead of three
  // (cfg2: 23.9 -> 22.9 us per step); SEPCORE_SINGLE_LAUNCH=0 brings the separate finalisation kernels back
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
  SEP_CUDA(cudaFuncSetAttribute(strip512_kernel<C, SCORE, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t total = static_cast<int64_t>(batch) * a.tiles;
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(sms) * CTAS_PER_SM, (total + W - 1) / W));
  profile_begin(stream, "strip512_kernel<C=%d,SCORE=%d,W=%d>", C, int(SCORE), W);
  strip512_kernel<C, SCORE, W><<<grid, W * 32, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (SCORE && counters == nullptr) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

int fused_strip512_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                       double *d_sums, Scratch &s, cudaStream_t stream, bool *handled) {
  *handled = false;
  if (p->size != 512 || p->shift != 128 || C > 3 || a.T < 4 || a.win2_t == nullptr) return SEP_OK;
  if (getenv("SEPCORE_FORCE_GENERIC") || getenv("SEPCORE_FORCE_TILES")) return SEP_OK;
  *handled = true;
  const bool score = a.refs != nullptr;
  switch (C) {
    case 1: return score ? launch_strip512<1, true>(p, a, batch, d_scores, d_sums, s, stream)
                         : launch_strip512<1, false>(p, a, batch, d_scores, d_sums, s, stream);
    case 2: return score ? launch_strip512<2, true>(p, a, batch, d_scores, d_sums, s, stream)
                         : launch_strip512<2, false>(p, a, batch, d_scores, d_sums, s, stream);
    default: return score ? launch_strip512<3, true>(p, a, batch, d_scores, d_sums, s, stream)
                          : launch_strip512<3, false>(p, a, batch, d_scores, d_sums, s, stream);
  }
}

}  // namespace sep
