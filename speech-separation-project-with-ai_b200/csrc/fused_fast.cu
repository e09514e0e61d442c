// fused_fast.cu -- register-resident specialisation of the fused hot path for
// size = 256 (shift 128 or 64), 1..4 sources.
//
// Same contract and tile decomposition as fused.cu (owned frames / recomputed
// halo / float64 tile partials), but:
//   * each half-warp transforms a PAIR of consecutive frames as one complex
//     256-point FFT held in registers (fft256.cuh): frame t rides in the real
//     part, frame t+1 in the imaginary part.  A frame pair costs 1 + C forward
//     and C inverse complex transforms (mixture + C references, C estimates).
//   * mask multiply, |X|, PSA labels and the PIT squared differences happen on
//     the split spectra in registers;  spec_c = mask_c * X (cell 41 identity).
//   * overlap-add is atomic-free and needs no frame buffer: windowed time
//     frames are added straight into a shared accumulator in size/shift phases
//     (frames with equal t mod size/shift never overlap), one __syncthreads
//     between phases.
// Reference lines: see fused.cu.
#include <cstdlib>

#include "common.cuh"
#include "fft256.cuh"
#include "fused.cuh"

namespace sep {

constexpr int kUnits = 8;                 // half-warps per CTA -> 128 threads
constexpr int kFastThreads = kUnits * 16;

template <int C, int R>
struct FastGeom {
  static constexpr int N = 256, SHIFT = N / R;
  static constexpr int FRAMES = 2 * kUnits;            // frames transformed per tile
  static constexpr int TB = FRAMES - (R - 1);          // owned frames / output hop-blocks
  static constexpr int TILE = (FRAMES - 1) * SHIFT + N;   // staged samples == accumulator span
  // Shared layout of a tile / accumulator: sample i lives at i + 16 * (i / (2 * SHIFT)).
  // The two half-warps of a warp work on frames 2 * SHIFT samples apart (a multiple of
  // 32 banks); the skew moves them 16 banks apart, so their 16-lane accesses never collide.
  static constexpr int SKEW_SPAN = 2 * SHIFT;
  static constexpr int TILE_SK = ((TILE - 1) + 16 * ((TILE - 1) / SKEW_SPAN) + 1 + 3) & ~3;
  static constexpr int WT = 16 * 18 + 8;                  // transposed window table, floats
  static constexpr int NV = FusedVals<C>::NV;
  static constexpr size_t smem(bool score) {
    return sizeof(float) * TILE_SK * ((score ? 1 + C : 1) + C)   // wave tiles + accumulators
         + sizeof(float2) * kXchFloat2 * kUnits                  // transpose buffers (also reduction scratch)
         + sizeof(float) * 2 * WT                                // analysis / synthesis windows [16][18]
         + sizeof(float2) * 256;                                 // W256^(p k) table, [k][p]
  }
  __host__ __device__ static constexpr int sk(int i) { return i + 16 * (i / SKEW_SPAN); }
};

template <int C, int R, bool SCORE>
__global__ void __launch_bounds__(kFastThreads, 3) fused256_kernel(const FusedArgs a) {
  using G = FastGeom<C, R>;
  constexpr int SHIFT = G::SHIFT, TILE = G::TILE, TSK = G::TILE_SK, TB = G::TB, SPAN = G::SKEW_SPAN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *tiles = reinterpret_cast<float *>(smem_raw);              // [1 (+C)][TSK], skewed
  float *acc = tiles + TSK * (SCORE ? 1 + C : 1);                  // [C][TSK], skewed
  float2 *xch_all = reinterpret_cast<float2 *>(acc + TSK * C);     // [kUnits][kXchFloat2]
  float *win = reinterpret_cast<float *>(xch_all + kXchFloat2 * kUnits);   // [16][18] 0.5 * analysis
  float *syn = win + G::WT;                                        // [16][18] synthesis
  float2 *twt = reinterpret_cast<float2 *>(syn + G::WT);           // [16][16] twiddles, [j][lane]

  const int T = a.T, b = blockIdx.y, tile = blockIdx.x;
  const int own_lo = tile * TB, own_hi = min(own_lo + TB, T);
  const int j0 = max(own_lo, R - 1), j1 = own_hi;                  // output hop-blocks
  const int t_lo = max(own_lo - (R - 1), 0);                       // first frame transformed
  const int64_t s0 = static_cast<int64_t>(t_lo) * SHIFT - a.pad;   // original index of tile[0]
  const int unit = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  float2 *xch = xch_all + unit * kXchFloat2;

  // ---- stage waveforms asynchronously (zeros outside [0, n)), clear accumulators ----
  constexpr int NSIG = SCORE ? 1 + C : 1;
#pragma unroll
  for (int sgn = 0; sgn < NSIG; ++sgn) {
    const float *row = sgn == 0 ? a.mix + static_cast<int64_t>(b) * a.n
                                : a.refs + (static_cast<int64_t>(b) * C + (sgn - 1)) * a.n;
    float *dst = tiles + TSK * sgn;
    const bool vec = ((reinterpret_cast<uintptr_t>(row + s0) & 15) == 0) && s0 >= 0 && s0 + TILE <= a.n;
    if (vec) {
      for (int i = threadIdx.x; i < TILE / 4; i += kFastThreads)
        cp_async16(dst + G::sk(4 * i), row + s0 + 4 * i);
    } else {
      for (int i = threadIdx.x; i < TILE; i += kFastThreads) {
        const int64_t g = s0 + i;
        if (g >= 0 && g < a.n) cp_async4(dst + G::sk(i), row + g); else dst[G::sk(i)] = 0.f;
      }
    }
  }
  // pull the tile that a CTA of the next wave will stage into L2 (no registers, no smem)
  if (a.lookahead > 0 && threadIdx.x == 0) {
    const int64_t next = static_cast<int64_t>(b) * a.tiles + tile + a.lookahead;
    const int nb = static_cast<int>(next / a.tiles), nt = static_cast<int>(next - static_cast<int64_t>(nb) * a.tiles);
    if (nb < a.batch) {
      const int n_lo = max(nt * TB - (R - 1), 0), n_hi = min(nt * TB + TB, T);
      const int64_t w0 = max(static_cast<int64_t>(n_lo) * SHIFT - a.pad, static_cast<int64_t>(0));
      const int64_t w1 = min(static_cast<int64_t>(n_lo) * SHIFT - a.pad + TILE, a.n);
#pragma unroll
      for (int sgn = 0; sgn < NSIG; ++sgn) {
        const float *row = sgn == 0 ? a.mix + static_cast<int64_t>(nb) * a.n
                                    : a.refs + (static_cast<int64_t>(nb) * C + (sgn - 1)) * a.n;
        const uintptr_t p0 = (reinterpret_cast<uintptr_t>(row + w0) + 15) & ~uintptr_t(15);
        const uintptr_t p1 = reinterpret_cast<uintptr_t>(row + w1) & ~uintptr_t(15);
        if (p1 > p0) prefetch_l2_bulk(reinterpret_cast<const void *>(p0), static_cast<unsigned>(p1 - p0));
      }
#pragma unroll
      for (int q = 0; q < C; ++q) {
        const float *base = a.masks + (static_cast<int64_t>(nb) * C + q) * T * 129;
        const uintptr_t p0 = (reinterpret_cast<uintptr_t>(base + static_cast<int64_t>(n_lo) * 129) + 15) & ~uintptr_t(15);
        const uintptr_t p1 = reinterpret_cast<uintptr_t>(base + static_cast<int64_t>(n_hi) * 129) & ~uintptr_t(15);
        if (p1 > p0) prefetch_l2_bulk(reinterpret_cast<const void *>(p0), static_cast<unsigned>(p1 - p0));
      }
    }
  }
  for (int i = threadIdx.x; i < TSK * C / 4; i += kFastThreads)
    reinterpret_cast<float4 *>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // the window tables ride the same async group
  if (threadIdx.x < G::WT / 4) {
    cp_async16(win + 4 * threadIdx.x, a.win_t + 4 * threadIdx.x);
    cp_async16(syn + 4 * threadIdx.x, a.syn_t + 4 * threadIdx.x);
  }
  cp_async16(twt + 2 * threadIdx.x, a.tw16 + 2 * threadIdx.x);   // symmetric: [j][p] == [p][j]

  // ---- this half-warp's frame pair; its mask rows are requested before the wait ----
  const int f0 = 2 * unit, f1 = f0 + 1;                // local frame indices
  const int ta = t_lo + f0, tb = t_lo + f1;            // global frame indices
  float ma[C][9], mb[C][9];
#pragma unroll
  for (int q = 0; q < C; ++q) {
    const float *base = a.masks + (static_cast<int64_t>(b) * C + q) * T * 129;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const bool bin_ok = r < 8 || l16 == 0;
      ma[q][r] = (bin_ok && ta < T) ? __ldg(base + static_cast<int64_t>(ta) * 129 + l16 + 16 * r) : 0.f;
      mb[q][r] = (bin_ok && tb < T) ? __ldg(base + static_cast<int64_t>(tb) * 129 + l16 + 16 * r) : 0.f;
    }
  }
  const float2 *tw = twt + l16;
  cp_async_wait_all();
  __syncthreads();

  // skewed positions of this lane's 16 samples of frame f0 (even) and f1 (odd): the skew term
  // is unit + a compile-time function of m
  const int base0 = f0 * SHIFT + l16 + 16 * unit;
#define SEP_POS_A(m) (base0 + 16 * (m) + 16 * ((16 * (m)) / SPAN))
#define SEP_POS_B(m) (base0 + SHIFT + 16 * (m) + 16 * ((SHIFT + 16 * (m)) / SPAN))
  const float2 *winp = reinterpret_cast<const float2 *>(win + 18 * l16);
  const float2 *synp = reinterpret_cast<const float2 *>(syn + 18 * l16);
  float2 v[16];
#pragma unroll
  for (int m = 0; m < 16; m += 2) {
    const float2 w = winp[m / 2];
    v[m] = cscale(make_float2(tiles[SEP_POS_A(m)], tiles[SEP_POS_B(m)]), w.x);
    v[m + 1] = cscale(make_float2(tiles[SEP_POS_A(m + 1)], tiles[SEP_POS_B(m + 1)]), w.y);
  }
  fft256<false>(v, tw, xch, l16);
  float2 Xa[9], Xb[9];
  split_pair(v, l16, Xa, Xb);

  float pit[C * C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) pit[i] = 0.f;
  if (SCORE) {
    const float len_f = a.lengths ? a.lengths[b] : static_cast<float>(T);
    const int len_i = static_cast<int>(len_f);
    // weights: frame counted once (owned), prediction gated by t < length (cell 28 :1031-1046)
    const float own_a = (ta >= own_lo && ta < own_hi) ? 1.f : 0.f;
    const float own_b = (tb >= own_lo && tb < own_hi) ? 1.f : 0.f;
    const float gate_a = ta < len_i ? 1.f : 0.f, gate_b = tb < len_i ? 1.f : 0.f;
    float inv_a[9], inv_b[9], mag_a[9], mag_b[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const float pa = fmaf(Xa[r].x, Xa[r].x, Xa[r].y * Xa[r].y);
      const float pb = fmaf(Xb[r].x, Xb[r].x, Xb[r].y * Xb[r].y);
      inv_a[r] = pa > 0.f ? rsqrtf(pa) : 0.f;
      inv_b[r] = pb > 0.f ? rsqrtf(pb) : 0.f;
      mag_a[r] = pa * inv_a[r] * gate_a;
      mag_b[r] = pb * inv_b[r] * gate_b;
    }
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const float *ref = tiles + TSK * (1 + j);
#pragma unroll
      for (int m = 0; m < 16; m += 2) {
        const float2 w = winp[m / 2];
        v[m] = cscale(make_float2(ref[SEP_POS_A(m)], ref[SEP_POS_B(m)]), w.x);
        v[m + 1] = cscale(make_float2(ref[SEP_POS_A(m + 1)], ref[SEP_POS_B(m + 1)]), w.y);
      }
      fft256<false>(v, tw, xch, l16);
      float2 Sa[9], Sb[9];
      split_pair(v, l16, Sa, Sb);
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const float bin_w = (r < 8 || l16 == 0) ? 1.f : 0.f;
        // label = |S| cos(angle X - angle S) = Re(S conj X) / |X| ; angle(0) = 0 -> Re S
        const float la = inv_a[r] > 0.f ? fmaf(Sa[r].x, Xa[r].x, Sa[r].y * Xa[r].y) * inv_a[r] : Sa[r].x;
        const float lb = inv_b[r] > 0.f ? fmaf(Sb[r].x, Xb[r].x, Sb[r].y * Xb[r].y) * inv_b[r] : Sb[r].x;
        const float wa = own_a * bin_w, wb = own_b * bin_w;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float da = fmaf(ma[i][r], mag_a[r], -la), db = fmaf(mb[i][r], mag_b[r], -lb);
          pit[i * C + j] = fmaf(wa * da, da, pit[i * C + j]);
          pit[i * C + j] = fmaf(wb * db, db, pit[i * C + j]);
        }
      }
    }
  }

  // ---- masked spectra -> time frames -> phased overlap-add ----
#pragma unroll
  for (int q = 0; q < C; ++q) {
    float2 L[9], Mi[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      // P = m_a X_a (frame ta, real part), Q = m_b X_b (frame tb, imaginary part)
      const float2 P = cscale(Xa[r], ma[q][r]), Q = cscale(Xb[r], mb[q][r]);
      L[r] = cadd_pi(P, Q);                                          // P + i Q
      Mi[r] = __fadd2_rn(make_float2(P.x, -P.y), make_float2(Q.y, Q.x));   // conj P + i conj Q
    }
    merge_pair(v, l16, L, Mi);
    fft256<true>(v, tw, xch, l16);
    float *accq = acc + TSK * q;
#pragma unroll
    for (int m = 0; m < 16; m += 2) {      // synthesis window on both frames at once
      const float2 w = synp[m / 2];
      v[m] = cscale(v[m], w.x);
      v[m + 1] = cscale(v[m + 1], w.y);
    }
#pragma unroll
    for (int ph = 0; ph < R; ++ph) {
      if ((f0 % R) == ph) {
#pragma unroll
        for (int m = 0; m < 16; ++m) accq[SEP_POS_A(m)] += v[m].x;
      }
      if ((f1 % R) == ph) {
#pragma unroll
        for (int m = 0; m < 16; ++m) accq[SEP_POS_B(m)] += v[m].y;
      }
      __syncthreads();
    }
  }

  // ---- write estimates, Gram partials ----
  double gram[C * C], ee[C], er[C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) gram[i] = 0.0;
#pragma unroll
  for (int i = 0; i < C; ++i) { ee[i] = 0.0; er[i] = 0.0; }
  const int64_t n_valid = (SCORE && a.valid) ? min(static_cast<int64_t>(a.valid[b]), a.n) : a.n;
  const int span = (j1 - j0) * SHIFT;
  const int first = (j0 - t_lo) * SHIFT;                      // accumulator index of output sample 0
  const int64_t g0 = static_cast<int64_t>(j0) * SHIFT - a.pad;   // its original sample index
  for (int i = threadIdx.x; i < span; i += kFastThreads) {
    const int64_t g = g0 + i;
    if (g >= a.n) break;
    float e[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      e[c] = acc[TSK * c + G::sk(first + i)];
      if (a.est) a.est[(static_cast<int64_t>(b) * C + c) * a.n + g] = e[c];
    }
    if (SCORE && g < n_valid) {
#pragma unroll
      for (int jr = 0; jr < C; ++jr) {
        const double r = static_cast<double>(tiles[TSK * (1 + jr) + G::sk(first + i)]);
        er[jr] = fma(r, r, er[jr]);
#pragma unroll
        for (int ie = 0; ie < C; ++ie) gram[ie * C + jr] = fma(static_cast<double>(e[ie]), r, gram[ie * C + jr]);
      }
#pragma unroll
      for (int ie = 0; ie < C; ++ie) {
        const double x = static_cast<double>(e[ie]);
        ee[ie] = fma(x, x, ee[ie]);
      }
    }
  }
#undef SEP_POS_A
#undef SEP_POS_B
  if (!SCORE) return;
  constexpr int NV = G::NV;
  double vals[NV];
#pragma unroll
  for (int i = 0; i < C * C; ++i) { vals[i] = static_cast<double>(pit[i]); vals[C * C + i] = gram[i]; }
#pragma unroll
  for (int i = 0; i < C; ++i) { vals[2 * C * C + i] = ee[i]; vals[2 * C * C + C + i] = er[i]; }
  block_sum<NV>(vals, reinterpret_cast<double *>(xch_all));   // transpose buffers are idle now
  if (threadIdx.x == 0) {
    double *dst = a.partials + (static_cast<int64_t>(b) * a.tiles + tile) * NV;
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = vals[i];
  }
  if (a.counters != nullptr) {      // single-launch mode: last tile of an utterance finalises it
    __shared__ int s_flag;
    finalize_in_kernel<C>(a, b, &s_flag);
  }
}

template <int C, int R, bool SCORE>
static int launch_fast(FusedArgs a, int batch, double *d_scores, double *d_sums, Scratch &s,
                       cudaStream_t stream) {
  using G = FastGeom<C, R>;
  a.tb = G::TB;
  a.tiles = (a.T + G::TB - 1) / G::TB;
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
  a.lookahead = 148 * 3;
  // one wave of resident CTAs (3 per SM)
  const size_t smem = G::smem(SCORE);
  SEP_CUDA(cudaFuncSetAttribute(fused256_kernel<C, R, SCORE>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(a.tiles, batch);
  profile_begin(stream, "fused256_kernel<C=%d,R=%d,SCORE=%d>", C, R, int(SCORE));
  fused256_kernel<C, R, SCORE><<<grid, kFastThreads, smem, stream>>>(a);
  profile_end(stream);
  SEP_LAUNCHED();
  if (SCORE && counters == nullptr) return launch_fused_finalize<C>(a, batch, d_scores, d_sums, stream);
  return SEP_OK;
}

template <int C>
static int dispatch_fast(const sep_plan *p, const FusedArgs &a, int batch, double *d_scores,
                         double *d_sums, Scratch &s, cudaStream_t stream) {
  const bool score = a.refs != nullptr;
  if (p->hops == 2)
    return score ? launch_fast<C, 2, true>(a, batch, d_scores, d_sums, s, stream)
                 : launch_fast<C, 2, false>(a, batch, d_scores, d_sums, s, stream);
  return score ? launch_fast<C, 4, true>(a, batch, d_scores, d_sums, s, stream)
               : launch_fast<C, 4, false>(a, batch, d_scores, d_sums, s, stream);
}

int fused_fast_try(const sep_plan *p, const FusedArgs &a, int batch, int C, double *d_scores,
                   double *d_sums, Scratch &s, cudaStream_t stream, bool *handled) {
  *handled = false;
  if (p->size != 256 || (p->hops != 2 && p->hops != 4)) return SEP_OK;
  if (getenv("SEPCORE_FORCE_GENERIC")) return SEP_OK;
  *handled = true;
  switch (C) {
    case 1: return dispatch_fast<1>(p, a, batch, d_scores, d_sums, s, stream);
    case 2: return dispatch_fast<2>(p, a, batch, d_scores, d_sums, s, stream);
    case 3: return dispatch_fast<3>(p, a, batch, d_scores, d_sums, s, stream);
    default: return dispatch_fast<4>(p, a, batch, d_scores, d_sums, s, stream);
  }
}

}  // namespace sep
