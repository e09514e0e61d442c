// bss.cu -- BSS Eval v4 (museval.metrics.bss_eval, images version) for the reference's call
//   museval.metrics.bss_eval(reference, estimated, window=np.inf, hop=np.inf, compute_permutation=True)
// at metrics/evaluate_metrics.py:79-81: SDR, ISR, SIR, SAR with 512-tap
// time-invariant distortion filters and the permutation chosen by mean SIR.
//
// museval forms the least-squares projections of every estimate on the delayed references in the time domain
// (FFT correlations -> Toeplitz Gram matrix G -> solve -> fftconvolve -> sums of squares).  Nothing of that needs to
// exist on the GPU: with d = <delayed references, estimate> and G = L L^T,
//     |P e|^2 = d^T G^-1 d = |L^-1 d|^2 ,     <P_j e, s_j> = <e, s_j> ,     |P_all e - P_j e|^2 = |P_all e|^2 - |P_j e|^2
// (nested subspaces), so every criterion is a function of the lag-0..L-1 correlation tables and of forward
// substitutions:
//     SDR = |s_j|^2 / |e_k - s_j|^2                       ISR = |s_j|^2 / (|P_j e_k|^2 - 2 <e_k, s_j> + |s_j|^2)
//     SIR = |P_j e_k|^2 / (|P_all e_k|^2 - |P_j e_k|^2)   SAR = |P_all e_k|^2 / (|e_k|^2 - |P_all e_k|^2)
// Three kernels, all float64 (museval is float64; G is ill-conditioned for speech):
//   xcorr_kernel     the correlation tables  R_ji[t] = sum_n s_j[n + t] s_i[n],  D_jk[t] = sum_n s_j[n - t] e_k[n],
//                    |e_k|^2, t < L: one CTA per (utterance, signal pair, 128 lags), 8 consecutive lags per thread on a
//                    sliding register window (64 FMAs per 16 shared-memory loads), fixed summation order
//   bss_solve_kernel one CTA per linear system: fills the (block-)Toeplitz Gram matrix from the tables, blocked
//                    right-looking Cholesky with the right-hand sides carried along as extra rows (so the forward
//                    substitution comes out of the same panel solves and trailing updates) -> |L^-1 d|^2
//   bss_final_kernel criteria table, argmax of the mean SIR over the permutations (numpy argmax semantics incl. NaN),
//                    silent source -> NaN, the NaN fallback of evaluate_metrics.py:83-86
// The CPU restatement of museval (oracle/bss_eval.py) follows the time-domain recipe literally, so the two agree only
// if both are right.  PARITY UNPINNED against museval itself (not installable here) -- see DESIGN.md section 3.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace sep {

constexpr int kBssLagBlock = 128;     // lags per xcorr CTA
constexpr int kBssChunk = 2048;       // samples staged per step
constexpr int kBssXs = kBssChunk + kBssLagBlock + 8;
constexpr int kNB = 32;               // Cholesky block size
constexpr int kTile = 64;             // trailing-update tile

__device__ __forceinline__ int skew(int i) { return i + (i >> 3); }   // one pad every 8 doubles: stride-8 reads are conflict-free

// corr[b][job][L]: job = j C + i          -> R_ji (x = ref j, y = ref i)
//                  job = C^2 + j C + k    -> D_jk (x = est k, y = ref j)
//                  job = 2 C^2 + k        -> |e_k|^2 in lag 0 (x = y = est k; only lag block 0 is computed)
__global__ void __launch_bounds__(256)
xcorr_kernel(const float *__restrict__ refs, const float *__restrict__ ests, const int64_t *__restrict__ roff,
             const int64_t *__restrict__ eoff, const int64_t *__restrict__ lens, int C, int L, double *__restrict__ corr) {
  __shared__ double ys[kBssChunk];
  __shared__ double xs[kBssXs + kBssXs / 8 + 8];
  const int b = blockIdx.y;
  const int lag_blocks = L / kBssLagBlock;
  const int job = blockIdx.x / lag_blocks, lb = blockIdx.x - job * lag_blocks;
  const int n_jobs = 2 * C * C + C;
  if (job >= 2 * C * C && lb > 0) return;
  const int64_t n = lens[b];
  const float *x, *y;
  if (job < C * C) {
    x = refs + roff[b * C + job / C];
    y = refs + roff[b * C + job % C];
  } else if (job < 2 * C * C) {
    const int q = job - C * C;
    x = ests + eoff[b * C + q % C];
    y = refs + roff[b * C + q / C];
  } else {
    x = y = ests + eoff[b * C + (job - 2 * C * C)];
  }
  const int lt = threadIdx.x & 15, g = threadIdx.x >> 4;
  const int lag0 = lb * kBssLagBlock;
  double acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.0;
  for (int64_t n0 = 0; n0 < n; n0 += kBssChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < kBssChunk; i += 256) ys[i] = n0 + i < n ? static_cast<double>(__ldg(y + n0 + i)) : 0.0;
    for (int i = threadIdx.x; i < kBssXs; i += 256) {
      const int64_t gi = n0 + lag0 + i;
      xs[skew(i)] = gi < n ? static_cast<double>(__ldg(x + gi)) : 0.0;
    }
    __syncthreads();
    // this thread: lags lag0 + 8 lt + q (q < 8), samples n0 + 128 g + s (s < 128); window w[m] = xs[128 g + s' + 8 lt + m]
    const int base = 128 * g + 8 * lt;
    double w[16];
#pragma unroll
    for (int m = 0; m < 8; ++m) w[m] = xs[skew(base + m)];
#pragma unroll 2
    for (int s = 0; s < 128; s += 8) {
#pragma unroll
      for (int m = 0; m < 8; ++m) w[8 + m] = xs[skew(base + s + 8 + m)];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double yv = ys[128 * g + s + u];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fma(yv, w[u + q], acc[q]);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) w[m] = w[8 + m];
    }
  }
  // sum the 16 sample groups in group order (deterministic)
  __syncthreads();
  double *red = xs;                                     // [16][128]
#pragma unroll
  for (int q = 0; q < 8; ++q) red[g * kBssLagBlock + 8 * lt + q] = acc[q];
  __syncthreads();
  if (threadIdx.x < kBssLagBlock) {
    double v = 0.0;
#pragma unroll
    for (int gg = 0; gg < 16; ++gg) v += red[gg * kBssLagBlock + threadIdx.x];
    corr[(static_cast<int64_t>(b) * n_jobs + job) * L + lag0 + threadIdx.x] = v;
  }
}

// Gram entry <s_j delayed p, s_i delayed q> = R_ji[q - p] (q >= p) or R_ij[p - q]
__device__ __forceinline__ double gram(const double *__restrict__ R, int C, int L, int j, int p, int i, int q) {
  return q >= p ? __ldg(R + (j * C + i) * L + (q - p)) : __ldg(R + (i * C + j) * L + (p - q));
}

// One CTA per system.  item = b (C) + sys:  sys 0 = all C references (n = C L; also yields the source-0 projection as the
// norm of the first L entries of L^-1 d: the leading block of G is G_00);  sys j >= 1 = reference j alone (n = L).
// proj[b][(C + 1) C]: rows j < C = |P_j e_k|^2, row C = |P_all e_k|^2.   A: per-CTA scratch, (n + C) x n row-major.
__global__ void __launch_bounds__(256, 1)
bss_solve_kernel(const double *__restrict__ corr, int batch, int C, int L, double *__restrict__ scratch,
                 int64_t slot_doubles, double *__restrict__ proj) {
  __shared__ double Dg[kNB][kNB + 1];
  __shared__ double invd[kNB];
  __shared__ double Pi[kNB][kTile + 2], Pj[kNB][kTile + 2];   // panel tiles, transposed: [p][row]
  __shared__ double red[256];
  double *A = scratch + static_cast<int64_t>(blockIdx.x) * slot_doubles;
  const int tid = threadIdx.x, n_jobs = 2 * C * C + C;
  const int items = batch * C;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / C, sys = item - b * C;
    const double *R = corr + static_cast<int64_t>(b) * n_jobs * L;
    const double *Dt = R + C * C * L;
    const int nsrc = sys == 0 ? C : 1, src0 = sys;          // sources in this system: src0 .. src0 + nsrc - 1
    const int n = nsrc * L, rows = n + C, ld = n;
    double maxdiag = 0.0;
    for (int j = 0; j < C; ++j) maxdiag = fmax(maxdiag, __ldg(R + (j * C + j) * L));
    const double thresh = 1e-13 * maxdiag;
    __syncthreads();
    // ---- fill the lower triangle and the right-hand-side rows ----
    for (int a = tid / 32; a < rows; a += 8) {
      const int lane = tid & 31;
      if (a < n) {
        const int j = src0 + a / L, p = a % L;
        for (int c = lane; c <= a; c += 32) A[static_cast<int64_t>(a) * ld + c] = gram(R, C, L, j, p, src0 + c / L, c % L);
      } else {
        const int k = a - n;
        for (int c = lane; c < n; c += 32) A[static_cast<int64_t>(a) * ld + c] = __ldg(Dt + ((src0 + c / L) * C + k) * L + c % L);
      }
    }
    __syncthreads();
    for (int kb = 0; kb < n; kb += kNB) {
      // ---- diagonal block: Cholesky in shared memory (warp 0) ----
      for (int e = tid; e < kNB * kNB; e += 256) {
        const int r = e / kNB, c = e % kNB;
        Dg[r][c] = c <= r ? A[static_cast<int64_t>(kb + r) * ld + kb + c] : 0.0;
      }
      __syncthreads();
      if (tid < 32) {
        const int i = tid;
        for (int j = 0; j < kNB; ++j) {
          const double piv = Dg[j][j];
          // numerically dependent direction (pivot at round-off level): drop it, i.e. project on the span of the
          // remaining delayed references -- what the eps * I of museval's solve amounts to
          const double inv = piv > thresh ? rsqrt(piv) : 0.0;
          __syncwarp();
          if (i == j) { Dg[j][j] = piv > thresh ? piv * inv : 0.0; invd[j] = inv; }
          if (i > j) Dg[i][j] *= inv;
          __syncwarp();
          if (i > j) {
            const double lij = Dg[i][j];
            for (int c = j + 1; c <= i; ++c) Dg[i][c] -= lij * Dg[c][j];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- panel: rows below the block (and the right-hand-side rows) times Dg^-T, one row per thread ----
      for (int r = kb + kNB + tid; r < rows; r += 256) {
        double *row = A + static_cast<int64_t>(r) * ld + kb;
        double x[kNB];
#pragma unroll
        for (int c = 0; c < kNB; c += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(row + c);
          x[c] = v.x;
          x[c + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
          double t = x[c];
#pragma unroll
          for (int p = 0; p < c; ++p) t = fma(-x[p], Dg[c][p], t);
          x[c] = t * invd[c];
          asm volatile("" ::: "memory");                   // keep the 496 shared-memory loads from being hoisted (spills)
        }
#pragma unroll
        for (int c = 0; c < kNB; c += 2) *reinterpret_cast<double2 *>(row + c) = make_double2(x[c], x[c + 1]);
      }
      __syncthreads();
      // ---- trailing update: A[I, J] -= P_I P_J^T over 64 x 64 tiles of the lower triangle (+ the RHS rows) ----
      const int t0 = kb + kNB;
      const int mt = (rows - t0 + kTile - 1) / kTile;           // row tiles (the last one holds the RHS rows)
      const int ct = (n - t0 + kTile - 1) / kTile;              // column tiles
      const int ty = tid >> 4, tx = tid & 15;
      for (int I = 0; I < mt; ++I) {
        for (int e = tid; e < kTile * kNB; e += 256) {
          const int r = e / kNB, p = e % kNB, gr = t0 + I * kTile + r;
          Pi[p][r] = gr < rows ? A[static_cast<int64_t>(gr) * ld + kb + p] : 0.0;
        }
        for (int J = 0; J <= I && J < ct; ++J) {
          __syncthreads();
          if (J == I) {
            for (int e = tid; e < kTile * kNB; e += 256) Pj[e % kNB][e / kNB] = Pi[e % kNB][e / kNB];
          } else {
            for (int e = tid; e < kTile * kNB; e += 256) {
              const int r = e / kNB, p = e % kNB, gr = t0 + J * kTile + r;
              Pj[p][r] = gr < n ? A[static_cast<int64_t>(gr) * ld + kb + p] : 0.0;
            }
          }
          __syncthreads();
          double acc[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll 8
          for (int p = 0; p < kNB; ++p) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&Pi[p][4 * ty]);
            const double2 a23 = *reinterpret_cast<const double2 *>(&Pi[p][4 * ty + 2]);
            const double2 b01 = *reinterpret_cast<const double2 *>(&Pj[p][4 * tx]);
            const double2 b23 = *reinterpret_cast<const double2 *>(&Pj[p][4 * tx + 2]);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int gr = t0 + I * kTile + 4 * ty + r;
            if (gr >= rows) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int gc = t0 + J * kTile + 4 * tx + c;
              if (gc < n && (gc <= gr || gr >= n)) A[static_cast<int64_t>(gr) * ld + gc] -= acc[r][c];
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- the right-hand-side rows now hold y_k = L^-1 d_k ----
    for (int k = 0; k < C; ++k) {
      const double *yrow = A + static_cast<int64_t>(n + k) * ld;
      double all = 0.0, first = 0.0;
      for (int c = tid; c < n; c += 256) {
        const double v = yrow[c];
        all = fma(v, v, all);
        if (c < L) first = fma(v, v, first);
      }
      for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        red[tid] = pass == 0 ? all : first;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
          if (tid < s) red[tid] += red[tid + s];
          __syncthreads();
        }
        if (tid == 0) {
          double *out = proj + static_cast<int64_t>(b) * (C + 1) * C;
          if (sys == 0) {
            if (pass == 0) out[C * C + k] = red[0];       // |P_all e_k|^2
            else out[0 * C + k] = red[0];                 // |P_0 e_k|^2
          } else if (pass == 0) {
            out[sys * C + k] = red[0];                    // |P_sys e_k|^2
          }
        }
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double safe_db(double num, double den) {
  if (den <= 0.0) return CUDART_INF;                        // museval._safe_db: den == 0 -> inf
  return 10.0 * log10(num / den);
}

// row[b]: [4][C][C] criteria table (sdr, isr, sir, sar; [jtrue][jest]), perm index (lexicographic, perm[jtrue] = jest),
// value (mean SDR of the selection with the NaN fallback), selected SDR per source [C]
__global__ void bss_final_kernel(const double *__restrict__ corr, const double *__restrict__ proj, int batch, int C, int L,
                                 double *__restrict__ rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int n_jobs = 2 * C * C + C, width = 4 * C * C + 2 + C;
  const double *R = corr + static_cast<int64_t>(b) * n_jobs * L;
  const double *P = proj + static_cast<int64_t>(b) * (C + 1) * C;
  double *row = rows + static_cast<int64_t>(b) * width;
  double ss[SEP_MAX_SOURCES], ee[SEP_MAX_SOURCES];
  bool silent = false;
  for (int j = 0; j < C; ++j) {
    ss[j] = R[(j * C + j) * L];
    ee[j] = R[(2 * C * C + j) * L];
    silent = silent || ss[j] == 0.0 || ee[j] == 0.0;        // museval._any_source_silent (an all-zero signal)
  }
  const double nan = CUDART_NAN;
  for (int j = 0; j < C; ++j)
    for (int k = 0; k < C; ++k) {
      const double dot = R[(C * C + j * C + k) * L], pj = P[j * C + k], pall = P[C * C + k];
      row[0 * C * C + j * C + k] = silent ? nan : safe_db(ss[j], ee[k] - 2.0 * dot + ss[j]);
      row[1 * C * C + j * C + k] = silent ? nan : safe_db(ss[j], pj - 2.0 * dot + ss[j]);
      row[2 * C * C + j * C + k] = silent ? nan : safe_db(pj, pall - pj);
      row[3 * C * C + j * C + k] = silent ? nan : safe_db(pall, ee[k] - pall);
    }
  int perm[SEP_MAX_SOURCES], best = 0;
  double best_v = 0.0;
  int P_count = 1;
  for (int i = 2; i <= C; ++i) P_count *= i;
  for (int p = 0; p < P_count; ++p) {
    nth_permutation(C, p, perm);
    double m = 0.0;
    for (int j = 0; j < C; ++j) m += row[2 * C * C + j * C + perm[j]];
    m /= C;
    // np.argmax: the first maximum; a NaN counts as the maximum (the first NaN wins)
    if (p == 0 || (!isnan(best_v) && (isnan(m) || m > best_v))) { best_v = m; best = p; }
  }
  nth_permutation(C, best, perm);
  double mean = 0.0, mean0 = 0.0;
  for (int j = 0; j < C; ++j) {
    const double v = row[j * C + perm[j]];
    row[4 * C * C + 2 + j] = v;
    mean += v;
    // np.nan_to_num: nan -> 0, +-inf -> +-DBL_MAX  (evaluate_metrics.py:85-86)
    mean0 += isnan(v) ? 0.0 : (isinf(v) ? copysign(1.7976931348623157e308, v) : v);
  }
  mean /= C;
  if (isnan(mean)) mean = mean0 / C;
  row[4 * C * C] = best;
  row[4 * C * C + 1] = mean;
}

}  // namespace sep

using namespace sep;

extern "C" int sep_bss_eval_row_width(int n_src) { return 4 * n_src * n_src + 2 + n_src; }

// ref_offsets / est_offsets / lengths are HOST arrays in both memory modes, like sep_score_batch_f32.
extern "C" int sep_bss_eval_f32(const float *refs, const float *ests, const int64_t *ref_offsets,
                                const int64_t *est_offsets, const int64_t *lengths, int batch, int n_src,
                                int64_t total_ref, int64_t total_est, int filters_len, double *rows, int mem,
                                void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(refs && ests && ref_offsets && est_offsets && lengths && rows, "sep_bss_eval_f32: null argument");
  SEP_REQUIRE(batch >= 1 && batch <= 65535, "sep_bss_eval_f32: batch must be in [1, 65535]");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES, "sep_bss_eval_f32: n_src=%d out of range", n_src);
  if (filters_len < kBssLagBlock || filters_len > 1024 || filters_len % kBssLagBlock != 0) {
    set_error("sep_bss_eval_f32: filters_len=%d unsupported (a multiple of %d in [%d, 1024]; museval's default is 512)",
              filters_len, kBssLagBlock, kBssLagBlock);
    return SEP_ERR_UNSUPPORTED;
  }
  int rc = check_mem(mem);
  if (rc) return rc;
  const int C = n_src, L = filters_len;
  const size_t n_off = static_cast<size_t>(batch) * C;
  for (int b = 0; b < batch; ++b) {
    SEP_REQUIRE(lengths[b] >= 1, "sep_bss_eval_f32: empty utterance %d", b);
    for (int c = 0; c < C; ++c) {
      SEP_REQUIRE(ref_offsets[b * C + c] >= 0 && ref_offsets[b * C + c] + lengths[b] <= total_ref,
                  "sep_bss_eval_f32: reference %d/%d out of bounds", b, c);
      SEP_REQUIRE(est_offsets[b * C + c] >= 0 && est_offsets[b * C + c] + lengths[b] <= total_est,
                  "sep_bss_eval_f32: estimate %d/%d out of bounds", b, c);
    }
  }
  Scratch s(stream);
  const float *d_refs, *d_ests;
  if ((rc = stage_in(s, refs, static_cast<size_t>(total_ref), mem, &d_refs))) return rc;
  if ((rc = stage_in(s, ests, static_cast<size_t>(total_est), mem, &d_ests))) return rc;
  int64_t *d_meta;
  if ((rc = s.alloc(&d_meta, 2 * n_off + batch))) return rc;
  // pageable host metadata: these copies return once the source has been staged
  SEP_CUDA(cudaMemcpyAsync(d_meta, ref_offsets, n_off * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  SEP_CUDA(cudaMemcpyAsync(d_meta + n_off, est_offsets, n_off * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  SEP_CUDA(cudaMemcpyAsync(d_meta + 2 * n_off, lengths, batch * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  const int n_jobs = 2 * C * C + C;
  double *corr, *proj, *scratch, *d_rows;
  if ((rc = s.alloc(&corr, static_cast<size_t>(batch) * n_jobs * L))) return rc;
  if ((rc = s.alloc(&proj, static_cast<size_t>(batch) * (C + 1) * C))) return rc;
  const int width = sep_bss_eval_row_width(C);
  if ((rc = stage_out(s, rows, static_cast<size_t>(batch) * width, mem, &d_rows))) return rc;
  // only lag 0 of the |e_k|^2 jobs is written by xcorr_kernel; the rest of those rows is never read
  const dim3 grid_x(n_jobs * (L / kBssLagBlock), batch);
  profile_begin(stream, "xcorr_kernel (fp64, %d lags) + bss_solve_kernel (blocked Cholesky, n=%d)", L, C * L);
  xcorr_kernel<<<grid_x, 256, 0, stream>>>(d_refs, d_ests, d_meta, d_meta + n_off, d_meta + 2 * n_off, C, L, corr);
  SEP_LAUNCHED();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = batch * C;
  const int ctas = std::min(items, 2 * sms);
  const int64_t slot = (static_cast<int64_t>(C) * L + C) * (static_cast<int64_t>(C) * L);
  if ((rc = s.alloc(&scratch, static_cast<size_t>(ctas) * slot))) return rc;
  bss_solve_kernel<<<ctas, 256, 0, stream>>>(corr, batch, C, L, scratch, slot, proj);
  SEP_LAUNCHED();
  bss_final_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(corr, proj, batch, C, L, d_rows);
  profile_end(stream);
  SEP_LAUNCHED();
  if ((rc = copy_back(s, rows, d_rows, static_cast<size_t>(batch) * width, mem))) return rc;
  return finish(s, mem);
}
