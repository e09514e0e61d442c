// pit.cu -- utterance-level permutation-invariant MSE (uPIT) on feature tensors.
//
// Reference: pit_with_outputsize / pit_loss, uPIT_baseline.ipynb:1023-1059 (cell
// 28); identical copy Raw_with_Convlayer.ipynb:338-374 (cell 12, F = 40).
//   y_true [B, T+1, C*F]: labels, last time row = valid length
//   y_pred [B, T,   C*F]
//   mask[b,t] = t < int(length_b)                       (:1031-1033)
//   pair[i][j] = sum_{t,f} (mask * pred_i - label_j)^2  (:1045-1052; labels unmasked)
//   cost_perm  = sum_c pair[perm[c]][c] / length_b ; idx = cost1 > cost2 (:1054)
//   loss       = sum_b min cost                         (:1055, SUM not mean)
// Every (t, f) element is read once; the C*C squared differences are formed in
// registers; partial sums are float64 and reduced in a fixed order.
#include "common.cuh"
#include "score.cuh"

namespace sep {

constexpr int kPitRows = 8;  // frames per CTA

template <int C>
__global__ void __launch_bounds__(256)
pit_pair_kernel(const float *__restrict__ y_true, const float *__restrict__ y_pred, int T, int F,
                int chunks, double *__restrict__ partials) {
  __shared__ double red[C * C * 8];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int W = C * F;
  const float *yt = y_true + static_cast<int64_t>(b) * (T + 1) * W;
  const float *yp = y_pred + static_cast<int64_t>(b) * T * W;
  const int len_i = static_cast<int>(yt[static_cast<int64_t>(T) * W]);
  const int t0 = chunk * kPitRows, t1 = min(t0 + kPitRows, T);
  float acc[C * C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) acc[i] = 0.f;
  const int total = (t1 - t0) * F;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int t = t0 + e / F, f = e % F;
    const float gate = t < len_i ? 1.f : 0.f;
    float p[C], l[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      p[c] = __ldg(yp + static_cast<int64_t>(t) * W + c * F + f) * gate;
      l[c] = __ldg(yt + static_cast<int64_t>(t) * W + c * F + f);
    }
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float d = p[i] - l[j];
        acc[i * C + j] = fmaf(d, d, acc[i * C + j]);
      }
  }
  double v[C * C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) v[i] = static_cast<double>(acc[i]);
  block_sum<C * C>(v, red);
  if (threadIdx.x == 0) {
    double *dst = partials + (static_cast<int64_t>(b) * chunks + chunk) * (C * C);
#pragma unroll
    for (int i = 0; i < C * C; ++i) dst[i] = v[i];
  }
}

// one warp per utterance; row = pair[C*C], costs[P], perm, loss
template <int C>
__global__ void pit_finalize_kernel(const double *__restrict__ partials, int chunks,
                                    const float *__restrict__ y_true, int T, int W,
                                    double *__restrict__ rows) {
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  const int b = blockIdx.x, lane = threadIdx.x;
  double v[C * C];
#pragma unroll
  for (int i = 0; i < C * C; ++i) v[i] = 0.0;
  for (int q = lane; q < chunks; q += 32) {
    const double *src = partials + (static_cast<int64_t>(b) * chunks + q) * (C * C);
#pragma unroll
    for (int i = 0; i < C * C; ++i) v[i] += src[i];
  }
#pragma unroll
  for (int i = 0; i < C * C; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
    const double len = static_cast<double>(y_true[(static_cast<int64_t>(b) * (T + 1) + T) * W]);
    finalize_pit<C>(v, len, rows + static_cast<int64_t>(b) * (C * C + P + 2));
  }
}

// Scatters rows into the caller's arrays and sums the loss in batch order.
__global__ void pit_outputs_kernel(const double *__restrict__ rows, int batch, int C, int P,
                                   double *__restrict__ pair, double *__restrict__ costs,
                                   int32_t *__restrict__ perm, double *__restrict__ loss) {
  const int stride = C * C + P + 2, lane = threadIdx.x;
  double s = 0.0;
  for (int b = lane; b < batch; b += 32) {
    const double *r = rows + static_cast<int64_t>(b) * stride;
    if (pair) for (int i = 0; i < C * C; ++i) pair[b * C * C + i] = r[i];
    if (costs) for (int p = 0; p < P; ++p) costs[b * P + p] = r[C * C + p];
    if (perm) perm[b] = static_cast<int32_t>(r[C * C + P]);
    s += r[C * C + P + 1];
  }
  s = warp_sum(s);
  if (lane == 0) *loss = s;
}

// d loss / d y_pred = 2 m (m p_i - l_c) / length for the source c the selected
// permutation assigns estimate i to (what TF autodiff yields for cell 28).
template <int C>
__global__ void pit_grad_kernel(const float *__restrict__ y_true, const float *__restrict__ y_pred,
                                const double *__restrict__ rows, int T, int F,
                                float *__restrict__ grad) {
  constexpr int P = (C == 1) ? 1 : (C == 2) ? 2 : (C == 3) ? 6 : 24;
  const int b = blockIdx.y, W = C * F;
  const float *yt = y_true + static_cast<int64_t>(b) * (T + 1) * W;
  const float *yp = y_pred + static_cast<int64_t>(b) * T * W;
  float *g = grad + static_cast<int64_t>(b) * T * W;
  const float len_f = yt[static_cast<int64_t>(T) * W];
  const int len_i = static_cast<int>(len_f);
  int perm[SEP_MAX_SOURCES], inv[SEP_MAX_SOURCES];
  nth_permutation(C, static_cast<int>(rows[static_cast<int64_t>(b) * (C * C + P + 2) + C * C + P]), perm);
  for (int c = 0; c < C; ++c) inv[perm[c]] = c;
  const int64_t total = static_cast<int64_t>(T) * W;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(e / W), w = static_cast<int>(e % W);
    const int i = w / F, f = w % F;
    const float gate = t < len_i ? 1.f : 0.f;
    const float lab = __ldg(yt + static_cast<int64_t>(t) * W + inv[i] * F + f);
    g[e] = 2.f * gate * (gate * __ldg(yp + e) - lab) / len_f;
  }
}

template <int C>
static int run_pit(const float *d_true, const float *d_pred, int batch, int T, int F, double *d_pair,
                   double *d_costs, int32_t *d_perm, double *d_loss, float *d_grad, Scratch &s,
                   cudaStream_t stream) {
  const int P = factorial(C), chunks = (T + kPitRows - 1) / kPitRows;
  double *partials, *rows;
  int rc;
  if ((rc = s.alloc(&partials, static_cast<size_t>(batch) * chunks * C * C))) return rc;
  if ((rc = s.alloc(&rows, static_cast<size_t>(batch) * (C * C + P + 2)))) return rc;
  profile_begin(stream, "pit_pair_kernel<C=%d>", C);
  pit_pair_kernel<C><<<dim3(chunks, batch), 256, 0, stream>>>(d_true, d_pred, T, F, chunks, partials);
  profile_end(stream);
  SEP_LAUNCHED();
  pit_finalize_kernel<C><<<batch, 32, 0, stream>>>(partials, chunks, d_true, T, C * F, rows);
  SEP_LAUNCHED();
  pit_outputs_kernel<<<1, 32, 0, stream>>>(rows, batch, C, P, d_pair, d_costs, d_perm, d_loss);
  SEP_LAUNCHED();
  if (d_grad) {
    const int64_t total = static_cast<int64_t>(T) * C * F;
    dim3 grid(static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 1024)), batch);
    pit_grad_kernel<C><<<grid, 256, 0, stream>>>(d_true, d_pred, rows, T, F, d_grad);
    SEP_LAUNCHED();
  }
  return SEP_OK;
}

}  // namespace sep

using namespace sep;

extern "C" int sep_pit_mse_f32(const float *y_true, const float *y_pred, int batch, int frames,
                               int feat, int n_src, double *pair, double *costs, int32_t *perm,
                               double *loss, float *grad, int mem, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SEP_REQUIRE(y_true && y_pred && loss, "sep_pit_mse_f32: null argument");
  SEP_REQUIRE(batch >= 1 && frames >= 1 && feat >= 1, "sep_pit_mse_f32: bad shape");
  SEP_REQUIRE(n_src >= 1 && n_src <= SEP_MAX_SOURCES, "sep_pit_mse_f32: n_src=%d out of range", n_src);
  int rc = check_mem(mem);
  if (rc) return rc;
  const int C = n_src, P = factorial(C), W = C * feat;
  Scratch s(stream);
  const float *d_true, *d_pred;
  double *d_pair, *d_costs, *d_loss;
  int32_t *d_perm;
  float *d_grad;
  const size_t n_true = static_cast<size_t>(batch) * (frames + 1) * W;
  const size_t n_pred = static_cast<size_t>(batch) * frames * W;
  if ((rc = stage_in(s, y_true, n_true, mem, &d_true))) return rc;
  if ((rc = stage_in(s, y_pred, n_pred, mem, &d_pred))) return rc;
  if ((rc = stage_out(s, pair, static_cast<size_t>(batch) * C * C, mem, &d_pair))) return rc;
  if ((rc = stage_out(s, costs, static_cast<size_t>(batch) * P, mem, &d_costs))) return rc;
  if ((rc = stage_out(s, perm, static_cast<size_t>(batch), mem, &d_perm))) return rc;
  if ((rc = stage_out(s, loss, static_cast<size_t>(1), mem, &d_loss))) return rc;
  if ((rc = stage_out(s, grad, n_pred, mem, &d_grad))) return rc;
  switch (C) {
    case 1: rc = run_pit<1>(d_true, d_pred, batch, frames, feat, d_pair, d_costs, d_perm, d_loss, d_grad, s, stream); break;
    case 2: rc = run_pit<2>(d_true, d_pred, batch, frames, feat, d_pair, d_costs, d_perm, d_loss, d_grad, s, stream); break;
    case 3: rc = run_pit<3>(d_true, d_pred, batch, frames, feat, d_pair, d_costs, d_perm, d_loss, d_grad, s, stream); break;
    default: rc = run_pit<4>(d_true, d_pred, batch, frames, feat, d_pair, d_costs, d_perm, d_loss, d_grad, s, stream); break;
  }
  if (rc) return rc;
  if ((rc = copy_back(s, pair, d_pair, static_cast<size_t>(batch) * C * C, mem))) return rc;
  if ((rc = copy_back(s, costs, d_costs, static_cast<size_t>(batch) * P, mem))) return rc;
  if ((rc = copy_back(s, perm, d_perm, static_cast<size_t>(batch), mem))) return rc;
  if ((rc = copy_back(s, loss, d_loss, static_cast<size_t>(1), mem))) return rc;
  if ((rc = copy_back(s, grad, d_grad, n_pred, mem))) return rc;
  return finish(s, mem);
}
