/*
 * sepcore.h -- C ABI of libsepcore.so: the B200-native (sm_100a) separation
 * signal path of jsjs4013/Speech-Separation-Project-with-AI.
 *
 * The reference has no FFI: its boundary is a set of Python functions called by
 * the notebooks.  Each entry point below is what a binding for that function
 * would call; the reference interface it replaces is cited as file:line
 * (notebooks: raw JSON line + 0-based cell).  The Python host side
 * (speech-separation-project-with-ai_b200/sepcore) binds these with ctypes and
 * re-exports the reference's own names and signatures.
 *
 * Conventions
 *   - every function returns SEP_OK (0) or a negative SEP_ERR_* code; the text
 *     of the last error on the calling thread is sep_last_error().  Nothing
 *     throws across the ABI.
 *   - all buffers are caller-owned.  `mem` says where they live:
 *       SEP_MEM_HOST   : host memory (pageable or pinned).  The library stages
 *                        H2D / D2H on `stream` and synchronises it before
 *                        returning (drop-in mode).
 *       SEP_MEM_DEVICE : device memory on the plan's device.  Fully
 *                        asynchronous on `stream`, no copies (throughput mode).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - a plan is immutable after creation; entry points are re-entrant across
 *     streams and threads (scratch is stream-ordered, cudaMallocAsync).
 *   - complex arrays are interleaved (re, im) float pairs.
 *   - F = size/2 + 1 bins; T frames; C sources; P = C! permutations in
 *     lexicographic order, perm[c] = index of the estimate assigned to source c.
 */
#ifndef SEPCORE_H_
#define SEPCORE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEP_OK               0
#define SEP_ERR_INVALID     (-1)  /* bad argument                                  */
#define SEP_ERR_UNSUPPORTED (-2)  /* legal in the reference, not built here       */
#define SEP_ERR_CUDA        (-3)  /* CUDA runtime error (text in sep_last_error)  */
#define SEP_ERR_NOMEM       (-4)

#define SEP_MEM_HOST   0
#define SEP_MEM_DEVICE 1

#define SEP_MAX_SOURCES 4

/* activation codes of sep_conv1d_f32 */
#define SEP_ACT_LINEAR  0
#define SEP_ACT_SIGMOID 1
#define SEP_ACT_RELU    2
/* padding codes of sep_conv1d_f32 (Keras names) */
#define SEP_PAD_VALID 0
#define SEP_PAD_SAME  1

typedef struct sep_plan sep_plan;

/* ---- library ----------------------------------------------------------- */
int         sep_version(void);
const char *sep_last_error(void);
/* Number of kernels this library has launched in the calling process so far
 * (all threads); bench.py reports the difference over the timed region. */
int64_t     sep_launch_count(void);
/* Name (with its template parameters) of the dominant kernel the last entry point called on this
 * thread launched, e.g. "wstrip256_kernel<C=2,R=2,SCORE=1>"; bench.py prints it in `roofline.kernel`. */
const char *sep_last_kernel(void);

/* Per-kernel timing for bench.py's roofline leg: while enabled, entry points
 * bracket their dominant kernel with CUDA events on the launching stream (not
 * during stream capture).  sep_profile_collect synchronises on the recorded
 * events, returns the summed kernel time and the number of bracketed launches,
 * and resets the log. */
int sep_profile_enable(int on);
int sep_profile_collect(double *total_ms, int *launches);

/* ---- plan: window, twiddles, frame geometry ---------------------------- */
/* `window` holds `size` float64 taps: the host evaluates the reference's
 * window callable (scipy, sym=True), including the window_length zero padding
 * of parallel_stft.py:183-187.  `fading` as in parallel_stft.py:169-173.
 * size: power of two in [32, 4096]; 1 <= shift <= size.  iSTFT and the fused
 * path additionally need size % shift == 0 (uPIT_baseline.ipynb:1245, cell 38).
 * The plan is bound to the CUDA device current at creation. */
int sep_plan_create(sep_plan **plan, int size, int shift, const double *window, int fading);
int sep_plan_destroy(sep_plan *plan);

/* _samples_to_stft_frames / _stft_frames_to_samples (parallel_stft.py:125-144)
 * applied the way stft() applies them (:169-180): frames for an utterance of
 * n_samples (before fade padding), and the iSTFT output length for `frames`
 * (uPIT_baseline.ipynb:1298-1305, cell 39). */
int sep_plan_frames(const sep_plan *plan, int64_t n_samples, int *frames);
int sep_plan_istft_samples(const sep_plan *plan, int frames, int64_t *n_samples);

/* _biorthogonal_window_loopy (uPIT_baseline.ipynb:1234-1259, cell 38), computed
 * once at plan creation in float64, including the reference's `index + 1 <
 * size` quirk.  out[size]. */
int sep_plan_synthesis_window(const sep_plan *plan, double *out);

/* ---- a2: segment_axis (parallel_stft.py:37-123) ------------------------ */
/* Device framing copy: in [batch, n] -> out [batch, frames, length] with
 * frames = 1 + (n - length) / (length - overlap); requires (n - length) %
 * (length - overlap) == 0 (the host handles end='cut'/'pad'/'wrap'). */
int sep_segment_axis_f32(const float *in, int batch, int64_t n, int length, int overlap,
                         float *out, int mem, void *stream);

/* ---- a3: stft (parallel_stft.py:146-196) ------------------------------- */
/* wave [batch, n_samples] (row stride `wave_stride` elements, >= n_samples) ->
 * spec [batch, T, F] complex64 interleaved. */
int sep_stft_f32(const sep_plan *plan, const float *wave, int batch, int64_t n_samples,
                 int64_t wave_stride, float *spec, int mem, void *stream);

/* ---- a4: |X|, angle X, PSA labels (parallel_stft.py:262-272) ----------- */
/* mix [batch, n], refs [batch, C, n] (may be NULL when n_src == 0) ->
 * feats  [batch, T, 2F] = |X| || angle X            (network input rows, :271)
 * labels [batch, T, C*F] = |S_c| cos(angle X - angle S_c)     (:272)
 * either output may be NULL. */
int sep_stft_features_f32(const sep_plan *plan, const float *mix, const float *refs,
                          int batch, int n_src, int64_t n_samples,
                          float *feats, float *labels, int mem, void *stream);

/* ---- a8: istft (uPIT_baseline.ipynb:1269-1307, cell 39) ---------------- */
/* spec [batch, T, F] complex64 -> wave [batch, L], L = sep_plan_istft_samples.
 * Imaginary parts of the DC and Nyquist bins are ignored like numpy's irfft. */
int sep_istft_f32(const sep_plan *plan, const float *spec, int batch, int frames,
                  float *wave, int mem, void *stream);

/* ---- a5+a6+a8: mask/phase recombination + istft ------------------------ */
/* cleaned [batch, T, C*F] (= mask_c * |X|, uPIT_baseline.ipynb:1087-1088 cell
 * 29), phase [batch, T, F] -> wave [batch, C, L]; spec_c = cleaned_c *
 * exp(j phase) (uPIT_baseline.ipynb:1385-1388, cell 41) is formed in registers
 * and never stored. */
int sep_recombine_istft_f32(const sep_plan *plan, const float *cleaned, const float *phase,
                            int batch, int n_src, int frames, float *wave,
                            int mem, void *stream);

/* ---- the fused hot path: a3 + a4 + a5 + a6 + a8 + a9 + a10/a11 (+a13) --- */
/* Layout of one row of `scores` (doubles; permutation indices are exact):
 *   [0,            C*C)  pit_pair[i][j] = sum_{t,f} (m*mask_i|X| - label_j)^2
 *   [.., +P)             pit_costs[p]   = sum_c pit_pair[perm_p[c]][c] / length
 *   +1                   pit_perm  (first minimum; tie -> perm 0, cell 28 :1054)
 *   +1                   pit_loss  (= pit_costs[pit_perm])
 *   +C*C                 si_pair[i][j]  = SI-SDR(ref_j, est_i) in dB
 *   +1                   si_best   (mean over sources for the best permutation;
 *                                   C=2: 0.5*max(sdr1, sdr2), evaluate_metrics.py:28-34)
 *   +1                   si_perm   (`sdr1 > sdr2` keeps perm 0, tie/NaN -> last)
 *   +C*C                 sdr_pair[i][j] = 10 log10(|r_j|^2 / |e_i - r_j|^2)
 *   +1                   sdr_best, +1 sdr_perm  (max mean; museval parity unpinned)
 * sep_score_stride(C) = 3*C*C + P + 6. */
int sep_score_stride(int n_src);

/* mix [batch, n], masks [batch, C, T, F] (T = sep_plan_frames(n)),
 * refs [batch, C, n] or NULL, frame_lengths [batch] float (valid frames of each
 * utterance, the `length` row of y_true; NULL = T), valid_samples [batch] int32
 * (samples scored by SI-SDR/SDR; NULL = n)
 *   -> est [batch, C, n] (may be NULL), scores [batch, sep_score_stride(C)]
 *      (NULL unless refs given), sums[4] = {sum pit_loss, sum si_best,
 *      sum sdr_best, batch} (may be NULL).
 * est_c = istft(mask_c * X)[:n]; spectra never touch HBM. */
int sep_fused_separate_f32(const sep_plan *plan, const float *mix, const float *masks,
                           const float *refs, const float *frame_lengths,
                           const int32_t *valid_samples, int batch, int n_src,
                           int64_t n_samples, float *est, double *scores, double *sums,
                           int mem, void *stream);

/* As sep_fused_separate_f32, with caller-provided device scratch so that no
 * allocation happens inside the call (CUDA-graph capture, steady-state
 * serving).  workspace must hold sep_fused_workspace_bytes() bytes of device
 * memory and must be ZERO-FILLED before its first use (it carries the
 * completion counters of the in-kernel finalisation; every successful call
 * leaves them at zero, so one memset at allocation time is enough).  Do not
 * share one workspace between calls that may run concurrently.  Only
 * meaningful with SEP_MEM_DEVICE (host mode ignores it). */
int sep_fused_workspace_bytes(const sep_plan *plan, int batch, int n_src, int64_t n_samples,
                              int64_t *bytes);
int sep_fused_separate_ws_f32(const sep_plan *plan, const float *mix, const float *masks,
                              const float *refs, const float *frame_lengths,
                              const int32_t *valid_samples, int batch, int n_src,
                              int64_t n_samples, float *est, double *scores, double *sums,
                              void *workspace, int64_t workspace_bytes, int mem, void *stream);

/* ---- a9: pit_loss (uPIT_baseline.ipynb:1023-1059, cell 28) ------------- */
/* y_true [batch, T+1, C*F] (last time row = valid length), y_pred [batch, T, C*F]
 *   -> pair [batch, C, C], costs [batch, P], perm [batch] (int32), loss[1]
 * (batch SUM of the selected costs, :1055).  Outputs other than loss may be NULL.
 * If grad is not NULL it receives d loss / d y_pred [batch, T, C*F]. */
int sep_pit_mse_f32(const float *y_true, const float *y_pred, int batch, int frames,
                    int feat, int n_src, double *pair, double *costs, int32_t *perm,
                    double *loss, float *grad, int mem, void *stream);

/* ---- a10-a13: SI-SDR / SDR scoring (metrics/evaluate_metrics.py:14-92) -- */
/* Ragged batch.  Utterance b has n_b = lengths[b] samples; its C reference
 * signals start at refs + ref_offsets[b*C + c] and its estimates at
 * ests + est_offsets[b*C + c] (element offsets; the host applies the
 * truncate-to-min-length rule of :46-48 by choosing lengths[b]).
 * scores [batch, 2*C*C + 4] = si_pair, si_best, si_perm, sdr_pair, sdr_best,
 * sdr_perm; sums[3] = {sum si_best, sum sdr_best, batch}.
 * total_ref / total_est = element counts of the two flat arrays (needed to
 * stage them in SEP_MEM_HOST mode).  ref_offsets / est_offsets / lengths are
 * HOST arrays in both memory modes (small metadata computed on the host). */
int sep_score_batch_f32(const float *refs, const float *ests,
                        const int64_t *ref_offsets, const int64_t *est_offsets,
                        const int64_t *lengths, int batch, int n_src,
                        int64_t total_ref, int64_t total_est,
                        double *scores, double *sums, int mem, void *stream);

/* pow_norm(s1, s2) = sum(s1 * s2) and pow_np_norm(s) = pow_norm(s, s)
 * (metrics/evaluate_metrics.py:14-20), accumulated in float64.  out[1]. */
int sep_dot_f32(const float *a, const float *b, int64_t n, double *out, int mem, void *stream);

/* ---- a14: Conv1D filterbank (Raw_with_Convlayer.ipynb:389, cell 13) ---- */
/* x [batch, rows, c_in], kernel [taps, c_in, filters] (Keras layout), bias
 * [filters] or NULL -> out [batch, rows_out, filters], activation fused.
 * rows_out = ceil(rows/stride) ('same') or (rows - taps)/stride + 1 ('valid'). */
int sep_conv1d_f32(const float *x, const float *kernel, const float *bias, int batch,
                   int rows, int c_in, int taps, int filters, int stride, int padding,
                   int activation, float *out, int mem, void *stream);

/* ---- BASELINE config 5: learned conv encoder / decoder filterbank on tcgen05 ---- */
/* A generalisation of a14 (the reference's Raw_with_Convlayer.ipynb:389 has an encoder
 * on non-overlapping segments and no decoder): frames[k] = wave[k*stride : k*stride+taps],
 * code = relu(frames @ enc), est_c = overlap_add((code * mask_c) @ dec, stride).
 * wave [batch, n]; enc [taps, filters]; dec [filters, taps]; masks [batch, C, K, filters],
 * K = (n - taps)/stride + 1  ->  est [batch, C, (K-1)*stride + taps]; code [batch, K,
 * filters] (optional, may be NULL).  Built for taps=16, filters=256, stride=8 (tensor
 * cores: tcgen05.mma kind::tf32 with 3xTF32 operand splitting, accumulators in TMEM);
 * other shapes return SEP_ERR_UNSUPPORTED.  n must be a multiple of 4. */
int sep_filterbank_separate_f32(const float *wave, const float *enc, const float *dec,
                                const float *masks, int batch, int n_src, int64_t n_samples,
                                int taps, int n_filters, int stride, float *est, float *code,
                                int mem, void *stream);

/* ---- sample formats either side of the path ---- */
/* int16 PCM -> float32 as wavread / librosa.load deliver the reference's 16-bit wavs
 * (metrics/evaluate_metrics.py:7-12, parallel_stft.py:213): out[i] = pcm[i] / 32768. */
int sep_pcm16_to_f32(const int16_t *pcm, int64_t n, float *out, int mem, void *stream);

/* audiowrite's sample conversion (uPIT_baseline.ipynb:1317-1354, cell 40) for float32 data
 * [batch, n], row by row: optional peak normalisation (data /= max|data|), data *= 32767,
 * clipped[row] = count(data > 32767), clip to [-32768, 32767], truncate to int16.
 * clipped may be NULL.  Arithmetic is float32 like numpy's on a float32 array. */
int sep_audiowrite_i16_f32(const float *data, int batch, int64_t n, int normalize, int16_t *out,
                           int64_t *clipped, int mem, void *stream);
/* The same conversion in float64 -- the dtype the reference's istft (cell 39) hands to audiowrite (cell 41
 * :1403-1404): numpy keeps a float64 array in float64, so truncation happens on the float64 product.
 * An all-zero row with normalize=1 is 0/0 = NaN in the reference, which astype(np.int16) turns into 0. */
int sep_audiowrite_i16_f64(const double *data, int batch, int64_t n, int normalize, int16_t *out,
                           int64_t *clipped, int mem, void *stream);

/* ---- per-batch reduction without a collective on the step ("B200-native refinement") ---- */
/* The path's only cross-utterance operation is the SUM of [loss, SI-SDR, SDR, n] over the batch (pit_loss
 * uPIT_baseline.ipynb:1055; eval_* evaluate_metrics.py:53,90).  With one process per GPU the reference design is one
 * all-reduce of those 32 bytes per batch -- 30-50 us of collective latency against a 21 us step.  Here the fused
 * kernel itself, in the epilogue that adds the batch sums, stores this rank's four float64 values into EVERY rank's
 * inbox over NVLink (peer-mapped memory, plain stores + a system-scope fence + an arrival counter): an all-gather by
 * one-sided puts; a reader adds the `world` rows once its own synchronisation says the step is done
 * (arrived[slot] == world).  No kernel ever waits on another rank.
 *   inbox layout (one per rank, sep_peer_alloc): [slots][world][4] float64, then [slots] uint64 arrival counters.
 *   sep_peer_alloc   cudaMalloc + zero fill + IPC handle (64 bytes) to send to the other ranks (any transport)
 *   sep_peer_open    maps another rank's inbox from its handle;  sep_peer_close / sep_peer_free undo the two
 *   sep_fused_separate_push_f32 = sep_fused_separate_ws_f32 in device mode + the push: peer_inboxes is a DEVICE array
 *   of `world` inbox pointers (this rank's own allocation at index `rank`), `slot` the row to write. */
int sep_peer_alloc(int64_t bytes, void **ptr, unsigned char *handle64);
int sep_peer_open(const unsigned char *handle64, void **ptr);
int sep_peer_close(void *ptr);
int sep_peer_free(void *ptr);
int sep_fused_separate_push_f32(const sep_plan *plan, const float *mix, const float *masks,
                                const float *refs, const float *frame_lengths,
                                const int32_t *valid_samples, int batch, int n_src,
                                int64_t n_samples, float *est, double *scores, double *sums,
                                void *workspace, int64_t workspace_bytes,
                                void *const *peer_inboxes, int world, int rank, int slot, int slots,
                                void *stream);

/* ---- feature / label records ---- */
/* The record the reference writes for every utterance: make_sequence_example(inputs, labels, length, name)
 * (parallel_stft_single.py:238-254, parallel_stft.py:217-229) serialised and framed by tf.io.TFRecordWriter
 * (:287-309): a tf.train.SequenceExample with feature lists 'inputs' [frames x width_inputs float32],
 * 'labels' [frames x width_labels], 'length' [1 float], 'name' [utf-8 bytes], inside the TFRecord framing
 * (u64 length, masked CRC-32C, payload, masked CRC-32C).  Host buffers only (byte formatting, no kernel); the
 * arrays are what sep_stft_features_f32 returns.  key_order: NULL (inputs, labels, length, name) or a permutation
 * of {0: inputs, 1: labels, 2: length, 3: name} -- protobuf leaves the order of map entries open and the
 * reference's committed files differ in it.  Byte-exact against those files (tests/test_records.py). */
/* TFRecord's masked CRC-32C of a byte string (frames payloads that were serialised elsewhere). */
uint32_t sep_record_masked_crc(const uint8_t *data, int64_t n);
int sep_record_size(int frames, int width_inputs, int width_labels, int name_len, int64_t *bytes);
int sep_record_encode(const float *inputs, const float *labels, int frames, int width_inputs,
                      int width_labels, float length, const char *name, int name_len,
                      const int *key_order, uint8_t *out, int64_t capacity, int64_t *written);

/* ---- BSS Eval v4 ---- */
/* museval.metrics.bss_eval(reference, estimated, window=np.inf, hop=np.inf, compute_permutation=True) as
 * eval_sdr calls it (metrics/evaluate_metrics.py:79-81): one window, `filters_len`-tap time-invariant distortion
 * filters (museval default 512), images criteria, single-channel signals.  Ragged batch in the layout of
 * sep_score_batch_f32 (flat signals, HOST offset / length arrays).  rows[b] holds sep_bss_eval_row_width(C) =
 * 4 C^2 + 2 + C float64 values:
 *   [0, C^2)        SDR[jtrue][jest]   = 10 log10 |s|^2 / |e - s|^2
 *   [C^2, 2C^2)     ISR[jtrue][jest]   = |s|^2 / |P_j e - s|^2
 *   [2C^2, 3C^2)    SIR[jtrue][jest]   = |P_j e|^2 / |P_all e - P_j e|^2
 *   [3C^2, 4C^2)    SAR[jtrue][jest]   = |P_all e|^2 / |e - P_all e|^2     (P: projection on the delayed references)
 *   4C^2            index of the permutation with the largest mean SIR (itertools.permutations order, numpy argmax
 *                   semantics; perm[jtrue] = jest)
 *   4C^2 + 1        mean SDR of that assignment with eval_sdr's NaN fallback (:83-86: mean of nan_to_num)
 *   4C^2 + 2 ..     SDR[jtrue][perm[jtrue]]
 * A silent (all-zero) reference or estimate makes every criterion NaN like museval.  float64 arithmetic on the
 * device; filters_len must be a multiple of 128 in [128, 1024].  PARITY UNPINNED against museval (not installable). */
int sep_bss_eval_row_width(int n_src);
int sep_bss_eval_f32(const float *refs, const float *ests, const int64_t *ref_offsets,
                     const int64_t *est_offsets, const int64_t *lengths, int batch, int n_src,
                     int64_t total_ref, int64_t total_est, int filters_len, double *rows, int mem,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEPCORE_H_ */
