/*
 * avtex.h — C ABI of libavtex.so: the B200 (sm_100a) transition-matrix engine.
 *
 * The reference (medhini/audio-video-textures) is pure Python/PyTorch and has no FFI or plugin
 * interface; its hot path sits behind Python callables.  Each entry point below replaces the
 * tensor arithmetic of one of those callables; the "replaces:" line cites the reference code
 * (paths relative to the reference root; classic/ = baselines/classic_video_textures/,
 * cvt/ = contrastive_video_textures/).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; avtex_last_error() then returns
 *     a thread-local, NUL-terminated message.  Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers on `device` unless the name starts with h_.
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS.
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); every call is
 *     asynchronous on it and performs no host synchronisation unless stated.
 *   - accumulators (`double *sum`, `unsigned long long *nnz`, ...) are ADDED to: zero them first
 *     (avtex_zero) — this is what lets row-shards and fused stages share one accumulator.
 *   - no hidden allocation: workspaces are caller-provided.
 */
#ifndef AVTEX_H_
#define AVTEX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVTEX_ABI_VERSION 3
#if defined(__GNUC__)
#define AVTEX_API __attribute__((visibility("default")))
#else
#define AVTEX_API
#endif

AVTEX_API int avtex_abi_version(void);
AVTEX_API const char *avtex_last_error(void);
/* Number of SMs / compute capability major*10+minor of `device` (host query, no stream). */
AVTEX_API int avtex_device_info(int device, int *sm_count, int *cc);
/* cudaMemsetAsync(ptr, 0, bytes). */
AVTEX_API int avtex_zero(void *ptr, int64_t bytes, int device, void *stream);

/* ---------------------------------------------------------------- K0: frame packing + norms
 * Packs frames [n, k] into centred signed bytes  packed[i, c] = frames[i, c] - 128  (zero padded to
 * kp columns, kp % 128 == 0) and writes the exact squared norms sqnorm[i] = sum_c packed[i,c]^2.
 * Pixel distances are translation invariant, so the centring changes no result; it keeps the
 * int32 tensor-core accumulator of avtex_gram_l2_s8 inside its range.
 * The f32 variant requires integer-valued inputs in [0, 255] (a uint8-derived video as float);
 * it sets flags[0] = 1 if any element is not, and the caller must then use
 * avtex_pairdist_direct_f32.  max_sqnorm (nullable, zeroed by the caller) receives max_i sqnorm[i]
 * (atomicMax), so the host can launch avtex_gram_l2_s8 speculatively and validate its domain
 * (4 * max < 2^32) later with one small D2H read instead of a reduction + sync in between.
 * replaces: the operand staging of classic/computeD1.py:61-83 (`frames[i:i+bs].cuda()`, repeat, view).
 */
AVTEX_API int avtex_pack_frames_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld,
                         int8_t *packed, int64_t kp, int64_t *sqnorm,
                         unsigned long long *max_sqnorm, int device, void *stream);
AVTEX_API int avtex_pack_frames_f32(const float *frames, int64_t n, int64_t k, int64_t ld,
                          int8_t *packed, int64_t kp, int64_t *sqnorm, int *flags,
                          unsigned long long *max_sqnorm, int device, void *stream);

/* ---------------------------------------------------------------- K1: pairwise L2 distances
 * D[(r - row0), j] = sqrt( sqnorm[r] + sqnorm[j] - 2 * <packed[r], packed[j]> ),  r in [row0, row0+rows),
 * j in [0, n).  The Gram is computed on tcgen05 (kind::i8, s8 x s8 -> s32 in TMEM) from TMA-fed
 * 128B-swizzled shared-memory tiles; the norm/sqrt epilogue is fused; d^2 is an exact integer so
 * duplicate frames give exactly 0 (the reference counts `nonzero(D1)`).  If `sum`/`nnz` are non-NULL
 * the fp64 sum and the non-zero count of the written block are added to them.
 * `symmetric` != 0 (requires row0 == 0, rows == n): only tiles on/above the diagonal are computed
 * and mirrored.  Precondition (checked by the host wrapper): 4 * max(sqnorm) < 2^32.
 * replaces: classic/computeD1.py:50-56 (dense) and :58-96 (tiled "slow" path), RGB branch.
 */
AVTEX_API int avtex_gram_l2_s8(const int8_t *packed, int64_t n, int64_t kp, const int64_t *sqnorm,
                     int64_t row0, int64_t rows, int symmetric, float *D, int64_t ldd,
                     double *sum, unsigned long long *nnz, int device, void *stream);

/* Same kernel on the RAW unsigned frames (no packed copy): frames [n, k] u8 with a row pitch `ld` that is
 * a multiple of 16 bytes, sqnorm[i] = sum_c frames[i,c]^2 from avtex_frame_norms_u8.  u8 x u8 products
 * are accumulated in the int32 TMEM accumulator modulo 2^32 (it wraps for typical video: sum x*y ~ 3e9 at
 * K = 150528); together with the mod-2^32 epilogue the result is exact under the same precondition as
 * above, which is stated on the CENTRED norms: 4 * max_i sum_c (frames[i,c]-128)^2 < 2^32
 * (avtex_frame_norms_u8 returns that maximum).  K is not padded: TMA zero-fills columns >= k. */
AVTEX_API int avtex_gram_l2_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, const int64_t *sqnorm,
                     int64_t row0, int64_t rows, int symmetric, float *D, int64_t ldd,
                     double *sum, unsigned long long *nnz, int device, void *stream);
/* General form (one launch, up to 32 rectangles): job j covers frames rows [row0, row0+rows) against
 * output columns [col0, col0+cols) and writes
 *     D [(r - d_row0)  * ldd + c] = d(r, c)      if D  != NULL   (direct)
 *     DT[(c - dt_row0) * ldt + r] = d(r, c)      if DT != NULL   (transposed)
 * `symmetric` jobs (row range == column range) compute only tiles touching the upper triangle and store
 * r <= c directly and r < c transposed.  DT may point into ANOTHER GPU's memory (peer-mapped, e.g. a
 * torch symmetric-memory buffer): the transposed tile is then pushed over NVLink by the same kernel that
 * computes it, which is how the row-sharded multi-GPU path keeps the factor-2 symmetry saving without a
 * separate exchange step (dist.py).  count_stats: add this job's stored elements to sum / nnz.
 * h_jobs is a HOST array, copied into the launch parameters.
 * operand: int8 centred frames (operand_signed = 1) or raw uint8 frames (0); ld = row pitch in bytes. */
typedef struct AvtexGramJob {
    int64_t row0, rows, col0, cols;
    float *D;
    int64_t d_row0, ldd;
    float *DT;
    int64_t dt_row0, ldt;
    int symmetric, count_stats;
    /* Residue-class jobs (stride-s pipelines, see avtex_diag_filter_pow_res).  A stride-s filter reads D1[i,j] only
     * where i = j (mod s); those entries are the s Gram matrices of the frames of one residue class each.  With the
     * clip viewed as [N/s, s*K] (row a = frames s*a .. s*a+s-1 concatenated; pass n = N/s, ld = s*K, k = K), class r
     * is the K-byte column range starting at k_off = r*K of every row, and the norm of its row a is
     * sqnorm[sq_off + a * sq_stride] with sq_off = r, sq_stride = s.  Zero-initialised fields mean the ordinary
     * job (k_off = 0, norms sqnorm[a]).  K must be a multiple of 128 when k_off is used. */
    int64_t k_off, sq_off, sq_stride;
} AvtexGramJob;
/* clock_probe (nullable, 2 x u64 on the device): CTA 0 writes the SM cycles (clock64) and the wall nanoseconds
 * (globaltimer) its first epilogue warp spent in the tile loop — their ratio is the SM clock the kernel really
 * ran at, which NVML's millisecond-scale sampling cannot see for a ~1 ms kernel (bench.py reports it). */
AVTEX_API int avtex_gram_l2_jobs(const void *operand, int operand_signed, int64_t n, int64_t k, int64_t ld,
                       const int64_t *sqnorm, const AvtexGramJob *h_jobs, int num_jobs, double *sum,
                       unsigned long long *nnz, unsigned long long *clock_probe, int device, void *stream);

/* avtex_gram_l2_jobs with K0 FUSED into the launch (raw uint8 frames, 2-CTA kernel): the epilogue warps of the Gram
 * kernel are idle until the first tiles' MMAs retire (0.43 ms at K = 150528), so they compute the norms meanwhile —
 * `norm_units` byte rows of k bytes, `norm_pitch` bytes apart starting at `frames` (for the residue-class view
 * [N/s, s*k]: norm_units = N, norm_pitch = k), written to sqnorm[0 .. norm_units) and, when max_centred != NULL
 * (zeroed by the caller), the maximum centred norm raised there as avtex_frame_norms_u8 does.  No epilogue reads a
 * norm before every warp of the grid has published its share (`sync_zeroed`: one uint32 the caller sets to 0).
 * Saves the separate pass over the frames (0.11 ms of a 1.5 ms step at C2) and its launch.
 * replaces: the same lines as avtex_frame_norms_u8 + avtex_gram_l2_jobs. */
AVTEX_API int avtex_gram_l2_jobs_fused_norms(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t norm_units,
                                   int64_t norm_pitch, int64_t *sqnorm, unsigned long long *max_centred,
                                   unsigned int *sync_zeroed, const AvtexGramJob *h_jobs, int num_jobs,
                                   double *sum, unsigned long long *nnz, int device, void *stream);

/* sqnorm[i] = sum_c frames[i,c]^2 (exact); max_centred (nullable, zeroed by the caller) receives
 * max_i sum_c (frames[i,c]-128)^2 via atomicMax.  HBM-bound: one read of the frames. */
AVTEX_API int avtex_frame_norms_u8(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t *sqnorm,
                         unsigned long long *max_centred, int device, void *stream);
/* Row-sharded form: `frames` points at global row `row0` of the clip and holds n rows; the norm of row row0+i is
 * written to h_sqnorm[d][row0 + i] for every destination d < num_dst (<= 8) and the centred maximum is raised in
 * h_max_centred[d] (nullable array / entries).  The destinations are the SAME full-length vector on every GPU
 * of the box (peer-mapped pointers, e.g. torch symmetric memory): each rank computes 1/G of the norms and pushes
 * them to its peers from the kernel, replacing an all-gather + all-reduce pair per step (dist.py). */
AVTEX_API int avtex_frame_norms_u8_push(const uint8_t *frames, int64_t n, int64_t k, int64_t ld, int64_t row0,
                              int64_t *const *h_sqnorm, unsigned long long *const *h_max_centred, int num_dst,
                              int device, void *stream);

/* Same contract by direct difference in fp32 (the reference's own formula), for arbitrary float
 * features; SIMT, no tensor cores.  x is [n, k] fp32. */
AVTEX_API int avtex_pairdist_direct_f32(const float *x, int64_t n, int64_t k, int64_t ld,
                              int64_t row0, int64_t rows, float *D, int64_t ldd,
                              double *sum, unsigned long long *nnz, int device, void *stream);
AVTEX_API int avtex_pairdist_direct_u8(const uint8_t *x, int64_t n, int64_t k, int64_t ld,
                             int64_t row0, int64_t rows, float *D, int64_t ldd,
                             double *sum, unsigned long long *nnz, int device, void *stream);

/* ---------------------------------------------------------------- statistics for sigma
 * sum += fp64 sum of D[rows, cols];  nnz += #{D != 0}.
 * replaces: `torch.nonzero(D).size(0)` and `D.sum()` of classic/computeD1.py:240-241,
 * classic/computeD2.py:44-45, classic/q_learning.py:53-54.
 */
AVTEX_API int avtex_sum_nnz(const float *D, int64_t rows, int64_t cols, int64_t ld,
                  double *sum, unsigned long long *nnz, int device, void *stream);

/* ---------------------------------------------------------------- K2: diagonal temporal filter
 * D2[a - a0, b] = sum_{k<fs} w[k] * D1[a*stride + k - in_row0, b*stride + k]   (k ascending, fp32)
 * for a in [a0, a0 + rows_out), b in [0, m).  `D1` points at global row `in_row0` of the distance
 * matrix and holds `in_rows` rows (row shards pass their halo'd block).  If D3 != NULL also writes D3 = powf(D2, p).
 * sum/nnz (nullable) accumulate the stats of the D2 rows written.
 * h_w: HOST pointer to the fs fp32 taps (passed to the kernel by value; fs <= 64 takes the
 * register-resident path, larger fs a generic path).
 * replaces: classic/computeD2.py:34-42 (F.conv2d with diag(binomial)) and classic/q_learning.py:34 (D2 ** p).
 */
AVTEX_API int avtex_diag_filter_pow(const float *D1, int64_t ld1, int64_t in_row0, int64_t in_rows, const float *h_w, int fs,
                          int stride, int64_t a0, int64_t rows_out, int64_t m,
                          float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                          double *sum, unsigned long long *nnz, int device, void *stream);

/* The same filter for a SYMMETRIC D1 (D1[i,j] == D1[j,i] bit for bit, as every matrix written by avtex_gram_l2_* is):
 * D2 is then symmetric too — both mirror images add the same values in the same order — so only the outputs on or
 * above the diagonal are computed and the strictly upper ones are mirrored through shared memory (64-byte runs).
 * Half the D1 bytes (stride 4 is HBM-bound), half the FMAs and pows (stride 1 is FP32-pipe bound); results are
 * bit-identical to avtex_diag_filter_pow on the same input.  Whole matrix only: D1 is [n_rows >= (m-1)*stride+fs, ld1],
 * D2 / D3 are [m, ld].  fs/stride pairs without a register-resident instantiation fall back to the general kernel.
 * replaces: classic/computeD2.py:34-42 + classic/q_learning.py:34 when D1 comes from compute_D1. */
AVTEX_API int avtex_diag_filter_pow_sym(const float *D1, int64_t ld1, int64_t n_rows, const float *h_w, int fs, int stride,
                              int64_t m, float *D2, int64_t ld2, float *D3, int64_t ld3, float p,
                              double *sum, unsigned long long *nnz, int device, void *stream);

/* The same filter over RESIDUE-CLASS planes, for stride >= 2.  A stride-s filter reads D1[i,j] only where
 * i = j (mod s); those entries are the s Gram matrices of the frames of one residue class each (see the k_off / sq_off
 * fields of AvtexGramJob): D1r[r][a, b] = d(frame s*a + r, frame s*b + r), plane r at D1r + r * plane, leading dimension
 * ld1.  In class coordinates the filter is a stride-1 filter of ceil(fs/s) taps per plane: the planes hold CLASS rows
 * [in_row0, in_row0 + in_rows) (row shards pass their halo'd block, like avtex_diag_filter_pow) and at least
 * ceil(((m-1)*s + fs) / s) columns.  The kernel walks a diagonal through the planes round-robin in the same tap order
 * k = 0..fs-1, so D2 / D3 are bit-identical to avtex_diag_filter_pow on the full matrix while K1 computed 1/s of the
 * pairs.  symmetric != 0 (whole matrix only): additionally use the upper-triangle form (every plane of a distance
 * matrix is symmetric).  Register-resident (fs, stride) pairs only: (40,4), (16,4).
 * replaces: classic/computeD1.py:47-96 + classic/computeD2.py:34-42 + classic/q_learning.py:34 when only D2 / D3
 * (not D1 / P1) are consumed, as in classic/video_textures.py:265-284 for -m 3. */
AVTEX_API int avtex_diag_filter_pow_res(const float *D1r, int64_t ld1, int64_t plane, int64_t in_row0, int64_t in_rows,
                              const float *h_w, int fs, int stride, int64_t a0, int64_t rows_out, int64_t m,
                              float *D2, int64_t ld2, float *D3, int64_t ld3, float p, double *sum,
                              unsigned long long *nnz, int symmetric, int device, void *stream);
/* out = D ** p elementwise (D >= 0, p > 0), same pow as the fused epilogue above.
 * replaces: `D3 = D2 ** p` of classic/q_learning.py:34 when D2 is handed in by the caller. */
AVTEX_API int avtex_pow_matrix(const float *D, int64_t ld, int64_t rows, int64_t cols, float p, float *out,
                     int64_t ld_out, int device, void *stream);

/* ---------------------------------------------------------------- K3/K4: future cost
 * One Jacobi sweep in vector form.  For the rows j in [row0, row0+rows) of D3
 * (pointer at row `row0`, m columns):
 *     m_new[j] = min_{k != j} ( D3[j,k] + fl(alpha * m_prev[k]) )     j >= 1   (m_prev == NULL: + 0)
 *     m_new[0] = min_{k != 0}   D3[0,k]
 * and, if eps_sum != NULL, adds  sum_{j>=1,k} ( fl(D3+fl(alpha*m_prev)) - X_prev )^2  with
 * X_prev = fl(D3 + fl(alpha*m_prev2)) (or D3 when m_prev2 == NULL): the numerator of the
 * reference's `eps` for the sweep that produced m_prev.
 * replaces: classic/q_learning.py:39-50.
 */
AVTEX_API int avtex_future_cost_sweep(const float *D3, int64_t ld, int64_t row0, int64_t rows, int64_t m,
                            const float *m_prev, const float *m_prev2, float alpha,
                            float *m_new, double *eps_sum, int device, void *stream);
/* All sweeps in ONE cooperative launch (single GPU): the same arithmetic as repeated
 * avtex_future_cost_sweep calls, with a grid-wide barrier between sweeps and the stop rule
 * `eps > eps_stop` evaluated on the device, so the host reads nothing until the end.
 * mbuf: 3 * mpad floats (mpad >= m, multiple of 4; scratch, rotated).  eps_trail: max_sweeps+1 doubles,
 * ZEROED by the caller; eps_trail[p] receives the numerator of sweep p.  info[0] = number of sweeps
 * (0 if max_sweeps was exhausted), info[1] = index (0..2) of the buffer in mbuf holding the vector m with
 * D3_new = D3 + fl(alpha*m); m_out (nullable, mpad floats) receives a copy of that vector, so the finalize kernel
 * can be launched without a host read of info.  replaces: classic/q_learning.py:39-51 including the while condition. */
AVTEX_API int avtex_future_cost_fused(const float *D3, int64_t ld, int64_t m, float alpha, float eps_stop,
                            int max_sweeps, float *mbuf, int64_t mpad, double *eps_trail, int *info,
                            float *m_out, int device, void *stream);
/* Row-sharded form of avtex_future_cost_fused (one cooperative kernel per GPU, launched on every rank):
 * this rank owns rows [row0, row0+rows) of D3.  After each sweep the kernel pushes its row minima into every
 * rank's m buffer and its eps numerator into every rank's slot array through PEER-MAPPED pointers
 * (h_mbuf[r], h_epsbuf[r], h_flags[r]: rank r's buffers as mapped in this process, e.g. a torch
 * symmetric-memory allocation; 3*mpad floats, (max_sweeps+1)*world doubles, world uint32), then
 * spins on the flag counters its peers write: the per-sweep all-gather and eps all-reduce happen inside the
 * kernel over NVLink.  Flags are monotonic: pass epoch_base = (calls so far) * (max_sweeps + 2), identical on
 * all ranks, and zero the flag/eps buffers once at allocation.  eps_local: [max_sweeps+1] doubles zeroed by
 * the caller (local scratch); eps_trail / info as in avtex_future_cost_fused (info: 4 ints, zeroed; info[2] = 1
 * when a peer did not arrive within ~10 s — the kernel then returns on every rank instead of trapping); the
 * result vector is h_mbuf[rank] + info[1]*mpad and is also copied to m_out (nullable, local, mpad floats).
 * max_ctas != 0 caps the cooperative grid (> 0: that many CTAs; -G: 1/G of full occupancy): "virtual ranks" that
 * share ONE device (tests) must all be resident at once for the flag barrier to close.  world <= 8. */
AVTEX_API int avtex_future_cost_fused_peer(const float *D3, int64_t ld, int64_t row0, int64_t rows, int64_t m,
                                 float alpha, float eps_stop, int max_sweeps, int rank, int world,
                                 float *const *h_mbuf, int64_t mpad, double *const *h_epsbuf,
                                 unsigned int *const *h_flags, unsigned int epoch_base, double *eps_local,
                                 double *eps_trail, int *info, float *m_out, int max_ctas, int device,
                                 void *stream);
/* D3_new[j,:] = D3[j,:] + fl(alpha * mvec)  (j >= 1),  row 0 copied.  sum/nnz nullable.
 * replaces: the materialised D3_new of classic/q_learning.py:48 at convergence. */
AVTEX_API int avtex_future_cost_finalize(const float *D3, int64_t ld, int64_t row0, int64_t rows, int64_t m,
                               const float *mvec, float alpha, float *D3_new, int64_t ld_out,
                               double *sum, unsigned long long *nnz, int device, void *stream);

/* ---------------------------------------------------------------- K5: transition probabilities
 * For output rows i in [0, rows_out): source row r = min(i + shift, rows_in - 1) of D,
 *     P[i,:] = exp(-D[r,:] / sigma) / rowsum,   P_new = (P < max_i - fl(th*max_i)) ? 0 : P.
 * P and P_new are nullable (P_new requires th >= 0).  counts (nullable) receives #nonzero of
 * P_new (or of P when P_new == NULL) per row.
 * replaces: classic/computeD1.py:242-245, classic/computeD2.py:47-50, classic/q_learning.py:56-64.
 */
AVTEX_API int avtex_transition_probs(const float *D, int64_t ld, int64_t rows_in, int64_t cols, float sigma,
                           int shift, int64_t rows_out, float *P, int64_t ldp, float th,
                           float *P_new, int64_t ldn, int *counts, int device, void *stream);

/* Non-zero structure of a matrix (ascending columns inside each row): the survivor lists the
 * sampling walk draws from.  counts[r] = #nonzero(row r);  colidx[rowptr[r] + t] = t-th non-zero.
 * replaces: `P[this_frame].nonzero().view(-1).detach().cpu()` per step,
 * classic/video_textures.py:76-78,149-151,186-188. */
AVTEX_API int avtex_row_nnz(const float *P, int64_t ld, int64_t rows, int64_t cols, int *counts,
                  int device, void *stream);
AVTEX_API int avtex_csr_fill(const float *P, int64_t ld, int64_t rows, int64_t cols, const int64_t *rowptr,
                   int *colidx, int device, void *stream);

/* ---------------------------------------------------------------- K6/K7: contrastive synthesis
 * y[i,:] = x[i,:] / max(||x[i,:]||_2, 1e-12).
 * replaces: nn.functional.normalize of cvt/models/models.py:351,412,433,436. */
AVTEX_API int avtex_l2_normalize_rows(const float *x, int64_t ld, int64_t rows, int64_t dim,
                            float *y, int64_t ldy, int device, void *stream);
/* out[w] = <qn, tn[w,:]> / temp  for w in [0, rows) — one query against every window.
 * replaces: torch.bmm(q, t).squeeze(1) / temp of cvt/models/models.py:416-417 (and :439,457). */
AVTEX_API int avtex_cosine_scores(const float *tn, int64_t ld, int64_t rows, int64_t dim, const float *qn,
                        float temp, float *out, int device, void *stream);
/* One selection step over the target list ids = [pos] ++ ascending({0..L-1} \ {q, pos}),
 * pos = min(q+1, L-1):   o /= sum(o);  a /= sum(a);  v = alpha*o + (1-alpha)*a  (a == NULL: v = o);
 * v[v < max - fl(th*max)] = 0;  v[nz] /= sum(v);  choices = window ids of the non-zeros in ids order.
 * one_minus_alpha is passed separately because the reference evaluates (1 - alpha) in float64.
 * vals (nullable, length L): the final v in window order (entry q undefined unless q == L-1).
 * replaces: cvt/validate.py:369-378, 524-527, 554, 558, 568. */
AVTEX_API int avtex_select_step(const float *o, const float *a, int64_t L, int64_t q, float alpha,
                      float one_minus_alpha, float th, int *choices, int *n_choices, float *vals,
                      int device, void *stream);
/* ONE launch per synthesis step: avtex_cosine_scores on the window table (and on the
 * source-audio table against the driving row when sn/dn != NULL) fused with avtex_select_step — same arithmetic,
 * same survivor list.  The list goes to choices / n_choices on the device AND, without a copy or a stream
 * synchronisation, to MAPPED PINNED host memory: host_out = [seq, n, first host_cap survivors]; the sequence
 * word `seq` is written last (after __threadfence_system), the host polls it.  Workspaces (device): ws_f32
 * 3*L floats; ws_acc 8 doubles and ws_max 4 uint32 (row counter + CTA ticket per parity), ZEROED once before the
 * first step (the kernel re-arms the other parity itself; seq must increase by 1 per call); ws_counts unused.
 * replaces: cvt/models/models.py:351-352,412-417,433-457 + cvt/validate.py:369-378,524-527,554,558,568 per step. */
AVTEX_API int avtex_synthesis_step(const float *tn, int64_t ld, int64_t L, int64_t dim, const float *qn,
                         const float *sn, int64_t lds, int64_t dimA, const float *dn, float temp,
                         int64_t q, float alpha, float one_minus_alpha, float th, float *ws_f32,
                         double *ws_acc, unsigned int *ws_max, int *ws_counts, int ws_counts_len,
                         int *choices, int *n_choices, float *vals, int *host_out, int host_cap,
                         int seq, int device, void *stream);
/* The WHOLE synthesis loop in one cooperative launch: n_steps times { avtex_synthesis_step arithmetic for query
 * q; idx = RandomState.randint(0, n) drawn ON THE DEVICE from numpy's own MT19937 state; q = choices[idx] }.
 * mt_state (device, 625 x 4 bytes: key[624] then pos, as returned by np.random.get_state()) is advanced in place
 * and must be copied back into numpy afterwards (np.random.set_state), so host and device consume ONE random
 * stream and the chosen sequence is bit-identical to the reference's host loop.  qn_table: normalised query rows
 * [L, dim]; dn_table: normalised driving-audio rows, row step+1 is used at step `step` (cvt/validate.py:417).
 * ws_acc: 8 doubles, ZEROED by the caller; q_scratch: one int64.  q_ids / nz [n_steps] receive the chosen window
 * and the survivor count of every step.  No host interaction until the kernel ends.
 * replaces: the `while len(new_frames) < max_length` loop of cvt/validate.py:324-572 at the embedding boundary. */
AVTEX_API int avtex_synthesis_loop(const float *tn, int64_t ld, int64_t L, int64_t dim, const float *qn_table,
                         int64_t ldq, const float *sn, int64_t lds, int64_t dimA, const float *dn_table,
                         int64_t ldd, float temp, float alpha, float one_minus_alpha, float th,
                         int64_t q_start, int n_steps, float *ws_f32, double *ws_acc, int *choices,
                         int *n_choices, void *mt_state, int64_t *q_scratch, int *q_ids, int *nz,
                         int device, void *stream);
/* Host test hook for the device generator: out[i] = RandomState.randint(0, n[i]) drawn from (key[624], *pos),
 * which are advanced in place.  No GPU involved. */
AVTEX_API int avtex_mt19937_randint_host(uint32_t *key, int *pos, const uint32_t *n, int count, uint32_t *out);
/* out[0] = arg max_w <normalize(x[w,:]), normalize(d)> with strict '>' against a running max that
 * starts at 0 (first maximum wins; 0 if no similarity is positive).
 * replaces: the start-segment search of cvt/validate.py:222-240. */
AVTEX_API int avtex_audio_start(const float *x, int64_t ld, int64_t rows, int64_t dim, const float *d,
                      float *sims_ws, int *out, int device, void *stream);

/* ---------------------------------------------------------------- (f1) window construction
 * out[o, :] = rows[idx[o], :] (row_bytes bytes each; idx[o] < 0 or >= n_rows gives a zero row): assembles the
 * [n_windows, window, frame] tensors an encoder consumes straight from the device-resident clip, from an index
 * plan (true windows w*S .. w*S+W, or the reference's per-step chunk layout with its zero padding).
 * replaces: the host-side fancy indexing + chunk copies of cvt/validate.py:380-395, cvt/utils/utils.py:233-260 and
 * the per-window slicing of cvt/models/models.py:355-362. */
AVTEX_API int avtex_gather_rows(const void *rows, int64_t row_bytes, int64_t pitch_bytes, int64_t n_rows,
                      const int *idx, int64_t n_out, void *out, int device, void *stream);

/* ---------------------------------------------------------------- (f4) output frame assembly
 * out[o] = video[ids[o]] (uint8 [h, w, 3] frames); with draw_bar the reference's progress bar is painted over rows
 * [h-25, h-10): black, red (255,0,0) in columns [mark_lo[o], mark_hi[o]) (host-computed with the reference's own
 * slice semantics; an empty range draws no marker).
 * replaces: classic/video_textures.py:215-226, cvt/validate.py:622-634 (one host copy + PIL image per frame). */
AVTEX_API int avtex_assemble_frames(const uint8_t *video, int64_t n_frames, int h, int w, const int *ids,
                          const int *mark_lo, const int *mark_hi, int draw_bar, int64_t n_out, uint8_t *out,
                          int device, void *stream);

/* ---------------------------------------------------------------- (f3) audio front end
 * Log-mel spectrogram in the reference's float64 arithmetic: frame f = samples [f*hop, f*hop + win_len) of `wave`
 * (device, fp64, [n_samples] or [n_samples, channels] averaged to mono) times `window`, zero padded to fft_len,
 * |rfft|, times the mel matrix `mel` ([fft_len/2+1, n_mel] row-major, device fp64), log(. + log_offset) -> out
 * [n_frames, n_mel] fp32.  window / mel are small host-computed tables (contrastive/audio_frontend.py).
 * replaces: cvt/utils/mel_features.py:72-93 (stft_magnitude), :188-223 (log_mel_spectrogram). */
AVTEX_API int avtex_logmel(const double *wave, int64_t n_samples, int channels, int win_len, int hop, int fft_len,
                 const double *window, const double *mel, int n_mel, double log_offset, float *out,
                 int64_t n_frames, int device, void *stream);
/* out[e, t, :] = logmel[e*hop + t, :] for t < win: the VGGish examples.
 * replaces: mel_features.frame on the feature rows, cvt/utils/vggish_utils.py:57-69. */
AVTEX_API int avtex_frame_examples(const float *logmel, int64_t n_frames, int n_mel, int win, int hop, float *out,
                         int64_t n_examples, int device, void *stream);

/* Test hook: the tile visiting order of avtex_gram_l2_s8 for TM x TN tiles (128 x 256).  Returns the
 * number of tiles; fills tm_out/tn_out (capacity entries) when non-NULL.  Host only. */
AVTEX_API int avtex_gram_tile_schedule(int TM, int TN, int symmetric, int *tm_out, int *tn_out, int capacity);
/* Same for the default 2-CTA kernel (256 x 256 tiles; symmetric keeps tn >= tm). */
AVTEX_API int avtex_gram_tile_schedule2(int TM, int TN, int symmetric, int *tm_out, int *tn_out, int capacity);
/* ... with an explicit group size (row-tiles per L2 super-tile: 8 for long K, 16 for K <= 16384). */
AVTEX_API int avtex_gram_tile_schedule2g(int TM, int TN, int symmetric, int group, int *tm_out, int *tn_out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* AVTEX_H_ */
