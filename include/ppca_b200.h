/*
 * ppca_b200.h — C ABI of the B200-native PPCA / PPCA-mixture EM engine.
 *
 * This is the drop-in boundary for the data-parallel hot path of viodotcom/ppca_rs.  The
 * reference has no C ABI of its own: its native surface is the Rust crate API
 * (ppca/src/lib.rs:22-25) wrapped by pyO3 (src/python_bindings.rs:15-26).  Every entry point
 * below names the reference symbol it replaces; a Rust `ppca` host crate would bind them 1:1
 * (see INTEGRATION.md), and the Python package `ppca_rs_b200` binds them with ctypes.
 *
 * Conventions
 *   - Plain pointers and sizes only.  All matrices are row-major f64.
 *   - `C` is the transform, d x k; `mu` the mean, d; `sigma` the isotropic noise STANDARD
 *     DEVIATION (ppca_model.rs:77-80).  Model parameters are tiny host arrays copied per call.
 *   - Datasets live on the device behind an opaque handle.  Input matrices mark missing entries
 *     with any non-finite value (dataset.rs:19-22 mask_non_finite).
 *   - Every function returns 0 on success, non-zero on failure; the message is available from
 *     ppca_b200_last_error() (thread-local).  Nothing throws or aborts across the boundary.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - A context is bound to one device and one stream and is not re-entrant.
 *   - Pointers named *_dev are DEVICE pointers (for the sharded multi-GPU path, where the caller
 *     all-reduces the statistics buffer with NCCL between em_stats and em_finish).
 */
#ifndef PPCA_B200_H
#define PPCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define PPCA_B200_ABI_VERSION 1

typedef struct ppca_b200_ctx ppca_b200_ctx;
typedef struct ppca_b200_dataset ppca_b200_dataset;

/* Status codes */
enum {
  PPCA_OK = 0,
  PPCA_ERR_INVALID = 1,   /* bad argument / shape mismatch (the reference panics: output_covariance.rs:124) */
  PPCA_ERR_CUDA = 2,      /* CUDA runtime error or no device */
  PPCA_ERR_EMPTY = 3,     /* empty dataset where the reference asserts (ppca_model.rs:52,358) */
  PPCA_ERR_NUMERIC = 4,   /* singular system where the reference `expect`s (output_covariance.rs:69, prior.rs:109) */
  PPCA_ERR_WEIGHTS = 5,   /* mixture iterate needs strictly positive weights (mix.rs:304-309,326) */
  PPCA_ERR_PRECISION = 6  /* ppca_b200_em_finish only: the precision guard of the int8-sliced contractions fired on the
                             (reduced) statistics and the context has moved to a wider arithmetic; repeat
                             em_stats (+ the all-reduce) and em_finish.  Single-call entry points retry internally. */
};

/* prior.rs:8-29 Prior.  mean_precision is the inverse of the prior mean covariance (prior.rs:36-41),
 * computed by the caller.  Pointers may be NULL when the corresponding flag is 0. */
typedef struct {
  int32_t has_mean_prior;
  const double *mean;            /* d */
  const double *mean_precision;  /* d x d */
  int32_t has_isotropic_noise_prior;
  double isotropic_noise_alpha;
  double isotropic_noise_beta;
  double transformation_precision;
} ppca_b200_prior;

/* ---- library ------------------------------------------------------------------------------- */
int32_t ppca_b200_abi_version(void);
const char *ppca_b200_last_error(void);
int32_t ppca_b200_device_count(int32_t *out);

/* ---- context ------------------------------------------------------------------------------- */
/* `cuda_stream` is a cudaStream_t to run on (e.g. torch's current stream) or NULL to create one. */
int32_t ppca_b200_ctx_create(int32_t device, void *cuda_stream, ppca_b200_ctx **out);
int32_t ppca_b200_ctx_destroy(ppca_b200_ctx *ctx);
int32_t ppca_b200_ctx_synchronize(ppca_b200_ctx *ctx);
/* The cudaStream_t every call of this context is enqueued on (the one passed to ctx_create, or the context's own).
 * Asynchronous entry points (ppca_b200_em_stats*, ppca_b200_mix_em_stats) return while their kernels run: foreign
 * work that touches their buffers (an NCCL all-reduce of the statistics) must be enqueued on, or ordered with, it. */
int32_t ppca_b200_ctx_stream(ppca_b200_ctx *ctx, void **out);
/* Tuning knob: samples per E/M-step chunk (0 = automatic). */
int32_t ppca_b200_ctx_set_chunk(ppca_b200_ctx *ctx, int64_t chunk_samples);
/* Arithmetic path of the two masked-Gram contractions (E-step Gram matrices, M-step second moments):
 *   mode 0 = FP64 tensor cores (mma.sync DMMA);
 *   mode 1 = exact int8-sliced evaluation on the int8 tensor path: the {0,1} mask times `slices` balanced base-256
 *            digit planes (int8) of the FP64 operand, int32 accumulation (exact), FP64 recombination.  slices in
 *            {6,7,8} keep every term to 46 / 54 / 62 bits below its column scale (6: below FP64 dot-product
 *            rounding; 7: every FP64 input exactly).  mma.sync IMMA.
 *   mode 2 = the same arithmetic on tcgen05.mma.kind::i8 with the accumulators in tensor memory.  slices = 4 is
 *            accepted in this mode only: the opt-in FP32-class FAST PATH (30 bits below each column scale, ~1e-9 per
 *            term: iterates agree with the FP64 path to ~1e-6, inside the 1e-4 the FP32 path is allowed) at two
 *            thirds of the tensor work and digit-plane bytes.
 * Default: mode 2 with 6 slices; the environment overrides it: PPCA_B200_GEMM=dmma|int8|tc, PPCA_B200_SLICES=6|7|8. */
int32_t ppca_b200_ctx_set_gemm(ppca_b200_ctx *ctx, int32_t mode, int32_t slices);
/* Precision guard of the int8-sliced modes (on by default; never active in mode 0 or with 4 slices).  Every term of
 * those contractions is kept to 8 slices - 2 bits below the largest entry of its COLUMN, so a sample (E-step) or an output
 * dimension (M-step) that only sees entries far below that maximum loses relative accuracy.  The engine bounds the loss
 * against the diagonal of the matrix it perturbs - terms(d_n) s_aa 2^-(8 slices - 1) <= 2^-eps_bits (sigma^2 + G_n[a][a])
 * per sample, the analogue with sum_n m_ni W_n[a][a] per output dimension - and, when any entry fails, repeats the whole
 * pass one rung up the ladder: configured slices -> 8 slices -> mode 0 (FP64 DMMA, the reference's own arithmetic).  The
 * rung that was needed is kept for the context's next 16 passes.  eps_bits = 0 keeps the current value (default 40).
 * Counter 13 of ppca_b200_ctx_variant_counts counts the repeated passes. */
int32_t ppca_b200_ctx_set_guard(ppca_b200_ctx *ctx, int32_t enabled, int32_t eps_bits);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int32_t ppca_b200_ctx_launch_count(ppca_b200_ctx *ctx, int64_t *out);
/* Launches so far per shape-dependent kernel variant, 16 counters (tests assert which code path a case ran):
 *   0 tbitgemm_kernel (A tile in shared memory, T = 7, 8)   1 tbitgemm_atm_kernel<T,2,1>   2 tbitgemm_atm_kernel<T,1,2>
 *   3 tbitgemm_atm2_kernel (two output tiles per mask stage)  4 ibitgemm_kernel (IMMA)      5 bitgemm_kernel (DMMA)
 *   6-8 solve_reg_kernel<8|16|32>   9 unused (kernel removed)   10 solve_tile_kernel (register-tiled, 32 < k <= 64)
 *   11 unused
 *   12 solve_kernel (generic)   13 passes repeated at a wider arithmetic by the precision guard
 *   14 batched mixture contraction launches   15 mixture chunk loops replayed from a captured CUDA graph */
int32_t ppca_b200_ctx_variant_counts(ppca_b200_ctx *ctx, int64_t *out16);
/* Device time accumulated since profiling was enabled (ppca_b200_ctx_set_profiling(ctx, 1) resets it), broken
 * down per kernel family, in ms; reading it synchronises the stream once, the profiled calls never do:
 * out[0]=model staging (Ksym, digit planes) out[1]=gram (masked contraction, E-step) out[2]=proj out[3]=solve
 * out[4]=moment (masked contraction, M-step) out[5]=cross+resid out[6]=finish out[7]=digit-plane slicing of W.
 * Measured with CUDA events recorded on the context's stream around each family's launches. */
int32_t ppca_b200_ctx_set_profiling(ppca_b200_ctx *ctx, int32_t enabled);
int32_t ppca_b200_ctx_last_profile(ppca_b200_ctx *ctx, double *out8);

/* ---- datasets: Dataset / MaskedSample / Mask (dataset.rs:11-14,93-100; utils.rs:27-28) --------- */
/* Dataset::new / new_with_weights from a host matrix (src/python_bindings.rs:32-64). weights may be NULL (all 1). */
int32_t ppca_b200_dataset_from_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d,
                                    const double *weights, ppca_b200_dataset **out);
/* PPCAMix::sample (mix.rs:124-134): every draw picks a component from exp(log_weights) (WeightedIndex) and samples that
 * model (sample_one, ppca_model.rs:164-191), on the device.  Cs = the d x k_j transforms back to back, mus = m x d. */
int32_t ppca_b200_mix_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t m, const int32_t *ks, const double *Cs,
                             const double *mus, const double *sigmas, const double *log_weights, double mask_prob,
                             uint64_t seed, ppca_b200_dataset **out);
/* PosteriorSampler / PosteriorSamplerMix (ppca_model.rs:581-626, mix.rs:505-532): one draw per inferred sample,
 * x = C_j (state + L xi) + mu_j + sigma_j eps with L L^T = covariance (Cholesky on the device; PPCA_ERR_NUMERIC where the
 * reference's `expect("Cholesky decomposition failed")` fires), component j drawn from the row's posterior probabilities
 * (m > 1; posteriors is n x m, unnormalised weights are accepted as by WeightedIndex).  states[j]: n x k_j,
 * covariances[j]: n x k_j x k_j, host arrays as returned by ppca_b200_infer.  Output samples are fully observed. */
int32_t ppca_b200_posterior_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t m, const int32_t *ks, const double *Cs,
                                   const double *mus, const double *sigmas, const double *posteriors,
                                   const double *const *states, const double *const *covariances, uint64_t seed,
                                   ppca_b200_dataset **out);
/* Dataset::new from a matrix that is ALREADY ON THE DEVICE (DLPack / __cuda_array_interface__ producers): no host round
 * trip; replaces the element-by-element numpy -> Rust copy of src/python_bindings.rs:41-54 for device-resident callers.
 * x: row-major, row_stride doubles between rows (>= d); weights: device pointer or NULL; producer_stream: the
 * cudaStream_t (as void*) the caller last wrote x / weights on, the ingest is ordered after it (NULL = default stream).
 * The dataset keeps its own packed copy: x may be freed when the call returns. */
int32_t ppca_b200_dataset_from_device(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, int64_t row_stride,
                                      const double *weights, void *producer_stream, ppca_b200_dataset **out);
/* Dataset::numpy (dataset.rs:64-72) into caller-provided device memory: nrows x d doubles, NaN at the masked slots. */
int32_t ppca_b200_dataset_to_device(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0, int64_t nrows,
                                    double *out_dev);
/* Synthetic data generated on the device with the reference's sampler semantics
 * (ppca_model.rs:164-191 sample_one): x = C_true xi + sigma_true eps, each entry masked with prob mask_prob.
 * C_true[i,a] ~ Bernoulli(0.1) as in examples/big_toy_model.py:6; n_components > 1 draws each sample
 * from one of n_components such models (PPCAMix::sample, mix.rs:124-134) with uniform mixing. */
int32_t ppca_b200_dataset_synthetic(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k_true,
                                    double sigma_true, double mask_prob, int32_t n_components,
                                    uint64_t seed, ppca_b200_dataset **out);
/* Rows [row_begin, row_begin + n) of the same synthetic dataset (the generator's counters are keyed by the global row):
 * the ranks of a sharded job pass one seed and their own row range and hold disjoint parts of ONE dataset. */
int32_t ppca_b200_dataset_synthetic_rows(ppca_b200_ctx *ctx, int64_t row_begin, int64_t n, int32_t d, int32_t k_true,
                                         double sigma_true, double mask_prob, int32_t n_components,
                                         uint64_t seed, ppca_b200_dataset **out);
/* PPCAModel::sample (ppca_model.rs:164-191): n draws x = C xi + mu + sigma eps from the given model, each entry masked
 * with probability mask_prob, generated on the device with a counter-based RNG keyed by `seed` (the reference is
 * unseeded: only the distribution can agree). */
int32_t ppca_b200_model_sample(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, const double *mu,
                               double sigma, double mask_prob, uint64_t seed, ppca_b200_dataset **out);
/* Dataset::with_weights (dataset.rs:171-176): shares the samples, new weights (host array of len n). */
int32_t ppca_b200_dataset_with_weights(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds,
                                       const double *weights, ppca_b200_dataset **out);
/* Dataset::len / output_size (dataset.rs:179-191) */
int32_t ppca_b200_dataset_len(const ppca_b200_dataset *ds, int64_t *out);
int32_t ppca_b200_dataset_output_size(const ppca_b200_dataset *ds, int32_t *out);
/* Dataset.numpy() (src/python_bindings.rs:81-92; dataset.rs:64-72 masked_vector): rows [row0,row0+nrows)
 * into out (nrows x d), NaN at masked slots. */
int32_t ppca_b200_dataset_to_host(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0,
                                  int64_t nrows, double *out);
/* Dataset.weights() (src/python_bindings.rs:106-108) */
int32_t ppca_b200_dataset_weights(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, double *out);
/* Dataset::empty_dimensions (dataset.rs:194-222): out[i] = 1 if dimension i is masked in every sample. */
int32_t ppca_b200_dataset_empty_dimensions(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, uint8_t *out);
/* DatasetChunks::__next__ (src/python_bindings.rs:151-165): copy of rows [row0,row0+nrows) with their weights. */
int32_t ppca_b200_dataset_slice(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int64_t row0,
                                int64_t nrows, ppca_b200_dataset **out);
/* Dataset.concat (src/python_bindings.rs:118-133) */
int32_t ppca_b200_dataset_concat(ppca_b200_ctx *ctx, const ppca_b200_dataset *const *list, int32_t count,
                                 ppca_b200_dataset **out);
int32_t ppca_b200_dataset_destroy(ppca_b200_dataset *ds);

/* ---- PPCAModel (ppca/src/ppca_model.rs) -------------------------------------------------------- */
/* PPCAModel::llks (ppca_model.rs:152-159): out has n entries. */
int32_t ppca_b200_llks(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                       const double *mu, double sigma, double *out);
/* PPCAModel::llk (ppca_model.rs:142-149): weighted sum. */
int32_t ppca_b200_llk(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                      const double *mu, double sigma, double *out);
/* PPCAModel::infer (ppca_model.rs:195-227): states n x k, covariances n x k x k (nullable). */
int32_t ppca_b200_infer(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                        const double *mu, double sigma, double *states, double *covariances);
/* InferredMasked::smoothed_covariance_diagonal (ppca_model.rs:485-508) and ::extrapolated_covariance_diagonal
 * (:542-577), batched as in src/python_bindings.rs:282-333: out[n][i] = sigma^2 + c_i^T Sigma_n c_i, evaluated as
 * the dense FP64 GEMM (packed Sigma_n) x (row-wise symmetric Kronecker table of C)^T on the tensor cores.
 * `covariances` is the host n x k x k array ppca_b200_infer returned.  With `masked_by` (nullable) the slots that
 * dataset observed are 0, as in the extrapolated variant.  Result: an all-observed dataset, weights 1. */
int32_t ppca_b200_covariance_diagonal(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, double sigma,
                                      const double *covariances, const ppca_b200_dataset *masked_by,
                                      ppca_b200_dataset **out);
/* InferredMasked::smoothed_covariance (masked_by NULL) / extrapolated_covariance (ppca_model.rs:471-477, 517-534): the full
 * d x d matrices sigma^2 I + C Sigma_n C^T, computed on the device into out (host, n x d x d); with masked_by the rows and
 * columns of the dimensions sample n observed are zero (all zeros when nothing is missing).  The reference warns about the
 * size (:466-470): n d^2 doubles — prefer ppca_b200_covariance_diagonal for anything large. */
int32_t ppca_b200_covariance_full(ppca_b200_ctx *ctx, int64_t n, int32_t d, int32_t k, const double *C, double sigma,
                                  const double *covariances, const ppca_b200_dataset *masked_by, double *out);
/* PPCAModel::smooth (ppca_model.rs:237-244) / ::extrapolate (:254-261): new all-observed dataset,
 * weights carried through. */
int32_t ppca_b200_smooth(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                         const double *mu, double sigma, ppca_b200_dataset **out);
int32_t ppca_b200_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                              const double *mu, double sigma, ppca_b200_dataset **out);
/* smooth (extrapolate = 0) / extrapolate (= 1) and, from the SAME E-step, PPCAModel::llks (ppca_model.rs:152-159) into
 * `llks` (host, n entries, nullable).  `reuse` (nullable) is a dataset an earlier smooth / extrapolate / reconstruct call
 * returned for an input of the same shape: its storage is overwritten in place (no allocation, no mask rebuild) and
 * *out == reuse; with reuse == NULL a new dataset is created.  This is what a streaming inference loop calls per batch:
 * the reference runs the per-sample algebra once for extrapolate and once more for llks. */
int32_t ppca_b200_reconstruct(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                              const double *mu, double sigma, int32_t extrapolate, ppca_b200_dataset *reuse,
                              double *llks, ppca_b200_dataset **out);
/* PPCAModel::iterate_with_prior (ppca_model.rs:277-393); prior may be NULL (= iterate, :267-269).
 * llk_in (nullable) receives the log-likelihood of the INPUT model on ds, which the E-step produces
 * for free (what PPCATrainer prints each iteration, python/ppca_rs/__init__.py:51). */
int32_t ppca_b200_iterate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                          const double *mu, double sigma, const ppca_b200_prior *prior, double *C_out,
                          double *mu_out, double *sigma_out, double *llk_in);

/* ---- out-of-core EM step: the samples stay in HOST memory ---------------------------------------- */
/* Same result as ppca_b200_dataset_from_host + ppca_b200_iterate (ppca_model.rs:277-393; the numpy -> Rust copy of
 * src/python_bindings.rs:41-54 happens per step instead of once), for datasets that do not fit the device or are
 * visited once: `x` (n x d row-major f64, non-finite = missing) and `weights` (n, nullable = 1) are streamed block
 * by block, the H2D copy of block i+1 on a second stream overlapping the ingest + E/M-step kernels of block i.
 * Page-lock the buffers first (ppca_b200_host_register, or cudaHostAlloc) for full PCIe bandwidth; pageable
 * memory works but copies synchronously through the driver's staging buffer. */
int32_t ppca_b200_iterate_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                               int32_t k, const double *C, const double *mu, double sigma,
                               const ppca_b200_prior *prior, double *C_out, double *mu_out, double *sigma_out,
                               double *llk_in);
/* Compact host format for the out-of-core path: only the OBSERVED values cross the bus (c2: -19 % bytes, c3: -29 %).
 *   vals   : the finite entries of x, row-major (rowptr[n] doubles)
 *   rowptr : n + 1 offsets into vals (rowptr[0] = 0)
 *   maskw  : n x ceil(d/32) mask words, bit b of word j <-> dimension 32 j + b (the layout of bit-vec's BitVec<u32>,
 *            utils.rs:27-28: a Rust host can hand its Mask storage over as it is)
 * ppca_b200_pack_host builds them from an n x d matrix on the host (non-finite = missing, dataset.rs:19-22); with
 * vals == NULL or maskw == NULL it only fills rowptr (rowptr[n] = number of values to allocate).  Page-lock the three
 * arrays (ppca_b200_host_register) for full PCIe bandwidth.  Same results as ppca_b200_iterate_host, bit for bit. */
int32_t ppca_b200_pack_host(const double *x, int64_t n, int32_t d, double *vals, int64_t *rowptr, uint32_t *maskw);
int32_t ppca_b200_iterate_packed_host(ppca_b200_ctx *ctx, const double *vals, const int64_t *rowptr,
                                      const uint32_t *maskw, int64_t n, int32_t d, const double *weights, int32_t k,
                                      const double *C, const double *mu, double sigma, const ppca_b200_prior *prior,
                                      double *C_out, double *mu_out, double *sigma_out, double *llk_in);
/* The sharded form of the same: this rank's host-resident rows into stats_dev (DEVICE, layout of
 * ppca_b200_em_stats_len), to be all-reduced by the caller and finished with ppca_b200_em_finish. */
int32_t ppca_b200_em_stats_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                                int32_t k, const double *C, const double *mu, double sigma, double *stats_dev);
/* Out-of-core inference: PPCAModel::smooth (extrapolate = 0) / ::extrapolate (= 1) (ppca_model.rs:237-261) and
 * ::llks (:152-159) over host-resident samples, streamed like ppca_b200_iterate_host; three streams overlap the H2D
 * of block i+1, the kernels of block i and the D2H of block i-1.  `out` (n x d, nullable) receives the
 * reconstruction (observed slots of extrapolate are bit copies of x), `llks` (n, nullable) the log-likelihoods. */
int32_t ppca_b200_reconstruct_host(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, int32_t k,
                                   const double *C, const double *mu, double sigma, int32_t extrapolate, double *out,
                                   double *llks);
/* cudaHostRegister / cudaHostUnregister of a caller-owned host range (e.g. a numpy array). */
int32_t ppca_b200_host_register(const void *p, uint64_t bytes);
int32_t ppca_b200_host_unregister(const void *p);

/* ---- sharded EM: one process per GPU, statistics all-reduced by the caller --------------------- */
/* Length (in doubles) of the additive sufficient-statistics buffer for (d, k):
 *   [ A: d x kkp | B: d x kp | tdev: d | totals: d | 8 scalars ]   kkp = roundup8(k(k+1)/2), kp = roundup8(k)
 * scalars: 0 = sum w (tr(C_o Sigma_n C_o^T) + |dev_n|^2)  (square_error + deviations_square_sum, ppca_model.rs:345-346; the
 *          residual norm is |x~|^2 - y^T z - sigma^2 |z|^2, never a pass over the residual itself), 1 = 0 (kept for
 *          layout compatibility), 2 = sum w llk_n, 3 = sum w (all samples), 4 = number of non-empty samples,
 *          5 = E-step and 6 = M-step precision-guard violations of this shard (see ppca_b200_ctx_set_guard), 7 reserved. */
int64_t ppca_b200_em_stats_len(int32_t d, int32_t k);
/* E-step + local M-step statistics of this shard (ppca_model.rs:278-358) into stats_dev (DEVICE memory,
 * overwritten).  Asynchronous on the context's stream. */
int32_t ppca_b200_em_stats(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                           const double *mu, double sigma, double *stats_dev);
/* M-step finish from (all-reduced) statistics: d row solves, sigma^2, mu, prior hooks
 * (ppca_model.rs:294-324 solve part, :360-392). */
int32_t ppca_b200_em_finish(ppca_b200_ctx *ctx, int32_t d, int32_t k, const double *C, const double *mu,
                            double sigma, const ppca_b200_prior *prior, const double *stats_dev,
                            double *C_out, double *mu_out, double *sigma_out, double *llk_in);

/* ---- PPCAMix (ppca/src/mix.rs) ------------------------------------------------------------------ */
/* Mixture parameters: m models; ks[j] state sizes; Cs = concatenation of the C_j (each d x ks[j]);
 * mus = m x d; sigmas = m; log_weights = m (already normalised, mix.rs:69). */
/* PPCAMix::llks (mix.rs:152-159) */
int32_t ppca_b200_mix_llks(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                           const double *Cs, const double *mus, const double *sigmas,
                           const double *log_weights, double *out);
/* PPCAMix::llk (mix.rs:162-174) */
int32_t ppca_b200_mix_llk(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                          const double *Cs, const double *mus, const double *sigmas,
                          const double *log_weights, double *out);
/* PPCAMix::infer_cluster (mix.rs:179-189): n x m log-posteriors. */
int32_t ppca_b200_mix_infer_cluster(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m,
                                    const int32_t *ks, const double *Cs, const double *mus,
                                    const double *sigmas, const double *log_weights, double *out);
/* PPCAMix::smooth / ::extrapolate (mix.rs:245-265): weights reset to 1. */
int32_t ppca_b200_mix_smooth(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                             const double *Cs, const double *mus, const double *sigmas,
                             const double *log_weights, ppca_b200_dataset **out);
int32_t ppca_b200_mix_extrapolate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m,
                                  const int32_t *ks, const double *Cs, const double *mus,
                                  const double *sigmas, const double *log_weights, ppca_b200_dataset **out);
/* PPCAMix::iterate_with_prior (mix.rs:281-337).  Outputs are laid out like the inputs;
 * llk_in (nullable) receives the mixture log-likelihood of the INPUT model. */
int32_t ppca_b200_mix_iterate(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                              const double *Cs, const double *mus, const double *sigmas,
                              const double *log_weights, const ppca_b200_prior *prior, double *Cs_out,
                              double *mus_out, double *sigmas_out, double *log_weights_out, double *llk_in);

/* Sharded mixture EM.  Step 1: log-posteriors of this shard into logpost_dev (n x m, DEVICE) and the
 * per-component local maxima of ln w_n + logpost[n][j] into comp_max (host, m; -inf if the shard is empty);
 * also the weighted mixture llk of the shard (nullable).  The caller all-reduces comp_max with MAX. */
int32_t ppca_b200_mix_posteriors(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m,
                                 const int32_t *ks, const double *Cs, const double *mus,
                                 const double *sigmas, const double *log_weights, double *logpost_dev,
                                 double *comp_max, double *llk_in);
/* Step 2, per component j: statistics with responsibilities r_n = exp(ln w_n + logpost[n][j] - comp_max_j)
 * as weights (mix.rs:304-326) into stats_dev (layout of ppca_b200_em_stats_len(d, ks[j])); scalar 3 of the
 * buffer is sum_n r_n.  The caller all-reduces with SUM and calls ppca_b200_em_finish per component. */
int32_t ppca_b200_mix_em_stats(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, int32_t j,
                               int32_t k, const double *C, const double *mu, double sigma,
                               const double *logpost_dev, double comp_max_j, double *stats_dev);

/* ---- sample-sharded EM with the collective inside the library -------------------------------------------------------
 * One process (or thread) per GPU, each with its own context and its own shard of the samples; the model is replicated.
 * The reference has no counterpart (its only parallelism is rayon over samples, ppca_model.rs:281-358): these entry
 * points are what a Rust `ppca` host binds to get the north star's "one NCCL all-reduce over NVLink per iteration".
 * Rank 0 creates a unique id and ships its 128 bytes to the other ranks through any channel it has (file, socket,
 * MPI, torch.distributed ...); every rank then calls ppca_b200_comm_init.  NCCL itself is bound at run time
 * (dlopen libnccl.so.2; PPCA_B200_NCCL_LIB overrides the name). */
#define PPCA_B200_UNIQUE_ID_BYTES 128
int32_t ppca_b200_comm_unique_id(uint8_t *out /* PPCA_B200_UNIQUE_ID_BYTES */);
int32_t ppca_b200_comm_init(ppca_b200_ctx *ctx, const uint8_t *unique_id, int32_t rank, int32_t world);
int32_t ppca_b200_comm_destroy(ppca_b200_ctx *ctx);
/* In-place all-reduce of `count` doubles (DEVICE memory) on the context's stream; op 0 = sum, 1 = max. */
int32_t ppca_b200_comm_allreduce(ppca_b200_ctx *ctx, double *buf_dev, int64_t count, int32_t op);
/* PPCAModel::iterate_with_prior (ppca_model.rs:277-393) over the union of all ranks' shards: local statistics, ONE
 * all-reduce(sum) of [A | B | tdev | totals | 8 scalars], M-step finish replicated on every rank (identical outputs).
 * A rank may hold an empty shard.  The precision ladder climbs on every rank together (the guard counters are part of
 * the reduced buffer). */
int32_t ppca_b200_iterate_sharded(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t k, const double *C,
                                  const double *mu, double sigma, const ppca_b200_prior *prior, double *C_out,
                                  double *mu_out, double *sigma_out, double *llk_in);
/* One EM step (ppca_model.rs:277-393) over rows [row_begin, row_begin + nrows) of the synthetic dataset that
 * ppca_b200_dataset_synthetic(n >= row_begin + nrows, d, k_true, sigma_true, mask_prob, n_components = 1, seed) would hold,
 * WITHOUT storing them: every chunk is regenerated on the device and consumed (the out-of-core form of BASELINE
 * configs[2], N = 100 M x 2048 = 1.6 TB, which fits no set of 8 GPUs).  sharded != 0: the statistics are all-reduced over
 * the context's communicator before the finish (every rank passes its own row range and gets the same model). */
int32_t ppca_b200_iterate_generated(ppca_b200_ctx *ctx, int64_t row_begin, int64_t nrows, int32_t d, int32_t k_true,
                                    double sigma_true, double mask_prob, uint64_t seed, int32_t k, const double *C,
                                    const double *mu, double sigma, const ppca_b200_prior *prior, int32_t sharded,
                                    double *C_out, double *mu_out, double *sigma_out, double *llk_in);
/* The same with this rank's samples in HOST memory, streamed every step (see ppca_b200_iterate_host). */
int32_t ppca_b200_iterate_host_sharded(ppca_b200_ctx *ctx, const double *x, int64_t n, int32_t d, const double *weights,
                                       int32_t k, const double *C, const double *mu, double sigma,
                                       const ppca_b200_prior *prior, double *C_out, double *mu_out, double *sigma_out,
                                       double *llk_in);
/* ... and in the compact host format (see ppca_b200_iterate_packed_host). */
int32_t ppca_b200_iterate_packed_host_sharded(ppca_b200_ctx *ctx, const double *vals, const int64_t *rowptr,
                                              const uint32_t *maskw, int64_t n, int32_t d, const double *weights,
                                              int32_t k, const double *C, const double *mu, double sigma,
                                              const ppca_b200_prior *prior, double *C_out, double *mu_out,
                                              double *sigma_out, double *llk_in);
/* PPCAMix::iterate_with_prior (mix.rs:281-337) over all ranks' shards: all-reduce(max) of the per-component
 * responsibility maxima (mix.rs:312-318), all-reduce(sum) of the statistics of every component. */
int32_t ppca_b200_mix_iterate_sharded(ppca_b200_ctx *ctx, const ppca_b200_dataset *ds, int32_t m, const int32_t *ks,
                                      const double *Cs, const double *mus, const double *sigmas,
                                      const double *log_weights, const ppca_b200_prior *prior, double *Cs_out,
                                      double *mus_out, double *sigmas_out, double *log_weights_out, double *llk_in);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PPCA_B200_H */
