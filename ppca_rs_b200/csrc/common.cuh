// common.cuh — shared definitions for the B200 PPCA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ppca_b200.h"

namespace ppca {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
struct Error {
  int code;
  std::string msg;
};
#define PPCA_THROW(code_, ...)                               \
  do {                                                       \
    char buf_[512];                                          \
    snprintf(buf_, sizeof(buf_), __VA_ARGS__);               \
    throw ::ppca::Error{(code_), std::string(buf_)};         \
  } while (0)
#define CUDA_CHECK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      PPCA_THROW(PPCA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                 __LINE__);                                                                     \
  } while (0)
#define REQUIRE(cond, ...) \
  do {                     \
    if (!(cond)) PPCA_THROW(PPCA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// shapes
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int tri(int k) { return k * (k + 1) / 2; }
// packed upper-by-rows index of (a, b), a <= b, in a k x k symmetric matrix
__host__ __device__ inline int tri_row_off(int a, int k) { return a * k - (a * (a - 1)) / 2; }
__host__ __device__ inline int tri_idx(int a, int b, int k) { return tri_row_off(a, k) + (b - a); }

struct Shape {
  int d, k;
  int kk;    // k(k+1)/2
  int kkp;   // kk rounded up to 8 (n8 MMA tiles, 64-byte rows)
  int kp;    // k rounded up to 8
  int d32;   // d rounded up to 32 (mask words)
  __host__ __device__ Shape() {}
  __host__ __device__ Shape(int d_, int k_) : d(d_), k(k_) {
    kk = tri(k_);
    kkp = (int)round_up(kk > 0 ? kk : 1, 8);
    kp = (int)round_up(k_ > 0 ? k_ : 1, 8);
    d32 = (int)round_up(d_, 32);
  }
};

// statistics buffer layout (see ppca_b200_em_stats_len)
struct StatsLayout {
  int64_t offA, offB, offTdev, offTotals, offScalars, len;
  StatsLayout(int d, int k) {
    Shape s(d, k);
    offA = 0;
    offB = offA + (int64_t)d * s.kkp;
    offTdev = offB + (int64_t)d * s.kp;
    offTotals = offTdev + d;
    offScalars = offTotals + d;
    len = offScalars + 8;
  }
};
enum { SC_SQERR = 0, SC_DEV2 = 1, SC_LLK = 2, SC_SUMW = 3, SC_NONEMPTY = 4, SC_UNSAFE_E = 5, SC_UNSAFE_M = 6 };

// ---------------------------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t count = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    count = 0;
  }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 1;
    CUDA_CHECK(cudaMalloc((void **)&p, n * sizeof(T)));
    count = n;
  }
  // grow-only
  void reserve(size_t n) {
    if (n > count) alloc(n);
  }
};

// immutable sample storage shared between datasets that differ only in weights
struct SampleStore {
  int64_t n = 0;       // samples
  int d = 0;           // output size
  int ldx = 0;         // row stride of X in doubles (multiple of 4)
  int dw = 0;          // mask words per sample = ceil(d/32)
  int64_t n_pad = 0;   // samples rounded up to 256 (rows of mask are allocated/zeroed up to here)
  int64_t nwT = 0;     // words per dimension row of maskT = n_pad / 32
  int d_pad = 0;       // rows of maskT, d rounded up to 256, zeroed
  DevBuf<double> X;        // n_pad x ldx ; masked slots and padding hold 0.0
  DevBuf<uint32_t> mask;   // n_pad x dw  ; bit b of word j <-> dimension 32 j + b (LSB first, like bit-vec)
  DevBuf<uint32_t> maskT;  // d_pad x nwT ; bit b of word j <-> sample 32 j + b
  DevBuf<int> dn;          // n_pad observed counts
  bool full = false;       // every slot observed (an output of smooth / extrapolate): may be overwritten in place
};

}  // namespace ppca

struct ppca_b200_dataset {
  std::shared_ptr<ppca::SampleStore> store;
  ppca::DevBuf<double> w;  // n_pad weights (padding 0)
  double min_w = 1.0;      // smallest weight (mixture EM needs > 0, mix.rs:304-309)
  int device = 0;
};

namespace ppca {

// Chunk workspaces and accumulators of one model.  A single model uses the context's own buffers; a mixture pass keeps
// one set per component alive at the same time (api.cu, mix_em_pass).
struct ModelWs {
  double *GW = nullptr, *YZ = nullptr, *WZ = nullptr, *nx = nullptr, *llk = nullptr, *tn = nullptr;
  int8_t *KsymQ = nullptr;
  double *KsymScale = nullptr;
  int8_t *WQ = nullptr;
  double *WScale = nullptr, *WScaleMax = nullptr;
  unsigned long long *colmax = nullptr;
  double *part_bg = nullptr, *part_cr = nullptr, *part_solve = nullptr;
  // Mask^T (w Z): the second M-step contraction (d x kp), its digit planes, scales and split-K partials
  int8_t *ZQ = nullptr;
  double *ZScale = nullptr, *MZ = nullptr, *part_mz = nullptr;
  unsigned long long *zcolmax = nullptr;
  double *dv = nullptr;       // per-sample |R_n|^2 of the chunk (identity form)
  int *rflag = nullptr;       // chunk flag: take the residual norms from the exact pass instead
  double *part_rx = nullptr;  // partial slots of the exact residual pass
  double *rscratch = nullptr; // SOLVE_SCRATCH doubles (block totals + completion counter of solve_reduce_kernel)
};

// per-iteration device model
struct DevModel {
  Shape s;
  double sigma;
  const double *C;     // d32 x kp   (zero padded)
  const double *mu;    // d32
  const double *Ksym;  // d32 x kkp  (zero padded rows and columns)
  const ModelWs *ws = nullptr;  // null = the context's buffers
  const double *sigma_dev = nullptr;  // device copy of sigma (mixture passes: lets the chunk loop replay from a CUDA graph)
};

// Kernel variants whose selection depends on the shape: counted per context so that tests can assert which code path
// a case actually ran (ppca_b200_ctx_variant_counts).
enum Variant {
  V_TC_SMEM_A = 0,   // tbitgemm_kernel<T>            (A tile in shared memory; T = 7, 8)
  V_TC_ATM_FEED,     // tbitgemm_atm_kernel<T, 2, 1>  (A tile in tensor memory, two producer groups)
  V_TC_ATM_DRAIN,    // tbitgemm_atm_kernel<T, 1, 2>  (short K loops, two epilogue sets)
  V_TC_ATM2,         // tbitgemm_atm2_kernel<T, 2, 1> (two output tiles per expanded mask stage)
  V_IMMA,            // ibitgemm_kernel
  V_DMMA,            // bitgemm_kernel
  V_SOLVE_REG8,
  V_SOLVE_REG16,
  V_SOLVE_REG32,
  V_UNUSED9,         // (solve_split64_kernel, removed: superseded by solve_tile_kernel)
  V_SOLVE_TILE,      // solve_tile_kernel<32 | 48 | 64> (register-tiled sweep, 16 < k <= 64)
  V_UNUSED11,
  V_SOLVE_GENERIC,
  V_PRECISION_RETRY,  // passes repeated at a wider arithmetic after the precision guard fired
  V_TC_MIX,           // batched mixture contraction launches
  V_GRAPH_REPLAYS,    // mixture chunk loops replayed from a captured CUDA graph instead of launched one by one
  V_COUNT
};

struct Launcher {
  cudaStream_t stream;
  int64_t *launch_counter;
  int sms;
  int64_t *variants = nullptr;  // V_COUNT counters (nullable)
  void count(int v) const {
    if (variants) ++variants[v];
  }
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute: the guarded block runs once per
// (call site, device), so contexts on several GPUs of one process each opt their kernels in.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool need() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (done.fetch_or(bit) & bit) == 0;
  }
};

// ---- kernels (host launchers) ----------------------------------------------------------------
// ingest.cu
void launch_ingest(const Launcher &L, const double *raw, int64_t nrows, int d, int64_t row0, SampleStore &st);
// compact host format (observed values + row offsets + mask words), see ppca_b200_iterate_packed_host
void launch_unpack(const Launcher &L, const double *vals, const int64_t *rowptr, const uint32_t *maskw, int64_t nrows, int d,
                   int64_t row0, SampleStore &st);
void launch_transpose_mask(const Launcher &L, SampleStore &st);
void launch_export(const Launcher &L, const SampleStore &st, int64_t row0, int64_t nrows, double *out_dev);
void launch_empty_dims(const Launcher &L, const SampleStore &st, uint8_t *out_dev);
void launch_synthetic(const Launcher &L, SampleStore &st, int k_true, double sigma_true, double mask_prob,
                      int n_components, uint64_t seed, int64_t row_begin = 0);
void launch_synth_truth(const Launcher &L, int d, int k_true, int n_components, uint64_t seed, double *Ct_dev, double *mut_dev);
// fast single-model generator (Xi -> DMMA row GEMM -> noise + mask): rows [row_offset, row_offset + rows) of the dataset into
// local rows [0, rows) of st; ws = synth_ws_doubles(rows, d, k) doubles of workspace
size_t synth_ws_doubles(int64_t rows, int d, int k);
void launch_generate_block(const Launcher &L, SampleStore &st, int64_t rows, int64_t row_offset, int k, const double *C_dev,
                           const double *mu_dev, double sigma, double mask_prob, uint64_t seed, double *ws);
// mixture / posterior sampling (sample_general_kernel, ingest.cu); every pointer is a device pointer
struct SamplerArgs {
  int64_t n;
  int d, m, kmax;
  const int *ks;            // m
  const int64_t *coff;      // m : offset of C_j in Cs
  const double *Cs, *mus, *sigmas;
  const double *cdf;        // m cumulative mixture weights (prior sampling of a mixture), or null
  const double *post;       // n x m posterior probabilities (posterior sampling of a mixture), or null
  int post_mode;            // 0: z ~ N(0, I)   1: z ~ N(state_n, covariance_n)
  const double *const *states;  // m device pointers, n x k_j   (post_mode)
  const double *const *covs;    // m device pointers, n x k_j x k_j
  double mask_prob;
  uint64_t seed;
  int *fail;                // set to 1 when a covariance is not positive definite
};
void launch_sample_general(const Launcher &L, SampleStore &st, const SamplerArgs &a);
void launch_model_sample(const Launcher &L, SampleStore &st, int k, const double *C_dev, const double *mu_dev,
                         double sigma, double mask_prob, uint64_t seed);
void launch_copy_rows(const Launcher &L, const SampleStore &src, int64_t src_row0, int64_t nrows, SampleStore &dst,
                      int64_t dst_row0);

// model.cu
void launch_prepare_model(const Launcher &L, const double *C_host_layout_dev, const double *mu_dev, int d, int k,
                          double *Cpad, double *mupad, double *Ksym);

// bitgemm.cu :  Out[M x Nq] (+)= Bits[M x 32*kblocks] * Bmat[32*kblocks x Nq]
struct BitGemmArgs {
  const uint32_t *bits;  // row-major words, row stride ldbits; word offset already applied
  int64_t ldbits;
  const double *Bmat;    // row-major, row stride ldb (even)
  int64_t ldb;
  double *Out;           // row-major, row stride ldo (even)
  int64_t ldo;
  int M;                 // rows of Out (bit rows)
  int Nq;                // columns, multiple of 8
  int kblocks;           // number of 32-wide K blocks
  int kcols;             // valid K columns (<= 32 * kblocks); MMA steps past it are skipped
  int accumulate;        // Out += result (else Out = result)
  double *partials;      // workspace for split-K (may be null when splitk == 1)
  int splitk;
  int defer_reduce;      // split-K only: partials += result (caller zeroed them) and no reduction is launched;
                         // the caller runs launch_bitgemm_reduce once after the last chunk
};
int bitgemm_pick_splitk(int M, int Nq, int kblocks, int sms);
size_t bitgemm_partials_len(int M, int Nq, int splitk);
void launch_bitgemm(const Launcher &L, const BitGemmArgs &a);
void launch_bitgemm_reduce(const Launcher &L, const double *partials, int splitk, int M, int Nq, double *Out,
                           int64_t ldo, int accumulate);

// proj.cu : Y[n][a] = sum_i sel(m_ni, x_ni - mu_i) C[i][a] ; nx[n] = sum_i sel(...)^2
void launch_proj(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m, double *Y,
                 double *nx);

// dense FP64 row GEMM on the projection kernel (covariance diagonals): Y = A * Bt, see proj.cu
void launch_rowgemm(const Launcher &L, const double *A, int lda, int rows_pad, int K, const double *Bt, int n8,
                    const uint32_t *ones, const double *zeros, double *Y, double *nx_scratch);

// solve.cu : per-sample k x k factorisation
struct SolveArgs {
  Shape s;
  double sigma;
  const double *sigma_dev = nullptr;  // when set the kernels read sigma from here (graph replays see the current model)
  int rows;            // valid samples in the chunk
  int rows_pad;        // rows of GW / YZ to (zero) fill, multiple of 32
  double *GW;          // rows_pad x kkp : in G (packed upper), out W = w (z z^T + Sigma)  [mode EM]
  double *YZ;          // rows_pad x kp  : in y, out z
  double *WZ;          // rows_pad x kp  : out w z   (nullable)
  const double *nx;    // rows
  const int *dn;       // rows
  const double *w;     // rows (nullable = 1)
  double *llk;         // rows : out per-sample log-likelihood (nullable)
  double *tn;          // rows : out per-sample tr(Sigma_n G_n) (nullable, mode 2)
  double *dv = nullptr;  // rows : out per-sample |R_n|^2, R_n = m (x~ - C z) (nullable, mode 2).  The residual is never
                         // formed: G z = y - sigma^2 z gives |R_n|^2 = nx - y^T z - sigma^2 |z|^2 (ppca_model.rs:338-346)
  double *cov;         // rows x k x k full covariances (nullable)
  double *part;        // SOLVE_SLOTS x 4 partial sums to accumulate into (nullable; needs llk): w t, w llk, w, #non-empty
  int mode;            // 0 = llk only, 1 = infer (z, cov), 2 = EM (z, W, wz, t)
  unsigned long long *colmax;  // kkp (nullable, mode 2, k <= 64; zeroed by the caller): atomicMax of the bit patterns
                               // of max_n |W[n][q]| — the column scales of the M-step digit planes, fused here
  // Precision guard of the int8-sliced E-step contraction (null gscale = FP64 contraction, no guard).  G_n arrives with
  // an error of at most terms(d_n) s_q 2^-(8T-1) per entry (s_q = power-of-two column scale of Ksym); the sample is
  // accepted when, for every a, that bound on the DIAGONAL entry is below eps (sigma^2 + G_aa), i.e. the perturbation
  // of M_n = sigma^2 I + G_n is small in the diagonally scaled sense under which its inverse / determinant are
  // well conditioned (off-diagonal scales obey s_ab <= 2 sqrt(s_aa s_bb)).  Violations are counted in unsafe[0]; the
  // host then repeats the pass at a wider arithmetic (api.cu, run_guarded).
  double *rscratch = nullptr;      // SOLVE_SCRATCH doubles for the reduction behind `part`
  const double *gscale = nullptr;  // kkp column scales of the Ksym digit planes
  double guard_coef = 0.0;         // 2^-(8T-1) / eps
  unsigned int *unsafe = nullptr;
};
// How many maximal quantisation errors e = s_q 2^-(8T-1) a sum of n terms can carry: all of them for short sums;
// for long ones 4 sqrt(n), a 7-sigma bound (independent round-to-nearest errors are uniform in [-e, e]: the sum has
// standard deviation 0.58 sqrt(n) e).
__host__ __device__ inline double guard_terms(double n) {
  const double r = 4.0 * sqrt(n);
  return n < r ? n : r;
}
enum { SOLVE_SLOTS = 128, SOLVE_SCRATCH = 128 * 6 + 2 };
void launch_solve(const Launcher &L, const SolveArgs &a);
// part[slot][0..3] += sum over the slot's rows of (w t, w llk, w, #non-empty) — what launch_solve does when a.part is set
// dv (nullable) = identity-form residual norms, folded into slot 0 unless the chunk's cancellation check fails, in which
// case *resid_flag = 1 and resid_exact_kernel supplies them; scratch = SOLVE_SCRATCH doubles (zeroed once at allocation)
void launch_solve_reduce(const Launcher &L, int rows, const double *llk, const double *tn, const double *dv,
                         const double *nx, const int *dn, const double *w, double *part, double *scratch, int *resid_flag);
// sums the partial slots (fixed order) into scalars[SC_SQERR, SC_LLK, SC_SUMW, SC_NONEMPTY]
void launch_solve_finish(const Launcher &L, const double *part, double *scalars);

// moments.cu : B += Xc^T (w z), weighted column sums of Xc, totals ; reconstruction writers
// accumulates into per-(slab, dimension block) partial slots (zeroed by the caller before the first chunk);
// launch_cross_resid_finish reduces them once, in fixed order, into the statistics buffer and forms
// tdev_i = Sx_i - sum_a C[i][a] MZ[i][a]  (MZ = Mask^T (w Z), d x kp, from the M-step contraction)
void launch_cross_resid(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m,
                        const double *WZ, const double *w, double *partials, int slabs_alloc);
void launch_cross_resid_finish(const Launcher &L, int d, int k, const double *partials, int slabs_alloc,
                               const double *Cpad, const double *MZ, double *statB, double *statTdev,
                               double *statTotals);
// exact sum_n w_n |m (x~ - C z)|^2 of a chunk into per-(slab, dimension block) slots pD — runs only when *flag != 0
// (every CTA returns at once otherwise); launch_resid_exact_finish adds the slots to scalars[SC_DEV2]
void launch_resid_exact(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m, const double *Z,
                        const double *w, const int *flag, double *pD, int slabs_alloc);
void launch_resid_exact_finish(const Launcher &L, int d, const double *pD, int slabs_alloc, double *scalars);
size_t resid_exact_partials_len(int d, int slabs_alloc);
int cross_resid_slabs(int d, int k, int rows, int sms);
size_t cross_resid_partials_len(int d, int k, int slabs_alloc);
// out = extrapolate ? (m ? x : C z + mu) : C z + mu ; scale != null multiplies by scale[n*scale_ld] and accumulates
void launch_reconstruct(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m,
                        const double *Z, int extrapolate, const double *scale, int64_t scale_ld, int accumulate,
                        double *outX, int64_t ldo);

// finish.cu : precision guard of the int8-sliced M-step contraction.  A_i[a][a] (a sum of positive terms) is compared with
// the bound terms(n) smax_aa 2^-(8T-1) on its quantisation error; dimensions nobody observed (totals == 0) are skipped.
// One block; writes scalars[SC_UNSAFE_E] = unsafe[0] (E-step violations counted by the solve kernels) and
// scalars[SC_UNSAFE_M] = violating (dimension, a) pairs.  smax == nullptr: only publishes the E-step counter.
void launch_mstep_guard(const Launcher &L, int d, int k, const double *statA, const double *totals, const double *smax,
                        double coef_terms, const unsigned int *unsafe, double *scalars);
// smax[q] = max(smax[q], scale[q])
void launch_scale_max(const Launcher &L, const double *scale, int n, double *smax);

// finish.cu : mean prior (prior.rs:97-110).  P_dev (n x n, prior precision on entry) is overwritten by the Cholesky factor of
// P + diag(totals / noise_sq); rhs_dev receives the solution.  *fail_dev != 0: not positive definite (caller falls back).
void launch_mean_prior_solve(const Launcher &L, double *P_dev, int n, const double *m0_dev, const double *totals_dev,
                             const double *mu_hat_dev, double noise_sq, double *rhs_dev, int *fail_dev);

// finish.cu : per-dimension solves (A_i + tau I) c = B_i
void launch_row_solve(const Launcher &L, int d, int k, const double *statA, const double *statB, double tau,
                      const double *Cold_pad, double *Cnew /* d x k dense */, int *flags);

// comm.cu : NCCL bound at run time (see the file header)
void comm_unique_id(uint8_t *out128);
void *comm_create(const uint8_t *id128, int rank, int world);
void comm_destroy(void *comm);
void comm_allreduce(void *comm, double *buf_dev, int64_t count, int op /* 0 = sum, 1 = max */, cudaStream_t stream);
int comm_version();

// mix.cu
void launch_log_softmax_rows(const Launcher &L, double *LP, int64_t n, int m, const double *logw_dev,
                             const double *w, double *mix_llk /* n, nullable */, double *comp_max /* m */,
                             double *llk_sum /* 1, nullable */);
void launch_responsibilities(const Launcher &L, const double *LP, int64_t n, int m, int j, const double *w,
                             double comp_max_j, double *r_out);
// ---- single-pass mixture EM (api.cu, mix_em_pass) ----
// accumulate != 0: llk_sum[0] += sum_n w_n mix_llk[n] (one block, fixed order), else overwrite
void launch_weighted_sum(const Launcher &L, const double *v, const double *w, int64_t n, double *out, int accumulate);
// run_max[j] <- max(run_max[j], chunk_max[j]) ; factor[j] = exp(old - new) (1 when unchanged, 0 on the first chunk)
void launch_mix_update_max(const Launcher &L, int m, double *run_max, const double *chunk_max, double *factor);
// buf[0, len) *= factor[0]; returns at once when the factor is exactly 1.  stride4 != 0: buf is [slots][4] and only the
// first three entries of each slot are scaled (the fourth is a count).
void launch_scale_by(const Launcher &L, double *buf, int64_t len, const double *factor, int stride4);
// r[n] = w_n > 0 ? exp(ln w_n + LP[n][j] - run_max[j]) : 0 for n < rows (0 up to rows_pad) ;
// W[n][q] = r_n V[n][q] in place (rows_pad x kkp) ; WZ[n][a] = r_n Z[n][a] ; colmax[q] = max_n |W[n][q]| (bit patterns)
void launch_mix_weight(const Launcher &L, const double *LP, int m, int j, const double *w, const double *run_max, int rows,
                       int rows_pad, int kkp, int kp, double *V, const double *Z, double *WZ, double *r,
                       unsigned long long *colmax);
// stats[..] *= exp(local_max[j] - global_max[j]) for the weight-linear entries (everything except the two counters and
// the guard slots)
void launch_mix_rescale_stats(const Launcher &L, double *stats, int64_t len_linear, double *scalars, const double *local_max,
                              const double *global_max, int j);

}  // namespace ppca

// ---- ibitgemm.cu : the same masked contraction evaluated exactly on the int8 tensor path ---------------
// Out[M x Nq] (+)= Bits[M x K] * Bmat[K x Nq] with Bmat pre-split into T signed 7-bit slices per column
// (column scale 2^e): every partial sum is an exact int32, slices are recombined in FP64 in the epilogue.
namespace ppca {
struct SlicedB {
  int8_t *q = nullptr;      // [kblocks][T][Nq][32]  (32 K-bytes permuted for the IMMA B fragment)
  double *scale = nullptr;  // [Nq] column scale s = 2^e >= max |column|
  int T = 0;
};
size_t sliced_bytes(int kblocks, int Nq, int T);
// colmax over rows [0,K) of Bmat, then digits; rows in [K, 32*kblocks) are treated as zero
void launch_slice(const Launcher &L, const double *Bmat, int64_t ldb, int K, int Nq, int kblocks, int T, int8_t *q,
                  double *scale, unsigned long long *colmax_scratch);
struct IBitGemmArgs {
  const uint32_t *bits;
  int64_t ldbits;
  const int8_t *Bq;
  const double *scale;
  int T;
  double *Out;
  int64_t ldo;
  int M, Nq, kblocks;
  int accumulate;
  double *partials;
  int splitk;
  int defer_reduce;
};
int ibitgemm_pick_splitk(int M, int Nq, int kblocks, int sms);
void launch_ibitgemm(const Launcher &L, const IBitGemmArgs &a);

// ---- tbitgemm.cu : the int8-sliced contraction on tcgen05 (TMEM accumulators) ----------------------------
size_t sliced_tc_bytes(int kblocks32, int Nq, int T);
// have_colmax: colmax_scratch already holds the column maxima (produced by the solve kernel); skip that pass
void launch_slice_tc(const Launcher &L, const double *Bmat, int64_t ldb, int K, int Nq, int kblocks32, int T, int8_t *q,
                     double *scale, unsigned long long *colmax_scratch, bool have_colmax = false);
int tbitgemm_pick_splitk(int M, int Nq, int ksteps, int sms);
void launch_tbitgemm(const Launcher &L, const uint32_t *bits, int64_t ldbits, int nwords, const int8_t *Bq,
                     const double *scale, int T, double *Out, int64_t ldo, int M, int Nq, int ksteps, int accumulate,
                     double *partials, int splitk, int defer_reduce);
}  // namespace ppca
