// ingest.cu — Dataset / MaskedSample / Mask on the device.
//
// Reference data model (dataset.rs:11-14,93-100; utils.rs:27-28): an array of heap vectors, each with a
// BitVec mask where bit = is_finite(x) (dataset.rs:19-22).  Here: one row-major f64 matrix X (masked slots
// sanitised to 0.0), a bit-packed mask with the same LSB-first u32 block layout as bit-vec, its 32x32-block
// transpose (left operand of the M-step contraction), and the per-sample observed counts.
#include "common.cuh"
#include "mma.cuh"

namespace ppca {

// one warp per sample row: mask = isfinite, sanitise, pack with ballot, popcount
__global__ void ingest_kernel(const double *__restrict__ raw, int64_t nrows, int d, int64_t row0, double *X, int ldx,
                              uint32_t *mask, int dw, int *dn) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const double *src = raw + row * d;
  double *dst = X + (row0 + row) * ldx;
  int count = 0;
  for (int j = 0; j < dw; ++j) {
    const int i = 32 * j + lane;
    double x = (i < d) ? src[i] : __longlong_as_double(0x7ff8000000000000LL);
    const bool fin = isfinite(x);
    const uint32_t word = __ballot_sync(0xffffffffu, fin);
    if (i < ldx) dst[i] = fin ? x : 0.0;
    if (lane == 0) mask[(row0 + row) * dw + j] = word;
    count += __popc(word);
  }
  if (lane == 0) dn[row0 + row] = count;
}

void launch_ingest(const Launcher &L, const double *raw, int64_t nrows, int d, int64_t row0, SampleStore &st) {
  if (nrows <= 0) return;
  const int threads = 256;
  const int64_t blocks = (nrows * 32 + threads - 1) / threads;
  ingest_kernel<<<(unsigned)blocks, threads, 0, L.stream>>>(raw, nrows, d, row0, st.X.p, st.ldx, st.mask.p, st.dw,
                                                            st.dn.p);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// The compact host format of the out-of-core path: only the OBSERVED values cross the bus.  One warp per sample row
// scatters vals[rowptr[row] - rowptr[0] + rank of the bit] into X, zeros elsewhere; the mask words are copied as they are.
__global__ void unpack_kernel(const double *__restrict__ vals, const int64_t *__restrict__ rowptr,
                              const uint32_t *__restrict__ maskw, int64_t nrows, int d, int64_t row0, double *X, int ldx,
                              uint32_t *mask, int dw, int *dn) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const double *src = vals + (rowptr[row] - rowptr[0]);
  double *dst = X + (row0 + row) * ldx;
  int count = 0;
  for (int j = 0; j < dw; ++j) {
    const int i = 32 * j + lane;
    uint32_t word = maskw[row * dw + j];
    if (32 * j + 32 > d) word &= (d - 32 * j >= 32) ? 0xffffffffu : ((1u << (d - 32 * j)) - 1u);  // bits past d never count
    const bool obs = (word >> lane) & 1u;
    const int pos = count + __popc(word & ((1u << lane) - 1u));
    if (i < ldx) dst[i] = obs ? src[pos] : 0.0;
    if (lane == 0) mask[(row0 + row) * dw + j] = word;
    count += __popc(word);
  }
  if (lane == 0) dn[row0 + row] = count;
}

void launch_unpack(const Launcher &L, const double *vals, const int64_t *rowptr, const uint32_t *maskw, int64_t nrows, int d,
                   int64_t row0, SampleStore &st) {
  if (nrows <= 0) return;
  const int threads = 256;
  const int64_t blocks = (nrows * 32 + threads - 1) / threads;
  unpack_kernel<<<(unsigned)blocks, threads, 0, L.stream>>>(vals, rowptr, maskw, nrows, d, row0, st.X.p, st.ldx, st.mask.p,
                                                            st.dw, st.dn.p);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// maskT[32 j + b][bs] bit l  =  mask[32 bs + l][j] bit b.   One warp per 32x32 bit block; 32 warps per CTA
// cover 32 consecutive sample blocks so the transposed words leave as 128-byte rows.
__global__ void __launch_bounds__(1024) transpose_mask_kernel(const uint32_t *__restrict__ mask, int dw, int64_t n_pad,
                                                              uint32_t *maskT, int64_t nwT) {
  __shared__ uint32_t tile[32][33];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int j = blockIdx.y;
  const int64_t bs = (int64_t)blockIdx.x * 32 + wi;
  uint32_t v = 0;
  if (bs < nwT) v = mask[(bs * 32 + lane) * dw + j];
  uint32_t mine = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const uint32_t t = __ballot_sync(0xffffffffu, (v >> b) & 1u);
    if (lane == b) mine = t;
  }
  tile[lane][wi] = mine;
  __syncthreads();
  const int64_t col = (int64_t)blockIdx.x * 32 + lane;
  if (col < nwT) maskT[((int64_t)32 * j + wi) * nwT + col] = tile[wi][lane];
}

void launch_transpose_mask(const Launcher &L, SampleStore &st) {
  if (st.n == 0) return;
  dim3 grid((unsigned)((st.nwT + 31) / 32), (unsigned)st.dw);
  transpose_mask_kernel<<<grid, 1024, 0, L.stream>>>(st.mask.p, st.dw, st.n_pad, st.maskT.p, st.nwT);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// Dataset.numpy(): NaN at masked slots (dataset.rs:64-72 masked_vector)
__global__ void export_kernel(const double *__restrict__ X, int ldx, const uint32_t *__restrict__ mask, int dw, int d,
                              int64_t row0, int64_t nrows, double *out) {
  const int64_t total = nrows * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / d;
    const int i = (int)(idx % d);
    const uint32_t w = mask[(row0 + row) * dw + (i >> 5)];
    out[idx] = ((w >> (i & 31)) & 1u) ? X[(row0 + row) * ldx + i] : __longlong_as_double(0x7ff8000000000000LL);
  }
}

void launch_export(const Launcher &L, const SampleStore &st, int64_t row0, int64_t nrows, double *out_dev) {
  if (nrows <= 0) return;
  const int64_t total = nrows * st.d;
  const int64_t want = (total + 255) / 256;
  const int blocks = (int)(want < (int64_t)L.sms * 16 ? want : (int64_t)L.sms * 16);
  export_kernel<<<blocks, 256, 0, L.stream>>>(st.X.p, st.ldx, st.mask.p, st.dw, st.d, row0, nrows, out_dev);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// Dataset::empty_dimensions (dataset.rs:194-222): OR of all masks; one warp per dimension over maskT
__global__ void empty_dims_kernel(const uint32_t *__restrict__ maskT, int64_t nwT, int d, uint8_t *out) {
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (i >= d) return;
  uint32_t acc = 0;
  for (int64_t j = lane; j < nwT; j += 32) acc |= maskT[(int64_t)i * nwT + j];
  acc = __reduce_or_sync(0xffffffffu, acc);
  if (lane == 0) out[i] = acc == 0u ? 1 : 0;
}

void launch_empty_dims(const Launcher &L, const SampleStore &st, uint8_t *out_dev) {
  if (st.d == 0) return;
  const int threads = 256;
  const int blocks = (st.d * 32 + threads - 1) / threads;
  empty_dims_kernel<<<blocks, threads, 0, L.stream>>>(st.maskT.p, st.nwT, st.d, out_dev);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_copy_rows(const Launcher &L, const SampleStore &src, int64_t src_row0, int64_t nrows, SampleStore &dst,
                      int64_t dst_row0) {
  if (nrows <= 0) return;
  CUDA_CHECK(cudaMemcpyAsync(dst.X.p + dst_row0 * dst.ldx, src.X.p + src_row0 * src.ldx,
                             sizeof(double) * nrows * src.ldx, cudaMemcpyDeviceToDevice, L.stream));
  CUDA_CHECK(cudaMemcpyAsync(dst.mask.p + dst_row0 * dst.dw, src.mask.p + src_row0 * src.dw,
                             sizeof(uint32_t) * nrows * src.dw, cudaMemcpyDeviceToDevice, L.stream));
  CUDA_CHECK(cudaMemcpyAsync(dst.dn.p + dst_row0, src.dn.p + src_row0, sizeof(int) * nrows, cudaMemcpyDeviceToDevice,
                             L.stream));
}

// ---------------------------------------------------------------------------------------------
// synthetic data (ppca_model.rs:164-191 sample_one semantics; mix.rs:124-134 for n_components > 1)
// counter-based RNG: splitmix64 of (seed, stream, counter) -> uniform -> Box-Muller
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ inline uint64_t rng_u64(uint64_t seed, uint64_t stream, uint64_t ctr) {
  return splitmix64(splitmix64(seed ^ (stream * 0xD1342543DE82EF95ULL)) + ctr);
}
__device__ inline double rng_uniform(uint64_t seed, uint64_t stream, uint64_t ctr) {  // (0,1)
  return ((double)(rng_u64(seed, stream, ctr) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ inline double rng_normal(uint64_t seed, uint64_t stream, uint64_t ctr) {
  const double u1 = rng_uniform(seed, stream, 2 * ctr), u2 = rng_uniform(seed, stream, 2 * ctr + 1);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// truth tables: Ct[j][i][a] ~ Bernoulli(0.1) (examples/big_toy_model.py:6), mut[j][i] = 0 for a single
// component, ~ N(0, 1) for mixtures so that components are distinguishable
__global__ void synth_truth_kernel(int d, int k_true, int n_components, uint64_t seed, double *Ct, double *mut) {
  const int64_t total = (int64_t)n_components * d * k_true;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x)
    Ct[idx] = rng_uniform(seed, 1, (uint64_t)idx) < 0.1 ? 1.0 : 0.0;
  const int64_t totm = (int64_t)n_components * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < totm;
       idx += (int64_t)gridDim.x * blockDim.x)
    mut[idx] = n_components > 1 ? rng_normal(seed, 2, (uint64_t)idx) : 0.0;
}

// one warp per sample
// row_offset: the block's first row in the dataset the counters are keyed on (a chunk regenerated on its own is the
// same rows a resident dataset of the whole job would hold); rows are written from local row 0
__global__ void __launch_bounds__(256) synth_kernel(int64_t n, int d, int k_true, int n_components, double sigma_true,
                                                    double mask_prob, uint64_t seed, const double *__restrict__ Ct,
                                                    const double *__restrict__ mut, double *X, int ldx, uint32_t *mask,
                                                    int dw, int *dn, int64_t row_offset) {
  extern __shared__ double xi_all[];  // warps x k_true
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t lrow = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (lrow >= n) return;
  const int64_t row = lrow + row_offset;
  X += (lrow - row) * (int64_t)ldx;  // outputs are indexed by the local row
  mask += (lrow - row) * (int64_t)dw;
  dn += (lrow - row);
  double *xi = xi_all + wi * k_true;
  for (int a = lane; a < k_true; a += 32) xi[a] = rng_normal(seed, 3, (uint64_t)row * k_true + a);
  __syncwarp();
  const int comp = n_components > 1 ? (int)(rng_u64(seed, 4, (uint64_t)row) % (uint64_t)n_components) : 0;
  const double *Cj = Ct + (int64_t)comp * d * k_true;
  const double *mj = mut + (int64_t)comp * d;
  int count = 0;
  for (int j = 0; j < dw; ++j) {
    const int i = 32 * j + lane;
    double x = 0.0;
    bool obs = false;
    if (i < d) {
      const double *crow = Cj + (int64_t)i * k_true;
      double acc = mj[i];
      for (int a = 0; a < k_true; ++a) acc += crow[a] * xi[a];
      x = acc + sigma_true * rng_normal(seed, 5, (uint64_t)row * d + i);
      obs = !(rng_uniform(seed, 6, (uint64_t)row * d + i) < mask_prob);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, obs);
    if (i < ldx) X[row * ldx + i] = obs ? x : 0.0;
    if (lane == 0) mask[row * dw + j] = word;
    count += __popc(word);
  }
  if (lane == 0) dn[row] = count;
}

void launch_generate_rows(const Launcher &L, int d, double *X, int ldx, uint32_t *mask, int dw, int *dn, int64_t rows,
                          int64_t row_offset, int k, const double *C_dev, const double *mu_dev, double sigma, double mask_prob,
                          uint64_t seed, double *ws);
size_t synth_ws_doubles(int64_t rows, int d, int k);

// rows of a resident dataset through the fast generator, block by block (the GEMM workspace is rows x d doubles)
static void generate_resident(const Launcher &L, SampleStore &st, int k, const double *C_dev, const double *mu_dev, double sigma,
                              double mask_prob, uint64_t seed, int64_t row_begin = 0) {
  const int64_t blk = std::max<int64_t>(128, std::min<int64_t>(((int64_t)1 << 30) / ((int64_t)st.d * 8) / 128 * 128, 1 << 18));
  DevBuf<double> ws;
  if (st.n <= 0) return;
  ws.alloc(synth_ws_doubles(std::min<int64_t>(blk, st.n), st.d, k));
  for (int64_t r0 = 0; r0 < st.n; r0 += blk) {
    const int64_t rows = std::min<int64_t>(blk, st.n - r0);
    launch_generate_rows(L, st.d, st.X.p + r0 * st.ldx, st.ldx, st.mask.p + r0 * st.dw, st.dw, st.dn.p + r0, rows, row_begin + r0,
                         k, C_dev, mu_dev, sigma, mask_prob, seed, ws.p);
  }
  CUDA_CHECK(cudaStreamSynchronize(L.stream));  // ws is freed on return
}

void launch_synthetic(const Launcher &L, SampleStore &st, int k_true, double sigma_true, double mask_prob,
                      int n_components, uint64_t seed, int64_t row_begin) {
  DevBuf<double> Ct, mut;
  Ct.alloc((size_t)n_components * st.d * k_true);
  mut.alloc((size_t)n_components * st.d);
  synth_truth_kernel<<<L.sms * 4, 256, 0, L.stream>>>(st.d, k_true, n_components, seed, Ct.p, mut.p);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  if (n_components == 1) {
    generate_resident(L, st, k_true, Ct.p, mut.p, sigma_true, mask_prob, seed, row_begin);
    return;
  }
  const int threads = 256;
  const int64_t blocks = (st.n * 32 + threads - 1) / threads;
  const size_t smem = sizeof(double) * (threads / 32) * k_true;
  synth_kernel<<<(unsigned)blocks, threads, smem, L.stream>>>(st.n, st.d, k_true, n_components, sigma_true, mask_prob,
                                                              seed, Ct.p, mut.p, st.X.p, st.ldx, st.mask.p, st.dw,
                                                              st.dn.p, row_begin);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  CUDA_CHECK(cudaStreamSynchronize(L.stream));  // Ct / mut are freed on return
}


// ---------------------------------------------------------------------------------------------
// Fast single-model generator: x = C xi + mu + sigma eps as a dense FP64 row GEMM (DMMA, launch_rowgemm) between two
// streaming kernels, for any block of rows of the dataset — every counter is keyed by the GLOBAL row, so a block
// regenerated on its own is bit-identical to the same rows of a resident dataset.  Used by Dataset.synthetic (one
// component), PPCAModel.sample and the out-of-core path that regenerates every chunk (ppca_b200_iterate_generated).
// The one-warp-per-sample synth_kernel above (k loads + k FMAs per element) stays for mixtures of truths.
// ---------------------------------------------------------------------------------------------
__global__ void synth_xi_kernel(int64_t rows, int64_t rows_pad, int k, int kpad, uint64_t seed, int64_t row_offset,
                                double *__restrict__ Xi) {
  const int64_t total = rows_pad * kpad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / kpad;
    const int a = (int)(idx % kpad);
    Xi[idx] = (r < rows && a < k) ? rng_normal(seed, 3, (uint64_t)(row_offset + r) * k + a) : 0.0;
  }
}

// Bt[a][i] = C[i][a], zero padded to kpad x n8 (the right operand layout of launch_rowgemm)
__global__ void synth_bt_kernel(const double *__restrict__ C, int d, int k, int kpad, int n8, double *__restrict__ Bt) {
  const int64_t total = (int64_t)kpad * n8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(idx / n8), i = (int)(idx % n8);
    Bt[idx] = (a < k && i < d) ? C[(int64_t)i * k + a] : 0.0;
  }
}

// one warp per sample: x = Y + mu + sigma eps, mask bit ~ Bernoulli(1 - mask_prob).  One Box-Muller pair serves the two
// elements 32 j + lane and 32 (j + 1) + lane of a pair of mask words.
__global__ void __launch_bounds__(256) synth_finish_kernel(int64_t rows, int d, int n8, const double *__restrict__ Y,
                                                           const double *__restrict__ mu, double sigma, double mask_prob,
                                                           uint64_t seed, int64_t row_offset, double *X, int ldx,
                                                           uint32_t *mask, int dw, int *dn) {
  const int lane = threadIdx.x & 31;
  const int64_t lrow = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (lrow >= rows) return;
  const uint64_t grow = (uint64_t)(row_offset + lrow);
  const int pairs = (dw + 1) / 2;
  const double *y = Y + lrow * (int64_t)n8;
  double *xo = X + lrow * (int64_t)ldx;
  int count = 0;
  for (int jp = 0; jp < pairs; ++jp) {
    const uint64_t c = (grow * pairs + jp) * 32 + lane;
    const double u1 = rng_uniform(seed, 5, 2 * c), u2 = rng_uniform(seed, 5, 2 * c + 1);
    const double r = sigma * sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    const double eps[2] = {r * cs, r * sn};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * jp + h;
      if (j >= dw) break;
      const int i = 32 * j + lane;
      double x = 0.0;
      bool obs = false;
      if (i < d) {
        x = (y[i] + mu[i]) + eps[h];
        obs = !(rng_uniform(seed, 6, grow * d + i) < mask_prob);
      }
      const uint32_t word = __ballot_sync(0xffffffffu, obs);
      if (i < ldx) xo[i] = obs ? x : 0.0;
      if (lane == 0) mask[lrow * dw + j] = word;
      count += __popc(word);
    }
  }
  if (lane == 0) dn[lrow] = count;
}

size_t synth_ws_doubles(int64_t rows, int d, int k) {  // workspace of one block: Xi | Y | Bt | zeros | ones (as doubles)
  const int64_t rows_pad = (rows + 127) / 128 * 128;
  const int kpad = (k + 31) / 32 * 32, n8 = (d + 7) / 8 * 8;
  return (size_t)rows_pad * kpad + (size_t)rows_pad * n8 + (size_t)kpad * n8 + (size_t)kpad + (size_t)kpad / 32 + 8 +
         (size_t)rows_pad;
}

// rows [row_offset, row_offset + rows) of the dataset into local rows [0, rows) of st; C_dev: d x k dense, mu_dev: d
void launch_generate_rows(const Launcher &L, int d, double *X, int ldx, uint32_t *mask, int dw, int *dn, int64_t rows,
                          int64_t row_offset, int k, const double *C_dev, const double *mu_dev, double sigma, double mask_prob,
                          uint64_t seed, double *ws) {
  if (rows <= 0) return;
  const int64_t rows_pad = (rows + 127) / 128 * 128;
  const int kpad = (k + 31) / 32 * 32, n8 = (d + 7) / 8 * 8;
  double *Xi = ws, *Y = Xi + rows_pad * kpad, *Bt = Y + rows_pad * n8, *zeros = Bt + (size_t)kpad * n8;
  uint32_t *ones = reinterpret_cast<uint32_t *>(zeros + kpad);
  double *nxs = zeros + kpad + kpad / 32 + 8;
  CUDA_CHECK(cudaMemsetAsync(zeros, 0, sizeof(double) * kpad, L.stream));
  CUDA_CHECK(cudaMemsetAsync(ones, 0xff, sizeof(uint32_t) * (kpad / 32), L.stream));
  const int gb = (int)std::min<int64_t>((int64_t)L.sms * 16, (rows_pad * kpad + 255) / 256);
  synth_xi_kernel<<<gb, 256, 0, L.stream>>>(rows, rows_pad, k, kpad, seed, row_offset, Xi);
  CUDA_CHECK(cudaGetLastError());
  const int bb = (int)std::min<int64_t>((int64_t)L.sms * 16, ((int64_t)kpad * n8 + 255) / 256);
  synth_bt_kernel<<<bb, 256, 0, L.stream>>>(C_dev, d, k, kpad, n8, Bt);
  CUDA_CHECK(cudaGetLastError());
  *L.launch_counter += 2;
  launch_rowgemm(L, Xi, kpad, (int)rows_pad, kpad, Bt, n8, ones, zeros, Y, nxs);
  const int64_t fb = (rows * 32 + 255) / 256;
  synth_finish_kernel<<<(unsigned)fb, 256, 0, L.stream>>>(rows, d, n8, Y, mu_dev, sigma, mask_prob, seed, row_offset, X, ldx,
                                                          mask, dw, dn);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_generate_block(const Launcher &L, SampleStore &st, int64_t rows, int64_t row_offset, int k, const double *C_dev,
                           const double *mu_dev, double sigma, double mask_prob, uint64_t seed, double *ws) {
  launch_generate_rows(L, st.d, st.X.p, st.ldx, st.mask.p, st.dw, st.dn.p, rows, row_offset, k, C_dev, mu_dev, sigma, mask_prob,
                       seed, ws);
}

// Truth tables of the synthetic generator (Ct d x k_true, mut d per component), kept by the caller
void launch_synth_truth(const Launcher &L, int d, int k_true, int n_components, uint64_t seed, double *Ct_dev, double *mut_dev) {
  synth_truth_kernel<<<L.sms * 4, 256, 0, L.stream>>>(d, k_true, n_components, seed, Ct_dev, mut_dev);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// PPCAModel::sample (ppca_model.rs:164-191) with a caller-supplied model: x = C xi + mu + sigma eps, masked with
// probability mask_prob; C_dev is the dense d x k transform, mu_dev the mean (both on the device).
void launch_model_sample(const Launcher &L, SampleStore &st, int k, const double *C_dev, const double *mu_dev,
                         double sigma, double mask_prob, uint64_t seed) {
  generate_resident(L, st, k, C_dev, mu_dev, sigma, mask_prob, seed);
}


// ---------------------------------------------------------------------------------------------
// Mixture / posterior sampling (mix.rs:124-134 PPCAMix::sample; ppca_model.rs:581-626 PosteriorSampler;
// mix.rs:505-532 PosteriorSamplerMix).  One warp per output sample:
//   1. component j: inverse-CDF draw from the mixture weights (cdf, shared by all rows) or from this row's
//      posterior probabilities (post, n x m; normalised here, WeightedIndex semantics); m == 1 -> 0
//   2. state: xi ~ N(0, I_k) (prior sampling) or z = state_n + L_n xi with L_n L_n^T = covariance_n, the Cholesky
//      factor formed here in shared memory (posterior sampling)
//   3. x = C_j z + mu_j + sigma_j eps, each entry masked with probability mask_prob
// A non-positive Cholesky pivot raises *fail (the reference `expect`s "Cholesky decomposition failed").
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_general_kernel(SamplerArgs a, double *X, int ldx, uint32_t *mask, int dw,
                                                             int *dn) {
  extern __shared__ double smem_sampler[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + wi;
  if (row >= a.n) return;
  const int kmax = a.kmax;
  double *z = smem_sampler + (size_t)wi * (kmax + (a.post_mode ? kmax * kmax + kmax : 0));
  double *xi = z + kmax;         // posterior mode only
  double *Lm = xi + kmax;        // kmax x kmax, posterior mode only
  // 1. component
  int comp = 0;
  if (a.m > 1) {
    const double u = rng_uniform(a.seed, 4, (uint64_t)row);
    if (a.post) {
      const double *pr = a.post + row * a.m;
      double tot = 0.0;
      for (int j = 0; j < a.m; ++j) tot += pr[j];
      double acc = 0.0;
      comp = a.m - 1;
      for (int j = 0; j < a.m; ++j) {
        acc += pr[j];
        if (u * tot < acc) {
          comp = j;
          break;
        }
      }
    } else {
      comp = a.m - 1;
      for (int j = 0; j < a.m; ++j)
        if (u < a.cdf[j]) {
          comp = j;
          break;
        }
    }
  }
  const int k = a.ks[comp];
  const double *Cj = a.Cs + a.coff[comp];
  const double *mj = a.mus + (int64_t)comp * a.d;
  const double sigma = a.sigmas[comp];
  // 2. state
  if (!a.post_mode) {
    for (int q = lane; q < k; q += 32) z[q] = rng_normal(a.seed, 3, (uint64_t)row * kmax + q);
  } else {
    const double *cov = a.covs[comp] + row * (int64_t)k * k;
    const double *stt = a.states[comp] + row * (int64_t)k;
    for (int q = lane; q < k * k; q += 32) Lm[q] = cov[q];
    for (int q = lane; q < k; q += 32) xi[q] = rng_normal(a.seed, 3, (uint64_t)row * kmax + q);
    __syncwarp();
    for (int p = 0; p < k; ++p) {  // right-looking Cholesky, lower triangle in place
      const double dpp = Lm[p * k + p];
      if (!(dpp > 0.0)) {
        if (lane == 0) atomicExch(a.fail, 1);
        break;
      }
      const double inv = 1.0 / sqrt(dpp);
      __syncwarp();
      for (int r = p + lane; r < k; r += 32) Lm[r * k + p] *= inv;  // the diagonal becomes sqrt(d)
      __syncwarp();
      for (int c = p + 1; c < k; ++c) {
        const double lcp = Lm[c * k + p];
        for (int r = c + lane; r < k; r += 32) Lm[r * k + c] = fma(-Lm[r * k + p], lcp, Lm[r * k + c]);
      }
      __syncwarp();
    }
    for (int q = lane; q < k; q += 32) {
      double acc = stt[q];
      for (int c = 0; c <= q; ++c) acc = fma(Lm[q * k + c], xi[c], acc);
      z[q] = acc;
    }
  }
  __syncwarp();
  // 3. outputs
  int count = 0;
  for (int j = 0; j < dw; ++j) {
    const int i = 32 * j + lane;
    double x = 0.0;
    bool obs = false;
    if (i < a.d) {
      const double *crow = Cj + (int64_t)i * k;
      double acc = mj[i];
      for (int q = 0; q < k; ++q) acc = fma(crow[q], z[q], acc);
      x = acc + sigma * rng_normal(a.seed, 5, (uint64_t)row * a.d + i);
      obs = !(rng_uniform(a.seed, 6, (uint64_t)row * a.d + i) < a.mask_prob);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, obs);
    if (i < ldx) X[row * ldx + i] = obs ? x : 0.0;
    if (lane == 0) mask[row * dw + j] = word;
    count += __popc(word);
  }
  if (lane == 0) dn[row] = count;
}

void launch_sample_general(const Launcher &L, SampleStore &st, const SamplerArgs &a) {
  if (st.n <= 0) return;
  const size_t per_warp = sizeof(double) * (a.kmax + (a.post_mode ? (size_t)a.kmax * a.kmax + a.kmax : 0));
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  REQUIRE(per_warp * warps <= 200 * 1024, "state_size %d too large for the sampler kernel", a.kmax);
  const size_t smem = per_warp * warps;
  static PerDeviceOnce configured;
  if (configured.need())
    CUDA_CHECK(cudaFuncSetAttribute(sample_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t blocks = (st.n + warps - 1) / warps;
  sample_general_kernel<<<(unsigned)blocks, warps * 32, smem, L.stream>>>(a, st.X.p, st.ldx, st.mask.p, st.dw, st.dn.p);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
