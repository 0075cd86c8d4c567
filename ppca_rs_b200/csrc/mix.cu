// mix.cu — mixture responsibilities in the log domain.
//
// Reference: robust_log_softmax / robust_log_softnorm (mix.rs:14-25); PPCAMix::llks_one / llk_one (:137-149);
// infer_cluster (:179-189); the responsibility weights of iterate_with_prior (:297-326):
//   lp_n = ln w_n + log_posterior[n][j] ; max_j = max_n lp_n ; r_n = exp(lp_n - max_j).
#include <cfloat>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

__device__ inline double atomic_max_double(double *addr, double v) {
  // monotone mapping of doubles onto signed 64-bit integers
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    if (!(v > cur)) break;
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
  return __longlong_as_double((long long)old);
}

// In place on LP (n x m): LP[n][j] <- log_softmax_j(LP[n][j] + logw[j]).
// mix_llk[n] = log sum_j exp(LP[n][j] + logw[j]) ; comp_max[j] = max_n (ln w_n + log-posterior) over w_n > 0.
__global__ void log_softmax_rows_kernel(double *LP, int64_t n, int m, const double *__restrict__ logw,
                                        const double *__restrict__ w, double *mix_llk, double *comp_max) {
  extern __shared__ double sh_max[];  // m
  for (int j = threadIdx.x; j < m; j += blockDim.x) sh_max[j] = -INFINITY;
  __syncthreads();
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n;
       row += (int64_t)gridDim.x * blockDim.x) {
    double *p = LP + row * m;
    double mx = -INFINITY;
    for (int j = 0; j < m; ++j) {
      const double v = p[j] + logw[j];
      p[j] = v;
      mx = fmax(mx, v);  // DVector::max
    }
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += exp(p[j] - mx);
    const double ln = log(s);
    if (mix_llk) mix_llk[row] = mx + ln;
    const double wn = w ? w[row] : 1.0;
    const double lw = log(wn);
    for (int j = 0; j < m; ++j) {
      const double lp = p[j] - mx - ln;
      p[j] = lp;
      if (comp_max && wn > 0.0) {
        const double v = lw + lp;
        if (v == v) atomic_max_double(&sh_max[j], v);  // NaN ignored (ordered_float::NotNan filter)
      }
    }
  }
  __syncthreads();
  if (comp_max)
    for (int j = threadIdx.x; j < m; j += blockDim.x)
      if (sh_max[j] > -INFINITY) atomic_max_double(&comp_max[j], sh_max[j]);
}

__global__ void fill_kernel(double *p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// fixed-order weighted sum: out[0] = sum_n w_n v_n  (single block)
__global__ void __launch_bounds__(1024) weighted_sum_kernel(const double *__restrict__ v, const double *__restrict__ w,
                                                            int64_t n, double *out, int accumulate = 0) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s = fma(w ? w[i] : 1.0, v[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum(sh[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
  }
}

void launch_log_softmax_rows(const Launcher &L, double *LP, int64_t n, int m, const double *logw_dev, const double *w,
                             double *mix_llk, double *comp_max, double *llk_sum) {
  if (comp_max) {
    fill_kernel<<<1, 64, 0, L.stream>>>(comp_max, m, -INFINITY);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
  }
  if (n > 0) {
    const int64_t want = (n + 255) / 256;
    const int blocks = (int)(want < (int64_t)L.sms * 8 ? want : (int64_t)L.sms * 8);
    log_softmax_rows_kernel<<<blocks, 256, sizeof(double) * m, L.stream>>>(LP, n, m, logw_dev, w, mix_llk, comp_max);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
  }
  if (llk_sum) {
    weighted_sum_kernel<<<1, 1024, 0, L.stream>>>(mix_llk, w, n, llk_sum);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
  }
}

// r_n = w_n > 0 ? exp(ln w_n + LP[n][j] - comp_max_j) : 0
__global__ void responsibilities_kernel(const double *__restrict__ LP, int64_t n, int m, int j,
                                        const double *__restrict__ w, double comp_max_j, double *r_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wn = w ? w[i] : 1.0;
    r_out[i] = wn > 0.0 ? exp(log(wn) + LP[i * m + j] - comp_max_j) : 0.0;
  }
}

void launch_responsibilities(const Launcher &L, const double *LP, int64_t n, int m, int j, const double *w,
                             double comp_max_j, double *r_out) {
  if (n <= 0) return;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < (int64_t)L.sms * 8 ? want : (int64_t)L.sms * 8);
  responsibilities_kernel<<<blocks, 256, 0, L.stream>>>(LP, n, m, j, w, comp_max_j, r_out);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_weighted_sum(const Launcher &L, const double *v, const double *w, int64_t n, double *out, int accumulate) {
  weighted_sum_kernel<<<1, 1024, 0, L.stream>>>(v, w, n, out, accumulate);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// ---- single-pass mixture EM: running per-component maxima with rescaling of what has been accumulated so far --------
// The reference scales the responsibilities of component j by exp(-max_n (ln w_n + lp_nj)) over the WHOLE dataset
// (mix.rs:312-318).  One pass over the samples only knows the maximum so far, so every accumulator of component j is
// kept relative to the running maximum and multiplied by exp(old - new) whenever a chunk raises it: the statistics are
// linear in the weights, the result is the reference's up to rounding.
__global__ void mix_update_max_kernel(int m, double *run_max, const double *__restrict__ chunk_max, double *factor) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double old = run_max[j], c = chunk_max[j];
  if (!(c > old)) {  // also when the chunk had no positive-weight sample (c = -inf)
    factor[j] = 1.0;
    return;
  }
  run_max[j] = c;
  factor[j] = old == -INFINITY ? 0.0 : exp(old - c);
}

void launch_mix_update_max(const Launcher &L, int m, double *run_max, const double *chunk_max, double *factor) {
  mix_update_max_kernel<<<(m + 63) / 64, 64, 0, L.stream>>>(m, run_max, chunk_max, factor);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

__global__ void scale_by_kernel(double *buf, int64_t len, const double *__restrict__ factor, int stride4) {
  const double f = factor[0];
  if (f == 1.0) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    if (!stride4 || (i & 3) != 3) buf[i] = f == 0.0 ? 0.0 : buf[i] * f;
}

void launch_scale_by(const Launcher &L, double *buf, int64_t len, const double *factor, int stride4) {
  if (len <= 0) return;
  const int64_t want = (len + 255) / 256;
  const int blocks = (int)(want < (int64_t)L.sms * 8 ? want : (int64_t)L.sms * 8);
  scale_by_kernel<<<blocks, 256, 0, L.stream>>>(buf, len, factor, stride4);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// one warp per sample row: r, W = r V (in place), WZ = r Z, running column maxima of |W| per warp slot in shared memory
__global__ void __launch_bounds__(256) mix_weight_kernel(const double *__restrict__ LP, int m, int j,
                                                         const double *__restrict__ w, const double *__restrict__ run_max,
                                                         int rows, int rows_pad, int kkp, int kp, double *V,
                                                         const double *__restrict__ Z, double *WZ, double *r_out,
                                                         unsigned long long *colmax) {
  extern __shared__ double cm_sh[];  // [8 warps][kkp]
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  double *cm = cm_sh + (size_t)wi * kkp;
  if (colmax)
    for (int q = lane; q < kkp; q += 32) cm[q] = 0.0;
  const double mx = run_max[j];
  for (int row = blockIdx.x * 8 + wi; row < rows_pad; row += gridDim.x * 8) {
    double r = 0.0;
    if (row < rows) {
      const double wn = w ? w[row] : 1.0;
      r = wn > 0.0 ? exp(log(wn) + LP[(int64_t)row * m + j] - mx) : 0.0;
    }
    if (lane == 0) r_out[row] = r;
    double *v = V + (int64_t)row * kkp;
    for (int q = 2 * lane; q < kkp; q += 64) {
      double2 x = *reinterpret_cast<double2 *>(v + q);
      x.x *= r;
      x.y *= r;
      *reinterpret_cast<double2 *>(v + q) = x;
      if (colmax) {
        cm[q] = fmax(cm[q], fabs(x.x));
        cm[q + 1] = fmax(cm[q + 1], fabs(x.y));
      }
    }
    for (int a = lane; a < kp; a += 32) WZ[(int64_t)row * kp + a] = r * Z[(int64_t)row * kp + a];
  }
  if (colmax) {
    __syncthreads();
    for (int q = threadIdx.x; q < kkp; q += blockDim.x) {
      double mq = 0.0;
      for (int s = 0; s < 8; ++s) mq = fmax(mq, cm_sh[(size_t)s * kkp + q]);
      if (mq > 0.0) atomicMax(colmax + q, (unsigned long long)__double_as_longlong(mq));
    }
  }
}

void launch_mix_weight(const Launcher &L, const double *LP, int m, int j, const double *w, const double *run_max, int rows,
                       int rows_pad, int kkp, int kp, double *V, const double *Z, double *WZ, double *r,
                       unsigned long long *colmax) {
  if (rows_pad <= 0) return;
  const size_t smem = colmax ? (size_t)8 * kkp * sizeof(double) : 0;
  REQUIRE(smem <= 200 * 1024, "mixture EM: state_size too large for the fused column maxima");
  static PerDeviceOnce configured;
  if (configured.need())
    CUDA_CHECK(cudaFuncSetAttribute(mix_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int64_t blocks = (rows_pad + 7) / 8;
  const int64_t cap = (int64_t)L.sms * (smem > 96 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4));
  if (blocks > cap) blocks = cap;
  mix_weight_kernel<<<(unsigned)blocks, 256, smem, L.stream>>>(LP, m, j, w, run_max, rows, rows_pad, kkp, kp, V, Z, WZ, r,
                                                              colmax);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

__global__ void mix_rescale_stats_kernel(double *stats, int64_t len_linear, double *scalars,
                                         const double *__restrict__ local_max, const double *__restrict__ global_max,
                                         int j) {
  const double lm = local_max[j], gm = global_max[j];
  if (lm == gm) return;
  const double f = lm == -INFINITY ? 0.0 : exp(lm - gm);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len_linear; i += (int64_t)gridDim.x * blockDim.x)
    stats[i] = f == 0.0 ? 0.0 : stats[i] * f;
  if (blockIdx.x == 0 && threadIdx.x < 4) {  // SC_SQERR, SC_DEV2, SC_LLK, SC_SUMW are linear in the weights
    scalars[threadIdx.x] = f == 0.0 ? 0.0 : scalars[threadIdx.x] * f;
  }
}

void launch_mix_rescale_stats(const Launcher &L, double *stats, int64_t len_linear, double *scalars, const double *local_max,
                              const double *global_max, int j) {
  const int64_t want = (len_linear + 255) / 256;
  const int blocks = (int)(want < (int64_t)L.sms * 8 ? want : (int64_t)L.sms * 8);
  mix_rescale_stats_kernel<<<blocks > 0 ? blocks : 1, 256, 0, L.stream>>>(stats, len_linear, scalars, local_max, global_max, j);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
