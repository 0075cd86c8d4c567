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
                                                            int64_t n, double *out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s = fma(w ? w[i] : 1.0, v[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum(sh[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = s;
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

}  // namespace ppca
