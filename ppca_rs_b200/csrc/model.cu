// model.cu — per-iteration model staging: zero-padded C, mu and the row-wise symmetric Kronecker table
//   Ksym[i][q(a,b)] = C[i][a] * C[i][b],  a <= b          (d32 x kkp)
// so that the per-sample Gram matrices G_n = C_o^T C_o (output_covariance.rs:57-59 after :123-131) become
// the single contraction Gs = Mask * Ksym (bitgemm.cu).  Symmetry is exploited by keeping only a <= b.
#include "common.cuh"

namespace ppca {

__global__ void prepare_model_kernel(const double *__restrict__ C, const double *__restrict__ mu, int d, int k,
                                     int kp, int kkp, int d32, double *Cpad, double *mupad, double *Ksym) {
  const int i = blockIdx.x;  // one CTA per (padded) output dimension
  const bool live = i < d;
  for (int a = threadIdx.x; a < kp; a += blockDim.x) Cpad[(int64_t)i * kp + a] = (live && a < k) ? C[(int64_t)i * k + a] : 0.0;
  if (threadIdx.x == 0) mupad[i] = live ? mu[i] : 0.0;
  const int kk = k * (k + 1) / 2;
  // q -> (a, b): walk rows; each thread handles a strided set of q
  for (int q = threadIdx.x; q < kkp; q += blockDim.x) {
    double v = 0.0;
    if (live && q < kk) {
      // invert q = a k - a(a-1)/2 + (b - a)
      int a = (int)floor(((2.0 * k + 1.0) - sqrt((2.0 * k + 1.0) * (2.0 * k + 1.0) - 8.0 * q)) * 0.5);
      if (a < 0) a = 0;
      while (a > 0 && tri_row_off(a, k) > q) --a;
      while (a + 1 < k && tri_row_off(a + 1, k) <= q) ++a;
      const int b = a + (q - tri_row_off(a, k));
      v = C[(int64_t)i * k + a] * C[(int64_t)i * k + b];
    }
    Ksym[(int64_t)i * kkp + q] = v;
  }
}

void launch_prepare_model(const Launcher &L, const double *C, const double *mu, int d, int k, double *Cpad,
                          double *mupad, double *Ksym) {
  Shape s(d, k);
  prepare_model_kernel<<<s.d32, 128, 0, L.stream>>>(C, mu, d, k, s.kp, s.kkp, s.d32, Cpad, mupad, Ksym);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
