// finish.cu — M-step row solves:  (A_i + tau I) c_i = B_i  for every output dimension i.
//
// Reference: ppca_model.rs:294-324 — total_second_moment + prior.transformation_precision() * I, then
// `.qr().solve(&cross_moment_row)` with a fallback to the old row when the system cannot be solved.
// A_i is a positive-weighted sum of SPD matrices (z z^T + Sigma_n), so it is SPD whenever at least one
// positive-weight sample observes dimension i; nalgebra's QR returns None only on an exactly-zero pivot,
// i.e. when A_i + tau I is the zero matrix (empty dimension, tau = 0).  We therefore solve by Cholesky
// (same solution to rounding) and keep the old row when the matrix is all-zero (flag 1) or when a pivot
// is not a positive finite number (flag 2).
#include "common.cuh"
#include "mma.cuh"

namespace ppca {

__global__ void __launch_bounds__(128) row_solve_kernel(int d, int k, int kp, int kkp, const double *__restrict__ A,
                                                        const double *__restrict__ B, double tau,
                                                        const double *__restrict__ Cold, double *Cnew, int *flags) {
  extern __shared__ double smem_fin[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int ldk = k | 1;
  const int per_warp = ((k + 1) * ldk + 1) & ~1;
  double *M = smem_fin + (size_t)wi * per_warp;
  for (int i = blockIdx.x * warps + wi; i < d; i += gridDim.x * warps) {
    const double *Ai = A + (int64_t)i * kkp;
    bool nz = false;
    for (int p = 0; p < k; ++p) {
      const int off = tri_row_off(p, k) - p;
      for (int b = p + lane; b < k; b += 32) {
        const double v = Ai[off + b] + (b == p ? tau : 0.0);
        nz |= (v != 0.0);
        M[b * ldk + p] = v;
      }
    }
    for (int q = lane; q < k; q += 32) M[k * ldk + q] = B[(int64_t)i * kp + q];
    nz = __any_sync(0xffffffffu, nz);
    __syncwarp();
    int flag = nz ? 0 : 1;
    if (nz) {
      for (int p = 0; p < k; ++p) {
        const double dpp = M[p * ldk + p];
        if (!(dpp > 0.0) || !isfinite(dpp)) { flag = 2; break; }
        const double inv = 1.0 / sqrt(dpp);
        __syncwarp();
        for (int rr = p + 1 + lane; rr <= k; rr += 32) M[rr * ldk + p] *= inv;
        if (lane == 0) M[p * ldk + p] = inv;  // keep 1 / L_pp on the diagonal
        __syncwarp();
        for (int cc = p + 1; cc < k; ++cc) {
          const double lcp = M[cc * ldk + p];
          for (int rr = cc + lane; rr <= k; rr += 32) M[rr * ldk + cc] = fma(-M[rr * ldk + p], lcp, M[rr * ldk + cc]);
        }
        __syncwarp();
      }
    }
    if (flag == 0) {
      // back substitution L^T c = u (u in row k)
      for (int p = k - 1; p >= 0; --p) {
        const double cp = M[k * ldk + p] * M[p * ldk + p];
        __syncwarp();
        if (lane == 0) M[k * ldk + p] = cp;
        for (int q = lane; q < p; q += 32) M[k * ldk + q] = fma(-M[p * ldk + q], cp, M[k * ldk + q]);
        __syncwarp();
      }
      for (int q = lane; q < k; q += 32) Cnew[(int64_t)i * k + q] = M[k * ldk + q];
    } else {
      for (int q = lane; q < k; q += 32) Cnew[(int64_t)i * k + q] = Cold[(int64_t)i * kp + q];
    }
    if (lane == 0 && flags) flags[i] = flag;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(1024) mstep_guard_kernel(int d, int k, int kkp, const double *__restrict__ A,
                                                           const double *__restrict__ totals,
                                                           const double *__restrict__ smax, double coef_terms,
                                                           const unsigned int *__restrict__ unsafe, double *scalars) {
  __shared__ unsigned int cnt;
  if (threadIdx.x == 0) cnt = 0u;
  __syncthreads();
  if (smax) {
    unsigned int mine = 0u;
    for (int idx = threadIdx.x; idx < d * k; idx += blockDim.x) {
      const int i = idx / k, a = idx % k;
      if (!(totals[i] > 0.0)) continue;  // nobody (with weight) observed this dimension: the row is kept as is
      const int q = tri_row_off(a, k);
      if (coef_terms * smax[q] > A[(int64_t)i * kkp + q]) ++mine;
    }
    if (mine) atomicAdd(&cnt, mine);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    scalars[SC_UNSAFE_E] = unsafe ? (double)unsafe[0] : 0.0;
    scalars[SC_UNSAFE_M] = (double)cnt;
  }
}

void launch_mstep_guard(const Launcher &L, int d, int k, const double *statA, const double *totals, const double *smax,
                        double coef_terms, const unsigned int *unsafe, double *scalars) {
  Shape s(d, k);
  mstep_guard_kernel<<<1, 1024, 0, L.stream>>>(d, k, s.kkp, statA, totals, smax, coef_terms, unsafe, scalars);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

__global__ void scale_max_kernel(const double *__restrict__ scale, int n, double *smax) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) smax[q] = fmax(smax[q], scale[q]);
}

void launch_scale_max(const Launcher &L, const double *scale, int n, double *smax) {
  if (n <= 0) return;
  scale_max_kernel<<<(n + 255) / 256, 256, 0, L.stream>>>(scale, n, smax);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_row_solve(const Launcher &L, int d, int k, const double *statA, const double *statB, double tau,
                      const double *Cold_pad, double *Cnew, int *flags) {
  if (d <= 0 || k <= 0) return;
  Shape s(d, k);
  const int ldk = k | 1;
  const size_t per_warp = (size_t)((((k + 1) * ldk) + 1) & ~1) * sizeof(double);
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  REQUIRE(per_warp * warps <= 227 * 1024, "state_size %d too large for the row-solve kernel", k);
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(row_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int blocks = (d + warps - 1) / warps;
  row_solve_kernel<<<blocks, warps * 32, per_warp * warps, L.stream>>>(d, k, s.kp, s.kkp, statA, statB, tau, Cold_pad,
                                                                      Cnew, flags);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}


// ---------------------------------------------------------------------------------------------
// Mean prior (prior.rs:97-110 smooth_mean): solve (Lambda_0 + diag(totals / sigma'^2)) mu = Lambda_0 m_0 + diag(...) mu_hat.
// The reference calls total_precision.qr().solve; the matrix is SPD whenever the prior covariance is, so the device
// path is a blocked right-looking Cholesky (32-wide panels: one CTA factors the diagonal block and the panel under it,
// a grid of 32 x 32 tiles applies the rank-32 update to the trailing lower triangle) and two triangular sweeps.  Round 1
// ran an O(d^3) column-strided Householder QR on one host thread (seconds per iteration at d = 2048); that routine is kept
// as the fallback for a precision matrix that is not positive definite.
// ---------------------------------------------------------------------------------------------
constexpr int CH_NB = 32;

// factors the diagonal block A[j0:j0+nb, j0:j0+nb] in shared memory (one CTA) and writes L back
__global__ void __launch_bounds__(256) chol_diag_kernel(double *A, int n, int j0, int *fail) {
  __shared__ double L[CH_NB][CH_NB + 1];
  const int nb = min(CH_NB, n - j0);
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
    const int r = idx / CH_NB, c = idx % CH_NB;
    L[r][c] = (r < nb && c < nb && c <= r) ? A[(int64_t)(j0 + r) * n + j0 + c] : 0.0;
  }
  __syncthreads();
  for (int p = 0; p < nb; ++p) {
    const double dpp = L[p][p];
    if (!(dpp > 0.0) || !isfinite(dpp)) {  // uniform: every thread reads the same pivot
      if (tid == 0) *fail = 1;
      return;
    }
    __syncthreads();
    const double root = sqrt(dpp), inv = 1.0 / root;
    if (tid == 0) L[p][p] = root;
    for (int r = p + 1 + tid; r < nb; r += 256) L[r][p] *= inv;
    __syncthreads();
    const int m = nb - p - 1;
    for (int idx = tid; idx < m * m; idx += 256) {
      const int r = p + 1 + idx / m, c = p + 1 + idx % m;
      if (c <= r) L[r][c] = fma(-L[r][p], L[c][p], L[r][c]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += 256) {
    const int r = idx / nb, c = idx % nb;
    if (c <= r) A[(int64_t)(j0 + r) * n + j0 + c] = L[r][c];
  }
}

// panel under the factored diagonal block: A[i, j0:j0+nb] <- A[i, j0:j0+nb] L^-T, one thread per row
__global__ void __launch_bounds__(256) chol_panel_kernel(double *A, int n, int j0, const int *fail) {
  __shared__ double L[CH_NB][CH_NB + 1];
  if (*fail) return;
  const int nb = min(CH_NB, n - j0);
  const int tid = threadIdx.x;
  for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
    const int r = idx / CH_NB, c = idx % CH_NB;
    L[r][c] = (r < nb && c < nb && c <= r) ? A[(int64_t)(j0 + r) * n + j0 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int i = j0 + nb + blockIdx.x * 256 + tid; i < n; i += gridDim.x * 256) {
    double x[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) x[c] = c < nb ? A[(int64_t)i * n + j0 + c] : 0.0;
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
      double v = x[c];
#pragma unroll
      for (int q = 0; q < CH_NB; ++q)
        if (q < c) v = fma(-x[q], L[c][q], v);
      x[c] = v / L[c][c];
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c)
      if (c < nb) A[(int64_t)i * n + j0 + c] = x[c];
  }
}

// trailing update of the lower triangle: A[i][j] -= sum_q A[i][j0+q] A[j][j0+q], tiles of 32 x 32, i >= j >= j0 + nb
__global__ void __launch_bounds__(256) chol_update_kernel(double *A, int n, int j0, int nb) {
  __shared__ double Pi[32][CH_NB + 1], Pj[32][CH_NB + 1];
  const int t0 = j0 + nb;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int i0 = t0 + 32 * ti, jj0 = t0 + 32 * tj;
  for (int idx = threadIdx.x; idx < 32 * CH_NB; idx += 256) {
    const int r = idx / CH_NB, q = idx % CH_NB;
    Pi[r][q] = (i0 + r < n && q < nb) ? A[(int64_t)(i0 + r) * n + j0 + q] : 0.0;
    Pj[r][q] = (jj0 + r < n && q < nb) ? A[(int64_t)(jj0 + r) * n + j0 + q] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 32; idx += 256) {
    const int r = idx / 32, c = idx % 32;
    const int i = i0 + r, j = jj0 + c;
    if (i < n && j < n && j <= i) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < CH_NB; ++q) s = fma(Pi[r][q], Pj[c][q], s);
      A[(int64_t)i * n + j] -= s;
    }
  }
}

// one CTA: forward substitution L y = b, then backward L^T x = y, in place on b (O(n^2), bandwidth of L from L2)
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double *__restrict__ A, int n, double *b) {
  __shared__ double piv;
  const int tid = threadIdx.x;
  for (int p = 0; p < n; ++p) {
    if (tid == 0) {
      b[p] /= A[(int64_t)p * n + p];
      piv = b[p];
    }
    __syncthreads();
    const double bp = piv;
    for (int i = p + 1 + tid; i < n; i += 1024) b[i] = fma(-A[(int64_t)i * n + p], bp, b[i]);
    __syncthreads();
  }
  for (int p = n - 1; p >= 0; --p) {
    if (tid == 0) {
      b[p] /= A[(int64_t)p * n + p];
      piv = b[p];
    }
    __syncthreads();
    const double bp = piv;
    for (int i = tid; i < p; i += 1024) b[i] = fma(-A[(int64_t)p * n + i], bp, b[i]);
    __syncthreads();
  }
}

// P <- prior precision + diag(totals / noise_sq) ; rhs_i <- (P0 m0)_i + (totals_i / noise_sq) mu_hat_i
__global__ void mean_prior_setup_kernel(double *P, int n, const double *__restrict__ m0, const double *__restrict__ totals,
                                        const double *__restrict__ mu_hat, double noise_sq, double *rhs) {
  const int i = blockIdx.x;
  __shared__ double sh[8];
  double acc = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) acc = fma(P[(int64_t)i * n + j], m0[j], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    const double t = totals[i] / noise_sq;
    rhs[i] = s + t * mu_hat[i];
    P[(int64_t)i * n + i] += t;
  }
}

// returns false (and leaves rhs untouched on the host side) when the matrix is not positive definite
void launch_mean_prior_solve(const Launcher &L, double *P_dev, int n, const double *m0_dev, const double *totals_dev,
                             const double *mu_hat_dev, double noise_sq, double *rhs_dev, int *fail_dev) {
  CUDA_CHECK(cudaMemsetAsync(fail_dev, 0, sizeof(int), L.stream));
  mean_prior_setup_kernel<<<n, 256, 0, L.stream>>>(P_dev, n, m0_dev, totals_dev, mu_hat_dev, noise_sq, rhs_dev);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  for (int j0 = 0; j0 < n; j0 += CH_NB) {
    const int nb = n - j0 < CH_NB ? n - j0 : CH_NB;
    const int below = n - j0 - nb;
    int pblocks = (below + 255) / 256;
    if (pblocks < 1) pblocks = 1;
    chol_diag_kernel<<<1, 256, 0, L.stream>>>(P_dev, n, j0, fail_dev);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
    if (below > 0) {
      chol_panel_kernel<<<pblocks, 256, 0, L.stream>>>(P_dev, n, j0, fail_dev);
      CUDA_CHECK(cudaGetLastError());
      ++*L.launch_counter;
      const int tiles = (below + 31) / 32;
      chol_update_kernel<<<dim3(tiles, tiles), 256, 0, L.stream>>>(P_dev, n, j0, nb);
      CUDA_CHECK(cudaGetLastError());
      ++*L.launch_counter;
    }
  }
  chol_solve_kernel<<<1, 1024, 0, L.stream>>>(P_dev, n, rhs_dev);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
