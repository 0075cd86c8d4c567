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

}  // namespace ppca
