// Probe: how fast does the FP64 pipe run the RANK-1 UPDATE pattern of the per-sample solve (A[i][j] -= f[i] * cv[j], a
// 4 x 8 register tile per thread, the 12 operands fresh from shared memory every step) compared with (a) the
// register-resident DFMA peak pattern c = fma(c, a, b) and (b) the same update as rank-4 DMMA (m8n8k4) over 16 tiles?
// Prints FMA rates in TFLOP/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rank1_probe tools/rank1_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) rank1_kernel(double *out, int iters, int from_smem) {
  __shared__ __align__(16) double ops[2][64 + 96];
  for (int i = threadIdx.x; i < 2 * 160; i += blockDim.x) (&ops[0][0])[i] = 1e-3 * (i % 7);
  __syncthreads();
  double A[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) A[i][j] = i + j;
  const int tr = threadIdx.x % 16, tc = (threadIdx.x / 16) % 8;
  double f[4] = {1e-3, 2e-3, 3e-3, 4e-3}, cv[8] = {1e-3, 2e-3, 3e-3, 4e-3, 5e-3, 6e-3, 7e-3, 8e-3};
  for (int it = 0; it < iters; ++it) {
    if (from_smem) {
      const double *ex = ops[it & 1];
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(ex + 64 + tr * 6 + i);
        f[i] = v.x; f[i + 1] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(ex + tc * 8 + j);
        cv[j] = v.x; cv[j + 1] = v.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) A[i][j] = fma(-f[i], cv[j], A[i][j]);
    if (!from_smem) {  // rotate the operands so nothing is loop invariant
      const double t = f[0]; f[0] = f[1]; f[1] = f[2]; f[2] = f[3]; f[3] = cv[0];
#pragma unroll
      for (int j = 0; j < 7; ++j) cv[j] = cv[j + 1];
      cv[7] = t;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += A[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// a warp holds 2 x 8 tiles of 8 x 8 (16 x 64 of the matrix): rank-4 update = 16 DMMA, 2 A fragments + 8 B fragments from smem
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) rank4_dmma_kernel(double *out, int iters) {
  __shared__ __align__(16) double S[2][64 * 4];
  for (int i = threadIdx.x; i < 2 * 256; i += blockDim.x) (&S[0][0])[i] = 1e-3 * (i % 5);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 3;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
    const double *s = S[it & 1];
    double a[2], b[8];
#pragma unroll
    for (int r = 0; r < 2; ++r) a[r] = s[((2 * w + r) * 8 + lane / 4) * 4 + lane % 4];
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) b[cb] = s[(cb * 8 + lane / 4) * 4 + lane % 4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cb = 0; cb < 8; ++cb) dmma884(c[r * 8 + cb][0], c[r * 8 + cb][1], -a[r], b[cb]);
  }
  double s2 = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s2 += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s2;
}

// The serial chain of one pivot inside a panel, alone on an SM sub-partition: broadcast the pivot from the diagonal lane,
// 1/sqrt (MUFU + third-order step), scale the lane's column entries, form the next pivot on the diagonal lane.
// One warp per CTA, one CTA per SM: cycles per pivot = the latency floor of the elimination for one sample.
__device__ __forceinline__ double probe_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
__global__ void chain_kernel(double *out, long long *cycles, int iters) {
  const int lane = threadIdx.x & 31;
  double a0 = 2.0 + lane * 1e-3, a1 = 3.0 + lane * 1e-3, d = 1.5;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double dq = __shfl_sync(0xffffu, d, it & 15);
    const double rinv = probe_rsqrt(dq);
    const double own0 = a0 * rinv, own1 = a1 * rinv;
    d = fma(-own1, own1, a1 + 4.0);   // next pivot (kept positive)
    a0 = fma(-own0, own1, a0) + 1e-3;
    a1 = own1 * rinv + 2.5;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + d;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class F> static float time_ms(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}

int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 2 * 256 * 2));
  const int iters = 200000;
  const double fl1 = 2.0 * 32 * iters * 256.0 * 2 * sms;   // rank-1: 32 FMA per thread per step
  float a = time_ms([&] { rank1_kernel<8><<<sms * 2, 256>>>(out, iters, 0); });
  float b = time_ms([&] { rank1_kernel<8><<<sms * 2, 256>>>(out, iters, 1); });
  const double fl4 = 2.0 * 16 * 256 * iters * 8.0 * 2 * sms;  // rank-4 DMMA: 16 x 256 FMA per warp per step
  float c = time_ms([&] { rank4_dmma_kernel<8><<<sms * 2, 256>>>(out, iters); });
  long long *cyc; CK(cudaMalloc(&cyc, sizeof(long long)));
  const int chain_iters = 100000;
  chain_kernel<<<sms, 32>>>(out, cyc, chain_iters); CK(cudaDeviceSynchronize());
  chain_kernel<<<sms, 32>>>(out, cyc, chain_iters); CK(cudaDeviceSynchronize());
  long long hc = 0; CK(cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
  printf("{\"pivot_chain_cycles_alone\": %.1f}\n", (double)hc / chain_iters);
  printf("{\"rank1_regs_tflops\": %.2f, \"rank1_smem_operands_tflops\": %.2f, \"rank4_dmma_smem_operands_tflops\": %.2f, \"warps_per_sm\": 16}\n",
         fl1 / a / 1e9, fl1 / b / 1e9, fl4 / c / 1e9);
  return 0;
}
