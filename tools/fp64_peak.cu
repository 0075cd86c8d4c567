// FP64 peak microbenchmark for the roofline denominator (MEASURED_PEAKS.json has
// only HBM GB/s and bf16 TFLOP/s; tcgen05 has no f64 kind, so the FP64 contractions
// of this engine run on mma.sync DMMA).  Measures:
//   1. DMMA  (mma.sync.m8n8k4.f64)  register-resident, no memory traffic
//   2. DFMA  (fma.rn.f64)           register-resident
//   3. STREAM-like copy (to cross-check hbm_gbs on this box)
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double4 *__restrict__ in, double4 *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

template <class F> static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
  const int iters = 20000;
  double best_dmma = 0, best_dfma = 0; int best_dmma_cfg = 0, best_dfma_cfg = 0;
  // sweep CTAs/SM (occupancy) x accumulators (ILP)
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    {
      float ms = time_ms([&] { dmma_kernel<8><<<sms * ctas, 256>>>(out, iters, 1.0, 1e-9); }, 3);
      double fl = 2.0 * 256 * 8 * (double)iters * 8 /*warps*/ * sms * ctas;
      double tf = fl / ms * 1e-9;
      if (tf > best_dmma) { best_dmma = tf; best_dmma_cfg = ctas * 100 + 8; }
      ms = time_ms([&] { dmma_kernel<16><<<sms * ctas, 256>>>(out, iters, 1.0, 1e-9); }, 3);
      fl = 2.0 * 256 * 16 * (double)iters * 8 * sms * ctas; tf = fl / ms * 1e-9;
      if (tf > best_dmma) { best_dmma = tf; best_dmma_cfg = ctas * 100 + 16; }
      ms = time_ms([&] { dmma_kernel<32><<<sms * ctas, 256>>>(out, iters, 1.0, 1e-9); }, 3);
      fl = 2.0 * 256 * 32 * (double)iters * 8 * sms * ctas; tf = fl / ms * 1e-9;
      if (tf > best_dmma) { best_dmma = tf; best_dmma_cfg = ctas * 100 + 32; }
    }
    {
      float ms = time_ms([&] { dfma_kernel<8><<<sms * ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
      double fl = 2.0 * 8 * (double)iters * 256 * sms * ctas; double tf = fl / ms * 1e-9;
      if (tf > best_dfma) { best_dfma = tf; best_dfma_cfg = ctas * 100 + 8; }
      ms = time_ms([&] { dfma_kernel<16><<<sms * ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
      fl = 2.0 * 16 * (double)iters * 256 * sms * ctas; tf = fl / ms * 1e-9;
      if (tf > best_dfma) { best_dfma = tf; best_dfma_cfg = ctas * 100 + 16; }
    }
  }
  // sustained DMMA: ~3 s back to back at the best config
  double sustained = 0;
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int ctas = best_dmma_cfg / 100; int reps = 0;
    CK(cudaEventRecord(e0));
    for (reps = 0; reps < 200; ++reps) dmma_kernel<16><<<sms * ctas, 256>>>(out, iters, 1.0, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 256 * 16 * (double)iters * 8 * sms * ctas * reps;
    sustained = fl / ms * 1e-9;
  }
  // copy bandwidth, 2 GiB in + 2 GiB out
  size_t nbytes = (size_t)2 << 30; double4 *a, *b;
  CK(cudaMalloc(&a, nbytes)); CK(cudaMalloc(&b, nbytes)); CK(cudaMemset(a, 1, nbytes));
  float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, nbytes / sizeof(double4)); }, 10);
  double gbs = 2.0 * nbytes / ms * 1e-6;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"dmma_tflops\": %.2f, \"dmma_cfg\": %d, "
         "\"dmma_tflops_sustained\": %.2f, \"dfma_tflops\": %.2f, \"dfma_cfg\": %d, \"copy_gbs\": %.1f}\n",
         p.name, sms, clk, best_dmma, best_dmma_cfg, sustained, best_dfma, best_dfma_cfg, gbs);
  return 0;
}
