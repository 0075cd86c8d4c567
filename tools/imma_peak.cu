// Legacy-path tensor throughput probes on sm_100a (register resident, no memory traffic):
//   IMMA  mma.sync.m16n8k32.s8.s8.s32   (int8 -> int32)
//   HMMA  mma.sync.m16n8k16.bf16 -> f32
// Used to decide whether an int8-sliced (Ozaki) FP64-equivalent contraction on mma.sync is worth building.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) imma_kernel(int *out, int iters, int a0, int b0) {
  int c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  int a[4] = {a0, a0 + 1, a0 + 2, a0 + 3};
  int b[2] = {b0, b0 + 1};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) hmma_kernel(float *out, int iters, unsigned a0, unsigned b0) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a[4] = {a0, a0, a0, a0};
  unsigned b[2] = {b0, b0};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static float best_ms(F f, int reps) {
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
  void *out; CK(cudaMalloc(&out, 4 * sms * 8 * 256));
  const int iters = 20000;
  double best_i = 0, best_h = 0;
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    float ms = best_ms([&] { imma_kernel<8><<<sms * ctas, 256>>>((int *)out, iters, 1, 2); }, 3);
    double ops = 2.0 * 16 * 8 * 32 * 8 * (double)iters * 8 * sms * ctas;
    if (ops / ms * 1e-9 > best_i) best_i = ops / ms * 1e-9;
    fprintf(stderr, "imma ctas/SM=%d nacc=8: %.1f TOPS\n", ctas, ops / ms * 1e-9);
    ms = best_ms([&] { imma_kernel<16><<<sms * ctas, 256>>>((int *)out, iters, 1, 2); }, 3);
    ops = 2.0 * 16 * 8 * 32 * 16 * (double)iters * 8 * sms * ctas;
    if (ops / ms * 1e-9 > best_i) best_i = ops / ms * 1e-9;
    fprintf(stderr, "imma ctas/SM=%d nacc=16: %.1f TOPS\n", ctas, ops / ms * 1e-9);
    ms = best_ms([&] { hmma_kernel<8><<<sms * ctas, 256>>>((float *)out, iters, 0x3f803f80u, 0x3f803f80u); }, 3);
    double fl = 2.0 * 16 * 8 * 16 * 8 * (double)iters * 8 * sms * ctas;
    if (fl / ms * 1e-9 > best_h) best_h = fl / ms * 1e-9;
    ms = best_ms([&] { hmma_kernel<16><<<sms * ctas, 256>>>((float *)out, iters, 0x3f803f80u, 0x3f803f80u); }, 3);
    fl = 2.0 * 16 * 8 * 16 * 16 * (double)iters * 8 * sms * ctas;
    if (fl / ms * 1e-9 > best_h) best_h = fl / ms * 1e-9;
  }
  printf("{\"gpu\": \"%s\", \"imma_s8_tops\": %.1f, \"hmma_bf16_tflops\": %.1f}\n", p.name, best_i, best_h);
  return 0;
}
