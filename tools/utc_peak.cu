// tcgen05.mma.kind::i8 issue-rate probe: one CTA per SM, one thread issues back-to-back MMAs (M=128, N, K=32) on
// whatever is in shared memory / tensor memory; no producers, no epilogue.  Prints int8 TOP/s for the SS form
// (A from shared memory) and the TS form (A from tensor memory), N = 192 and 256.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) utc_kernel(int iters, int *sink) {
  extern __shared__ unsigned char raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t *>(raw + (base - smem_u32(raw)))[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  if (threadIdx.x == 0) {
    const uint64_t ad = desc_sw128(base), bd = desc_sw128(base + 16384);
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                       ::"r"(tmem), "r"(tmem + 384u + 8u * k), "l"(bd + 2 * k), "r"(IDESC), "r"(1u), "r"(0u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}"
                       ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(IDESC), "r"(1u), "r"(0u) : "memory");
      }
      if ((it & 15) == 15 || it == iters - 1) {  // bound the number of MMAs in flight
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = lane;
}

template <int N, bool TS> static double run(int sms) {
  const int iters = 4000;
  const size_t smem = 16384 + N * 128 + 1024;
  CK(cudaFuncSetAttribute(utc_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  utc_kernel<N, TS><<<sms, 128, smem>>>(100, nullptr); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0)); utc_kernel<N, TS><<<sms, 128, smem>>>(iters, nullptr); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return 2.0 * 128 * N * 32 * 4.0 * iters * sms / best * 1e-9;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"utcimma_ss_n192_tops\": %.1f, \"utcimma_ss_n256_tops\": %.1f, \"utcimma_ts_n192_tops\": %.1f, \"utcimma_ts_n256_tops\": %.1f}\n",
         p.name, run<192, false>(sms), run<256, false>(sms), run<192, true>(sms), run<256, true>(sms));
  return 0;
}
