// mma.cuh — FP64 tensor-core and async-copy primitives (sm_100a).
//
// tcgen05.mma has no f64 kind (ptxas rejects kind::f64), so FP64 contractions run on
// mma.sync.m8n8k4.f64, which lowers to one DMMA.8x8x4 SASS instruction.  Fragment layout (PTX ISA,
// "Matrix Fragments for mma.m8n8k4 with .f64"):  with r = lane / 4, c = lane % 4
//   A (8x4, row)  : a    = A[r][c]
//   B (4x8, col)  : b    = B[c][r]          (k = c, n = r)
//   C/D (8x8)     : c0,1 = C[r][2c], C[r][2c+1]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppca {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1.0 if bit `b` of `w` is set else 0.0, built with integer ops only (no I2F on the slow pipe).
__device__ __forceinline__ double bit_to_double(uint32_t w, int b) {
  uint32_t hi = (0u - ((w >> b) & 1u)) & 0x3FF00000u;
  return __hiloint2double((int)hi, 0);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global -> shared, L2 only (.cg); src_bytes in {0,16}: 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes));
}
// 8-byte async copy (.ca only supports 4/8/16)
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ppca
