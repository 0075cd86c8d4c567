// ibitgemm.cu — the masked-Gram contraction evaluated EXACTLY on the int8 tensor path.
//
//   Out[M x Nq] (+)= Bits[M x K] * Bmat[K x Nq]
//
// The left operand is a {0,1} bit matrix, so every output is a plain sum of selected FP64 numbers.  Each column
// of Bmat is scaled by a power of two s_q > max|column| and split into T balanced base-256 digits
// (int8, [-128, 127]; the leading digit stays within +-65):  v = rint(x / s_q * 2^(8T-2)) = sum_t digit_t 256^(T-1-t),
// i.e. x is kept to 8T-2 bits below its column scale (46 bits for T = 6, 54 for T = 7).
// One int8 x int8 -> int32 tensor MMA per digit plane accumulates sum_k bit * digit EXACTLY (|sum| <= 128 K < 2^31),
// and the epilogue recombines the T planes in FP64 by Horner.  The only error is that 2^-(8T-1) s_q rounding of each
// term (below the rounding error of an FP64 dot product for T >= 6); there is no accumulation error.
// B200: legacy mma.sync.m16n8k32.s8 runs at 1144 TOP/s (profiles/r01_imma_peak.json), i.e. 163 TFLOP/s
// FP64-equivalent at T = 7 against 37 TFLOP/s for DMMA.  Same operands, same references as bitgemm.cu:
//   E-step Gs = Mask * Ksym (output_covariance.rs:57-59 after :123-131), M-step A += Mask^T W (ppca_model.rs:297-306).
#include <cstdio>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

// ---------------------------------------------------------------------------------------------
// column scales and digit planes
// ---------------------------------------------------------------------------------------------
__global__ void colmax_kernel(const double *__restrict__ B, int64_t ldb, int K, int Nq, unsigned long long *cm) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Nq) return;
  const int per = (K + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(K, lo + per);
  double m = 0.0;
  for (int r = lo; r < hi; ++r) m = fmax(m, fabs(B[(int64_t)r * ldb + q]));
  if (m > 0.0) atomicMax(cm + q, (unsigned long long)__double_as_longlong(m));  // order-independent
}

// position of K-row r (0..31) inside the 32-byte group: thread t of a quad reads bytes [8t, 8t+8) as (b0, b1)
__host__ __device__ constexpr int perm32(int r) { return ((r & 15) >> 2) * 8 + (r >> 4) * 4 + (r & 3); }

// rint(xs) as T balanced base-256 digits packed little-endian (byte t = digit of weight 256^t); |xs| <= 2^(8T-2)
template <int T>
__device__ __forceinline__ unsigned long long digit_bytes(double xs) {
  constexpr unsigned long long BIAS = 0x0080808080808080ull & ((1ull << (8 * (T - 1))) - 1ull);  // +128 below the top digit
  const long long v = __double2ll_rn(xs) + (long long)BIAS;
  return (unsigned long long)v ^ BIAS;  // low digits: byte - 128 == byte ^ 0x80 as int8; top digit: signed as is
}

template <int T>
__global__ void __launch_bounds__(128) slice_kernel(const double *__restrict__ B, int64_t ldb, int K, int Nq,
                                                    const unsigned long long *__restrict__ cm, int8_t *out,
                                                    double *scale) {
  const int q = blockIdx.x * 128 + threadIdx.x;
  const int kb = blockIdx.y;
  if (q >= Nq) return;
  const double m = __longlong_as_double((long long)cm[q]);
  int ex = 0;
  if (m > 0.0) frexp(m, &ex);  // m = f 2^ex, f in [0.5, 1)  =>  |x| < 2^ex
  const double s = ldexp(1.0, ex);
  if (kb == 0) scale[q] = s;
  const double inv = ldexp(1.0, 8 * T - 2 - ex);
  uint32_t words[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) words[t][j] = 0u;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int row = kb * 32 + r;
    const double x = (row < K) ? B[(int64_t)row * ldb + q] : 0.0;
    const int p = perm32(r);
    const unsigned long long v = digit_bytes<T>(x * inv);
#pragma unroll
    for (int t = 0; t < T; ++t)  // plane 0 = most significant digit
      words[t][p >> 2] |= (uint32_t)((v >> (8 * (T - 1 - t))) & 0xffull) << (8 * (p & 3));
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    uint4 *dst = reinterpret_cast<uint4 *>(out + (((int64_t)kb * T + t) * Nq + q) * 32);
    dst[0] = make_uint4(words[t][0], words[t][1], words[t][2], words[t][3]);
    dst[1] = make_uint4(words[t][4], words[t][5], words[t][6], words[t][7]);
  }
}

size_t sliced_bytes(int kblocks, int Nq, int T) { return (size_t)kblocks * T * Nq * 32; }

void launch_slice(const Launcher &L, const double *Bmat, int64_t ldb, int K, int Nq, int kblocks, int T, int8_t *q,
                  double *scale, unsigned long long *cm) {
  if (Nq <= 0 || kblocks <= 0) return;
  CUDA_CHECK(cudaMemsetAsync(cm, 0, sizeof(unsigned long long) * Nq, L.stream));
  {
    int slabs = K / 256;
    if (slabs < 1) slabs = 1;
    if (slabs > 8 * L.sms) slabs = 8 * L.sms;
    dim3 grid((Nq + 127) / 128, slabs);
    colmax_kernel<<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, cm);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
  }
  dim3 grid((Nq + 127) / 128, kblocks);
  if (T == 6) slice_kernel<6><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, cm, q, scale);
  else if (T == 7) slice_kernel<7><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, cm, q, scale);
  else if (T == 8) slice_kernel<8><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, cm, q, scale);
  else PPCA_THROW(PPCA_ERR_INVALID, "int8 path: T must be 6, 7 or 8 (got %d)", T);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// ---------------------------------------------------------------------------------------------
// the IMMA kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void imma16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 4 mask bits -> 4 bytes of 0/1 (bit j -> byte j): the partial products occupy disjoint bit ranges, no carries
__device__ __forceinline__ uint32_t nib_to_bytes(uint32_t w, int shift) {
  return (((w >> shift) & 0xFu) * 0x00204081u) & 0x01010101u;
}

template <int T_, int WM_, int WN_>
struct IbCfg {
  static constexpr int T = T_, WM = WM_, WN = WN_;
  static constexpr int BM = 32 * WM, NQ = 16 * WN, THREADS = 32 * WM * WN;
  static constexpr int BKB = 4, STAGES = 3;
  static constexpr int STAGE_BYTES = BKB * T * NQ * 32;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) ibitgemm_kernel(IBitGemmArgs a) {
  constexpr int T = Cfg::T, WM = Cfg::WM, BM = Cfg::BM, NQ = Cfg::NQ, THREADS = Cfg::THREADS;
  constexpr int BKB = Cfg::BKB, STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) unsigned char smem_i8[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.x * BM, q0 = blockIdx.y * NQ;

  const int kb_per = (a.kblocks + a.splitk - 1) / a.splitk;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(a.kblocks, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);
  const int nst = (nkb + BKB - 1) / BKB;

  auto load_stage = [&](int stage, int st) {
    unsigned char *sB = smem_i8 + stage * Cfg::STAGE_BYTES;
    const int kb0 = kb_begin + st * BKB;
    constexpr int CHUNKS = BKB * T * NQ * 2;  // 16-byte chunks
    for (int idx = tid; idx < CHUNKS; idx += THREADS) {
      const int half = idx & 1;
      const int qi = (idx >> 1) % NQ;
      const int st_t = (idx >> 1) / NQ;  // s * T + t
      const int s = st_t / T, t = st_t % T;
      const bool ok = (kb0 + s) < kb_end && (q0 + qi) < a.Nq;
      const int8_t *src = ok ? a.Bq + ((((int64_t)(kb0 + s) * T + t) * a.Nq + q0 + qi) * 32 + half * 16) : a.Bq;
      cp_async16(sB + ((st_t * NQ + qi) * 32 + half * 16), src, ok ? 16 : 0);
    }
  };

  // this thread's four bit rows: g, g+8, g+16, g+24 of the warp's 32-row slab
  const uint32_t *brow[4];
  bool brow_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = m0 + wm * 32 + 8 * j + g;
    brow_ok[j] = row < a.M;
    brow[j] = a.bits + (int64_t)(brow_ok[j] ? row : 0) * a.ldbits;
  }
  auto load_words = [&](uint32_t (&w)[BKB][4], int st) {
    const int kb0 = kb_begin + st * BKB;
#pragma unroll
    for (int s = 0; s < BKB; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[s][j] = (brow_ok[j] && (kb0 + s) < kb_end) ? __ldg(brow[j] + kb0 + s) : 0u;
  };

  int acc[2][2][T][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mi][ni][t][e] = 0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nst) load_stage(s, s);
    cp_async_commit();
  }
  uint32_t wcur[BKB][4], wnext[BKB][4];
  if (nst > 0) load_words(wcur, 0);

  for (int it = 0; it < nst; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < nst) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    if (it + 1 < nst) load_words(wnext, it + 1);
    const unsigned char *sB = smem_i8 + (it % STAGES) * Cfg::STAGE_BYTES + ((wn * 16 + g) * 32 + 8 * tq);
#pragma unroll
    for (int s = 0; s < BKB; ++s) {
      if (it * BKB + s >= nkb) break;
      uint32_t af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        af[mi][0] = nib_to_bytes(wcur[s][2 * mi], 4 * tq);
        af[mi][1] = nib_to_bytes(wcur[s][2 * mi + 1], 4 * tq);
        af[mi][2] = nib_to_bytes(wcur[s][2 * mi], 16 + 4 * tq);
        af[mi][3] = nib_to_bytes(wcur[s][2 * mi + 1], 16 + 4 * tq);
      }
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const uint2 b = *reinterpret_cast<const uint2 *>(sB + ((s * T + t) * NQ + 8 * ni) * 32);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) imma16832(acc[mi][ni][t], af[mi], b.x, b.y);
        }
    }
#pragma unroll
    for (int s = 0; s < BKB; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) wcur[s][j] = wnext[s][j];
  }
  cp_async_wait<0>();

  // epilogue: recombine the digit planes in FP64 (Horner, smallest plane first), apply the column scale
  double *out;
  int64_t ldo;
  bool accumulate;
  if (a.splitk > 1) {
    out = a.partials + (int64_t)blockIdx.z * a.M * a.Nq;
    ldo = a.Nq;
    accumulate = a.defer_reduce != 0;
  } else {
    out = a.Out;
    ldo = a.ldo;
    accumulate = a.accumulate != 0;
  }
#pragma unroll
  for (int ni = 0; ni < 2; ++ni) {
    const int col = q0 + wn * 16 + 8 * ni + 2 * tq;
    if (col >= a.Nq) continue;
    const double sc0 = a.scale[col] * (1.0 / 64.0), sc1 = a.scale[col + 1] * (1.0 / 64.0);
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm * 32 + 16 * mi + 8 * h + g;
        if (row >= a.M) continue;
        double v0 = (double)acc[mi][ni][T - 1][2 * h], v1 = (double)acc[mi][ni][T - 1][2 * h + 1];
#pragma unroll
        for (int t = T - 2; t >= 0; --t) {
          v0 = fma(v0, 1.0 / 256.0, (double)acc[mi][ni][t][2 * h]);
          v1 = fma(v1, 1.0 / 256.0, (double)acc[mi][ni][t][2 * h + 1]);
        }
        double2 v = make_double2(v0 * sc0, v1 * sc1);
        double2 *p = reinterpret_cast<double2 *>(out + (int64_t)row * ldo + col);
        if (accumulate) {
          const double2 o = *p;
          v.x += o.x;
          v.y += o.y;
        }
        *p = v;
      }
  }
}

template <int T>
using IbDefault = IbCfg<T, 4, 2>;  // 128 rows x 32 columns x T planes, 8 warps

int ibitgemm_pick_splitk(int M, int Nq, int kblocks, int sms) {
  const int BM = 128, NQ = 32;
  const int64_t tiles = round_up(M, BM) / BM * (round_up(Nq, NQ) / NQ);
  const int64_t max_s = kblocks / 16 > 0 ? kblocks / 16 : 1;  // at least 16 K-blocks (512 rows) per slab
  if (tiles >= sms) {
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 8 && s <= max_s; ++s) {
      const int64_t ctas = tiles * s;
      const double eff = (double)ctas / (double)(round_up(ctas, sms));
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        best = s;
      }
      if (eff >= 0.95) return s;
    }
    return best;
  }
  int64_t s = sms / tiles;
  if (s > max_s) s = max_s;
  return (int)(s < 1 ? 1 : s);
}

template <class Cfg>
static void launch_ib(const Launcher &L, const IBitGemmArgs &a) {
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(ibitgemm_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  dim3 grid((unsigned)(round_up(a.M, Cfg::BM) / Cfg::BM), (unsigned)(round_up(a.Nq, Cfg::NQ) / Cfg::NQ),
            (unsigned)a.splitk);
  ibitgemm_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(V_IMMA);
}

void launch_ibitgemm(const Launcher &L, const IBitGemmArgs &a) {
  REQUIRE(a.Nq % 8 == 0 && a.ldo % 2 == 0, "ibitgemm: Nq must be a multiple of 8, output pitch even");
  REQUIRE(a.splitk >= 1 && (a.splitk == 1 || a.partials != nullptr), "ibitgemm: split-K needs a partials workspace");
  if (a.M <= 0 || a.Nq <= 0) return;
  if (a.T == 6) launch_ib<IbDefault<6>>(L, a);
  else if (a.T == 7) launch_ib<IbDefault<7>>(L, a);
  else if (a.T == 8) launch_ib<IbDefault<8>>(L, a);
  else PPCA_THROW(PPCA_ERR_INVALID, "int8 path: T must be 6, 7 or 8 (got %d)", a.T);
  if (a.splitk > 1 && !a.defer_reduce)
    launch_bitgemm_reduce(L, a.partials, a.splitk, a.M, a.Nq, a.Out, a.ldo, a.accumulate);
}

}  // namespace ppca
