// tbitgemm.cu — the exact int8-sliced masked-Gram contraction on the 5th-generation tensor cores (tcgen05).
//
//   Out[M x Nq] (+)= Bits[M x K] * Bmat[K x Nq],   Bmat pre-split into T signed 7-bit digit planes (ibitgemm.cu)
//
// Same arithmetic as ibitgemm.cu (exact int32 partial sums, FP64 Horner recombination), but the MMAs are
// tcgen05.mma.kind::i8 issued by one thread, with the accumulators of all T planes of a 128 x 32 output tile in
// tensor memory (N = 32 T columns, double buffered).  Warp-specialised persistent kernel, one CTA per SM:
//   warps 0-3  producers: expand the bit-packed mask rows to int8 {0,1} straight into the SWIZZLE_128B K-major
//              shared-memory layout the MMA reads (the mask never exists as bytes in HBM); the digit-plane tile,
//              stored pre-swizzled, arrives by one cp.async.bulk per stage (mbarrier complete_tx)
//   warp  4    MMA issuer: tcgen05.mma M=128, N=32T, K=32, four per 128-byte K step; tcgen05.commit frees the
//              stage ("empty" mbarrier) and, after the last K step, publishes the accumulator ("tmem_full")
//   warps 5-8  epilogue: tcgen05.ld the T planes of 8 columns at a time, recombine in FP64, scale, store
// References as in bitgemm.cu: E-step Gs = Mask Ksym (output_covariance.rs:57-59 after :123-131),
// M-step A += Mask^T W (ppca_model.rs:297-306).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

namespace tb {

constexpr int MH = 1;             // 128-row MMA halves per tile (MH = 2 with NQ = 16 measured slower: producer bound)
constexpr int BM = 128 * MH, NQ = 32, BKB = 128, STAGES = 4;
constexpr int PRODUCERS = 128, THREADS = 288, THREADS_ATM = 448;
constexpr int ATM_DEPTH = 6, ATM_RING = 8;  // mask-word prefetch distance (own jobs) and ring slots of the ATM kernel

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Whole-warp wait with a single polling lane: 32 lanes spinning on one mbarrier serialise in the shared-memory
// pipe (ncu counted them as ~4e8 bank conflicts per launch), so lane 0 polls and __syncwarp() releases the rest.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory (lane = row, 4 K-bytes per 32-bit column), B from shared memory
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_c),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups of 1024 bytes
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte offset of 16-byte chunk c of row r in the swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t nib4(uint32_t b, int shift) {
  return (((b >> shift) & 0xFu) * 0x00204081u) & 0x01010101u;
}


// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256); the address must be 32-byte aligned
__device__ __forceinline__ void st_global_256(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_global_256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}

// Recombination of the T digit-plane sums d_t (exact int32) of one output: value = sum_t d_t 256^(T-1-t).
// The planes are first merged in exact 64-bit integer arithmetic in groups of three (|group| < 2^48 for the
// three-plane groups), so an output costs ceil(T/3) int64->double conversions and ceil(T/3)-1 FMAs instead of T
// conversions and T-1 FMAs: the FP64 conversion unit was the epilogue's narrowest pipe.  Returns the value itself;
// combine_scale<T>() carries the 2^(-8 (T-1)) normalisation and the 1/64 of the digit scale.
template <int T>
__device__ __forceinline__ double combine_planes(const int (&d)[T]) {
  constexpr int NG = (T + 2) / 3, FIRST = T - 3 * (NG - 1);  // planes in the leading group
  long long g = 0;
#pragma unroll
  for (int t = 0; t < FIRST; ++t) g = g * 256 + (long long)d[t];
  double acc = (double)g;
#pragma unroll
  for (int j = 1; j < NG; ++j) {
    const int t0 = FIRST + 3 * (j - 1);
    const long long gj = ((long long)d[t0] * 256 + (long long)d[t0 + 1]) * 256 + (long long)d[t0 + 2];
    acc = fma(acc, 16777216.0, (double)gj);
  }
  return acc;
}
template <int T>
__device__ __forceinline__ constexpr double combine_scale() {
  // the old Horner form returned value / 256^(T-1); keep that normalisation: 2^(-8 (T-1)) / 64
  double s = 1.0 / 64.0;
  for (int i = 0; i < T - 1; ++i) s *= 1.0 / 256.0;
  return s;
}

}  // namespace tb

struct TBitGemmArgs {
  const uint32_t *bits;
  int64_t ldbits;     // words per bit row
  int nwords;         // valid words per bit row (reads at or beyond return 0)
  const int8_t *Bq;   // [ksteps][qtiles][T*NQ][128]
  const double *scale;
  double *Out;
  int64_t ldo;
  int M, Nq;
  int ksteps;         // 128-bit K steps
  int accumulate;
  double *partials;
  int splitk;
  int defer_reduce;
};

template <int T>
__global__ void __launch_bounds__(tb::THREADS, 1) tbitgemm_kernel(TBitGemmArgs a) {
  using namespace tb;
  constexpr int N = T * NQ;
  constexpr uint32_t A_BYTES = BM * BKB, B_BYTES = N * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int TCOLS = MH * N;  // TMEM columns per accumulator buffer
  static_assert(MH == 1 && (STAGES & (STAGES - 1)) == 0, "producer groups assume one 128-row half and 2^n stages");
  static_assert(2 * TCOLS <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 64 + 1);  // the 64 mask expanders of the group that owns the job + its bulk-copy issue
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int mtiles = (a.M + BM - 1) / BM, qtiles = (a.Nq + NQ - 1) / NQ;
  const int ntiles = mtiles * qtiles * a.splitk;
  const int ks_per = (a.ksteps + a.splitk - 1) / a.splitk;

  if (warp < 4) {
    // ===================== producers =====================
    // The (tile, K step) jobs of this CTA form one sequence.  The four producer warps work as two groups of 64
    // threads that take alternate jobs, so each group has two MMA periods per job and its mask-word loads (issued
    // one own job ahead) are covered.  Per job: one thread issues ONE cp.async.bulk of the pre-swizzled digit-plane
    // tile (completion counted in bytes on the stage's "full" mbarrier); every thread of the group expands two
    // mask rows into the A tile, fences the generic->async proxy and arrives.
    const int grp = warp >> 1, gt = tid & 63;  // group, thread within group: rows gt and gt + 64
    int tiles_left = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    int tile_i = blockIdx.x, ks_i = 0, ks_end_i = 0, qt_i = 0;
    const uint32_t *wrow[2] = {a.bits, a.bits};
    bool row_ok[2] = {false, false};
    auto open_tile = [&]() {
      const int z = tile_i / (mtiles * qtiles), rem = tile_i % (mtiles * qtiles);
      qt_i = rem / mtiles;
      const int mt = rem % mtiles;
      ks_i = z * ks_per;
      ks_end_i = min(a.ksteps, ks_i + ks_per);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = mt * BM + 64 * h + gt;
        row_ok[h] = row < a.M;
        wrow[h] = a.bits + (int64_t)(row_ok[h] ? row : 0) * a.ldbits;
      }
    };
    auto advance = [&]() {  // to the next job of the sequence
      ++ks_i;
      while (tiles_left > 0 && ks_i >= ks_end_i) {
        tile_i += gridDim.x;
        if (--tiles_left > 0) open_tile();
      }
    };
    if (tiles_left > 0) open_tile();
    while (tiles_left > 0 && ks_i >= ks_end_i) {  // skip empty K slabs
      tile_i += gridDim.x;
      if (--tiles_left > 0) open_tile();
    }
    auto load_words = [&](uint32_t (&w)[2][4]) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int wi = 4 * ks_i + j;
          w[h][j] = (row_ok[h] && wi < a.nwords) ? __ldg(wrow[h] + wi) : 0u;
        }
    };
    int job = 0;
    if (grp == 1 && tiles_left > 0) {  // group 1 starts at job 1
      advance();
      job = 1;
    }
    uint32_t w[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) w[h][0] = w[h][1] = w[h][2] = w[h][3] = 0u;
    if (tiles_left > 0) load_words(w);
    while (tiles_left > 0) {
      const int stage = job & (STAGES - 1);
      const uint32_t phase = (uint32_t)(job / STAGES) & 1u;
      mbar_wait_warp(empty_bar(stage), phase ^ 1u);
      const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
      if (gt == 0) {
        const int8_t *src = a.Bq + ((int64_t)ks_i * qtiles + qt_i) * (int64_t)B_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(stage)), "r"(B_BYTES)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB),
            "l"(src), "r"(B_BYTES), "r"(full_bar(stage))
            : "memory");
      }
      uint32_t wc[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) wc[h][j] = w[h][j];
      // move the cursor two jobs on (the other group owns the next one) and fetch that job's mask words now
      advance();
      if (tiles_left > 0) advance();
      job += 2;
      if (tiles_left > 0) load_words(w);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t b = (wc[h][c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + sw128_off(64 * h + gt, c)),
                       "r"(nib4(b, 0)), "r"(nib4(b, 4)), "r"(nib4(b, 8)), "r"(nib4(b, 12))
                       : "memory");
        }
      fence_proxy_async();  // the st.shared rows above must be visible to the tensor core (async proxy)
      mbar_arrive(full_bar(stage));
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int stage = 0, buf = 0;
    uint32_t phase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qtiles);
      const int ks_begin = z * ks_per, ks_end = min(a.ksteps, ks_begin + ks_per);
      mbar_wait_warp(tempty_bar(buf), tphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_c = tmem_base + (uint32_t)(buf * TCOLS);
      if (ks_begin >= ks_end) {  // empty K slab (never produced by the host-side split): publish immediately
        if (lane == 0) tc_commit(tfull_bar(buf));
        __syncwarp();
      }
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        mbar_wait_warp(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
          const uint64_t bdesc = smem_desc_sw128(sB);
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k)
#pragma unroll
            for (int h = 0; h < MH; ++h)
              tc_mma_i8(tmem_c + (uint32_t)(h * N), smem_desc_sw128(sA + h * (128 * BKB)) + (uint64_t)(2 * k),
                        bdesc + (uint64_t)(2 * k), IDESC, (ks > ks_begin || k > 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (ks == ks_end - 1) tc_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++buf == 2) {
        buf = 0;
        tphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue =====================
    int buf = 0;
    uint32_t tphase = 0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qtiles), rem = tile % (mtiles * qtiles);
      const int qt = rem / mtiles, mt = rem % mtiles;
      const bool empty_slab = z * ks_per >= min(a.ksteps, z * ks_per + ks_per);
      double *out;
      int64_t ldo;
      bool accumulate;
      if (a.splitk > 1) {
        out = a.partials + (int64_t)z * a.M * a.Nq;
        ldo = a.Nq;
        accumulate = a.defer_reduce != 0;
      } else {
        out = a.Out;
        ldo = a.ldo;
        accumulate = a.accumulate != 0;
      }
      mbar_wait_warp(tfull_bar(buf), tphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < MH; ++h) {
        const int row = mt * BM + 128 * h + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TCOLS + h * N);
#pragma unroll 1
        for (int qc = 0; qc < NQ / 8; ++qc) {
          int rg[T][8];
#pragma unroll
          for (int t = 0; t < T; ++t) tc_ld8(taddr + (uint32_t)(t * NQ + 8 * qc), rg[t]);
          tc_wait_ld();
          const int q = qt * NQ + 8 * qc;
          if (row < a.M && q < a.Nq) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int d[T];
#pragma unroll
              for (int t = 0; t < T; ++t) d[t] = rg[t][j];
              v[j] = empty_slab ? 0.0 : combine_planes<T>(d) * (a.scale[q + j] * combine_scale<T>());
            }
            // 32-byte accesses: a lane's 8 outputs are two full L2 sectors (16-byte stores wrote every sector twice)
            double *p = out + (int64_t)row * ldo + q;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (accumulate) {
                double o0, o1, o2, o3;
                ld_global_256(p + 4 * j, o0, o1, o2, o3);
                v[4 * j] += o0;
                v[4 * j + 1] += o1;
                v[4 * j + 2] += o2;
                v[4 * j + 3] += o3;
              }
              st_global_256(p + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (++buf == 2) {
        buf = 0;
        tphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// NGRP producer groups (128 threads each, taking jobs round-robin) and NEPI sets of four epilogue warps (each set
// owns NQ / NEPI of the tile's columns); NGRP + NEPI = 3 keeps the CTA at 14 warps.  <2, 1> feeds long K loops
// (mask expansion is the scarce resource), <1, 2> drains short ones (d <= 512: a tile is only a few K steps, the
// FP64 recombination of its 32 x T accumulator columns is what bounds it; ncu: producers idle, epilogue warps busy).
template <int T, int NGRP, int NEPI>
__global__ void __launch_bounds__(tb::THREADS_ATM, 1) tbitgemm_atm_kernel(TBitGemmArgs a) {
  static_assert(32 * (4 * NGRP + 2 + 4 * NEPI) == tb::THREADS_ATM, "role split must add up to the CTA size");
  using namespace tb;
  constexpr int N = T * NQ;
  // the A tile never touches shared memory here: producers write it to tensor memory (tcgen05.st), the MMA
  // reads it from there, so shared memory only carries the digit-plane tile (write once, read once per K step)
  constexpr uint32_t B_BYTES = N * BKB, STAGE_BYTES = B_BYTES;
  constexpr int ACOL = 2 * MH * N;  // first TMEM column of the A stages (32 columns = 128 K-bytes each)
  static_assert(ACOL + 32 * STAGES <= 512, "A stages + two accumulator buffers must fit the 512 TMEM columns");
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int TCOLS = MH * N;  // TMEM columns per accumulator buffer
  static_assert(MH == 1 && (STAGES & (STAGES - 1)) == 0, "producer groups assume one 128-row half and 2^n stages");
  static_assert(2 * TCOLS <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  // warps 0-3 / 4-7: two producer groups (alternate jobs; warp w writes TMEM lane quarter w & 3), warp 8: MMA issuer,
  // warps 9-12: epilogue (quarters 1,2,3,0), warp 13: digit-plane loader
  constexpr int MMA_WARP = 4 * NGRP, EPI_WARP0 = MMA_WARP + 1, LOAD_WARP = EPI_WARP0 + 4 * NEPI;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), PRODUCERS / 32 + 1);  // one elected lane per producer warp + the bulk-copy issue
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128 * NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int mtiles = (a.M + BM - 1) / BM, qtiles = (a.Nq + NQ - 1) / NQ;
  const int ntiles = mtiles * qtiles * a.splitk;
  const int ks_per = (a.ksteps + a.splitk - 1) / a.splitk;

  if (warp < 4 * NGRP) {
    // ===================== producers =====================
    // The (tile, K step) jobs of this CTA form one sequence; producer group g = warp / 4 takes jobs j = g (mod NGRP).
    // Per job every thread of the group expands its mask row (row = 32 (warp & 3) + lane = TMEM lane) to int8
    // {0,1} in registers and writes the 128 bytes to the stage's tensor-memory columns with one tcgen05.st;
    // tcgen05.wait::st, fence, arrive.  The 16 bytes of mask a thread needs per job are fetched DEPTH own jobs
    // ahead with cp.async into a private shared-memory ring (mask rows are strided: latency, not bandwidth).
    constexpr int DEPTH = ATM_DEPTH, RING = ATM_RING;
    const int grp = warp >> 2, gt = tid & 127;  // group, thread within the group = row of the tile
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES) +
                     grp * (RING * 128 * 4);
    struct Cursor {
      int tiles_left, tile, ks, ks_end;
      const uint32_t *wrow;
      bool row_ok;
    };
    auto open_tile = [&](Cursor &c) {
      const int z = c.tile / (mtiles * qtiles), rem = c.tile % (mtiles * qtiles);
      const int mt = rem % mtiles;
      c.ks = z * ks_per;
      c.ks_end = min(a.ksteps, c.ks + ks_per);
      const int row = mt * BM + gt;
      c.row_ok = row < a.M;
      c.wrow = a.bits + (int64_t)(c.row_ok ? row : 0) * a.ldbits;
    };
    auto settle = [&](Cursor &c) {
      while (c.tiles_left > 0 && c.ks >= c.ks_end) {
        c.tile += gridDim.x;
        if (--c.tiles_left > 0) open_tile(c);
      }
    };
    auto step1 = [&](Cursor &c) {  // to the next job of the sequence
      if (c.tiles_left > 0) {
        ++c.ks;
        settle(c);
      }
    };
    auto step = [&](Cursor &c) {  // to this group's next job
#pragma unroll
      for (int g = 0; g < NGRP; ++g) step1(c);
    };
    auto start = [&](Cursor &c) {
      c.tiles_left = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      c.tile = blockIdx.x;
      c.ks = c.ks_end = 0;
      c.wrow = a.bits;
      c.row_ok = false;
      if (c.tiles_left > 0) open_tile(c);
      settle(c);
      if (grp == 1) step1(c);  // group 1 starts at job 1
    };
    auto fetch_words = [&](const Cursor &c, int slot) {  // 4 x 4-byte cp.async (mask rows are only 4-byte aligned)
      const uint32_t dst = smem_u32(ring + (slot * 128 + gt) * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int wi = 4 * c.ks + j;
        const bool ok = c.tiles_left > 0 && c.row_ok && wi < a.nwords;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * j), "l"(ok ? c.wrow + wi : a.bits),
                     "r"(ok ? 4 : 0));
      }
      asm volatile("cp.async.commit_group;" ::);
    };
    Cursor cp, cf;  // processing cursor, fetch cursor (DEPTH own jobs ahead)
    start(cp);
    start(cf);
    int nf = 0;
    for (; nf < DEPTH; ++nf) {
      fetch_words(cf, nf % RING);
      step(cf);
    }
    int job = grp, own = 0;
    while (cp.tiles_left > 0) {
      fetch_words(cf, nf % RING);
      ++nf;
      step(cf);
      const int stage = job & (STAGES - 1);
      const uint32_t phase = (uint32_t)(job / STAGES) & 1u;
      mbar_wait_warp(empty_bar(stage), phase ^ 1u);
      tc_fence_after();
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");  // this job's own mask words have landed
      const uint4 wq = *reinterpret_cast<const uint4 *>(ring + ((own % RING) * 128 + gt) * 4);
      const uint32_t wc[4] = {wq.x, wq.y, wq.z, wq.w};
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = nib4(wc[j >> 3], 4 * (j & 7));
      tc_st32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ACOL + 32 * stage), v);
      tc_wait_st();
      tc_fence_before();
      // one arrive per warp: 128 lanes arriving on one mbarrier word serialise in the shared-memory pipe (ncu counted
      // ~1e9 LSU bank conflicts per E-step launch, the same pipe the bulk copies and the mask ring use)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      step(cp);
      job += NGRP;
      ++own;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == LOAD_WARP) {
    // ===================== digit-plane loader =====================
    // one thread streams the pre-swizzled digit-plane tiles with cp.async.bulk as soon as a stage is free,
    // independently of the mask expansion (completion counted in bytes on the stage's "full" mbarrier)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int z = tile / (mtiles * qtiles), rem = tile % (mtiles * qtiles);
        const int qt = rem / mtiles;
        const int ks_begin = z * ks_per, ks_end = min(a.ksteps, ks_begin + ks_per);
        for (int ks = ks_begin; ks < ks_end; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sB = smem_base + stage * STAGE_BYTES;
          const int8_t *src = a.Bq + ((int64_t)ks * qtiles + qt) * (int64_t)B_BYTES;
          const uint32_t nbytes = B_BYTES;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(stage)), "r"(nbytes)
                       : "memory");
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB),
              "l"(src), "r"(nbytes), "r"(full_bar(stage))
              : "memory");
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    int stage = 0, buf = 0;
    uint32_t phase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qtiles);
      const int ks_begin = z * ks_per, ks_end = min(a.ksteps, ks_begin + ks_per);
      mbar_wait_warp(tempty_bar(buf), tphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_c = tmem_base + (uint32_t)(buf * TCOLS);
      if (ks_begin >= ks_end) {  // empty K slab (never produced by the host-side split): publish immediately
        if (lane == 0) tc_commit(tfull_bar(buf));
        __syncwarp();
      }
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        mbar_wait_warp(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sB = smem_base + stage * STAGE_BYTES;
          const uint64_t bdesc = smem_desc_sw128(sB);
          const uint32_t tmem_a = tmem_base + (uint32_t)(ACOL + 32 * stage);
#pragma unroll
          for (int k = 0; k < BKB / 32; ++k)
            tc_mma_i8_ts(tmem_c, tmem_a + (uint32_t)(8 * k), bdesc + (uint64_t)(2 * k), IDESC,
                         (ks > ks_begin || k > 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (ks == ks_end - 1) tc_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++buf == 2) {
        buf = 0;
        tphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue =====================
    int buf = 0;
    uint32_t tphase = 0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access (hardware rule: warp id mod 4)
    const int cset = (warp - EPI_WARP0) >> 2;  // which NQ / NEPI column set of the tile this warp drains
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qtiles), rem = tile % (mtiles * qtiles);
      const int qt = rem / mtiles, mt = rem % mtiles;
      const bool empty_slab = z * ks_per >= min(a.ksteps, z * ks_per + ks_per);
      double *out;
      int64_t ldo;
      bool accumulate;
      if (a.splitk > 1) {
        out = a.partials + (int64_t)z * a.M * a.Nq;
        ldo = a.Nq;
        accumulate = a.defer_reduce != 0;
      } else {
        out = a.Out;
        ldo = a.ldo;
        accumulate = a.accumulate != 0;
      }
      mbar_wait_warp(tfull_bar(buf), tphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < MH; ++h) {
        const int row = mt * BM + 128 * h + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TCOLS + h * N);
#pragma unroll 1
        for (int qc = cset * (NQ / 8 / NEPI); qc < (cset + 1) * (NQ / 8 / NEPI); ++qc) {
          int rg[T][8];
#pragma unroll
          for (int t = 0; t < T; ++t) tc_ld8(taddr + (uint32_t)(t * NQ + 8 * qc), rg[t]);
          tc_wait_ld();
          const int q = qt * NQ + 8 * qc;
          if (row < a.M && q < a.Nq) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int d[T];
#pragma unroll
              for (int t = 0; t < T; ++t) d[t] = rg[t][j];
              v[j] = empty_slab ? 0.0 : combine_planes<T>(d) * (a.scale[q + j] * combine_scale<T>());
            }
            // 32-byte accesses: a lane's 8 outputs are two full L2 sectors (16-byte stores wrote every sector twice)
            double *p = out + (int64_t)row * ldo + q;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (accumulate) {
                double o0, o1, o2, o3;
                ld_global_256(p + 4 * j, o0, o1, o2, o3);
                v[4 * j] += o0;
                v[4 * j + 1] += o1;
                v[4 * j + 2] += o2;
                v[4 * j + 3] += o3;
              }
              st_global_256(p + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (++buf == 2) {
        buf = 0;
        tphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// Two 32-column output tiles per expanded mask stage ("Q2", the default for long K loops; PPCA_B200_TC_Q2=0 disables).  The fast-path experiment showed
// the main loop is bound by the mask expansion / stage round trip, not by the tensor pipe (a third less MMA work
// changed nothing), so here one A stage in tensor memory feeds TWO N = 32 T MMA groups: tile = 128 rows x 64 columns,
// accumulators 2 x 32 T columns single-buffered (384 + 4 x 32 A columns = the 512 TMEM columns), the two digit-plane
// tiles of a K step are adjacent in memory and arrive with one bulk copy.  Same roles and barriers as above.
template <int T, int NGRP, int NEPI>
__global__ void __launch_bounds__(tb::THREADS_ATM, 1) tbitgemm_atm2_kernel(TBitGemmArgs a) {
  static_assert(32 * (4 * NGRP + 2 + 4 * NEPI) == tb::THREADS_ATM, "role split must add up to the CTA size");
  using namespace tb;
  constexpr int N = T * NQ;
  // the A tile never touches shared memory here: producers write it to tensor memory (tcgen05.st), the MMA
  // reads it from there, so shared memory only carries the digit-plane tile (write once, read once per K step)
  constexpr int NQT = 2;  // 32-column output tiles per tile (sharing every A stage)
  constexpr uint32_t B_BYTES = N * BKB, STAGE_BYTES = NQT * B_BYTES;
  constexpr int ACOL = NQT * MH * N;  // first TMEM column of the A stages (32 columns = 128 K-bytes each)
  static_assert(ACOL + 32 * STAGES <= 512, "A stages + the two accumulator tiles must fit the 512 TMEM columns");
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(MH == 1 && (STAGES & (STAGES - 1)) == 0, "producer groups assume one 128-row half and 2^n stages");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  // warps 0-3 / 4-7: two producer groups (alternate jobs; warp w writes TMEM lane quarter w & 3), warp 8: MMA issuer,
  // warps 9-12: epilogue (quarters 1,2,3,0), warp 13: digit-plane loader
  constexpr int MMA_WARP = 4 * NGRP, EPI_WARP0 = MMA_WARP + 1, LOAD_WARP = EPI_WARP0 + 4 * NEPI;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * STAGES + 2 + b); };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), PRODUCERS / 32 + 1);  // one elected lane per producer warp + the bulk-copy issue
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128 * NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int mtiles = (a.M + BM - 1) / BM, qtiles = (a.Nq + NQ - 1) / NQ;
  const int qgroups = (qtiles + NQT - 1) / NQT;  // pairs of q tiles
  const int ntiles = mtiles * qgroups * a.splitk;
  const int ks_per = (a.ksteps + a.splitk - 1) / a.splitk;

  if (warp < 4 * NGRP) {
    // ===================== producers =====================
    // The (tile, K step) jobs of this CTA form one sequence; producer group g = warp / 4 takes jobs j = g (mod NGRP).
    // Per job every thread of the group expands its mask row (row = 32 (warp & 3) + lane = TMEM lane) to int8
    // {0,1} in registers and writes the 128 bytes to the stage's tensor-memory columns with one tcgen05.st;
    // tcgen05.wait::st, fence, arrive.  The 16 bytes of mask a thread needs per job are fetched DEPTH own jobs
    // ahead with cp.async into a private shared-memory ring (mask rows are strided: latency, not bandwidth).
    constexpr int DEPTH = ATM_DEPTH, RING = ATM_RING;
    const int grp = warp >> 2, gt = tid & 127;  // group, thread within the group = row of the tile
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES) +
                     grp * (RING * 128 * 4);
    struct Cursor {
      int tiles_left, tile, ks, ks_end;
      const uint32_t *wrow;
      bool row_ok;
    };
    auto open_tile = [&](Cursor &c) {
      const int z = c.tile / (mtiles * qgroups), rem = c.tile % (mtiles * qgroups);
      const int mt = rem % mtiles;
      c.ks = z * ks_per;
      c.ks_end = min(a.ksteps, c.ks + ks_per);
      const int row = mt * BM + gt;
      c.row_ok = row < a.M;
      c.wrow = a.bits + (int64_t)(c.row_ok ? row : 0) * a.ldbits;
    };
    auto settle = [&](Cursor &c) {
      while (c.tiles_left > 0 && c.ks >= c.ks_end) {
        c.tile += gridDim.x;
        if (--c.tiles_left > 0) open_tile(c);
      }
    };
    auto step1 = [&](Cursor &c) {  // to the next job of the sequence
      if (c.tiles_left > 0) {
        ++c.ks;
        settle(c);
      }
    };
    auto step = [&](Cursor &c) {  // to this group's next job
#pragma unroll
      for (int g = 0; g < NGRP; ++g) step1(c);
    };
    auto start = [&](Cursor &c) {
      c.tiles_left = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      c.tile = blockIdx.x;
      c.ks = c.ks_end = 0;
      c.wrow = a.bits;
      c.row_ok = false;
      if (c.tiles_left > 0) open_tile(c);
      settle(c);
      if (grp == 1) step1(c);  // group 1 starts at job 1
    };
    auto fetch_words = [&](const Cursor &c, int slot) {  // 4 x 4-byte cp.async (mask rows are only 4-byte aligned)
      const uint32_t dst = smem_u32(ring + (slot * 128 + gt) * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int wi = 4 * c.ks + j;
        const bool ok = c.tiles_left > 0 && c.row_ok && wi < a.nwords;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * j), "l"(ok ? c.wrow + wi : a.bits),
                     "r"(ok ? 4 : 0));
      }
      asm volatile("cp.async.commit_group;" ::);
    };
    Cursor cp, cf;  // processing cursor, fetch cursor (DEPTH own jobs ahead)
    start(cp);
    start(cf);
    int nf = 0;
    for (; nf < DEPTH; ++nf) {
      fetch_words(cf, nf % RING);
      step(cf);
    }
    int job = grp, own = 0;
    while (cp.tiles_left > 0) {
      fetch_words(cf, nf % RING);
      ++nf;
      step(cf);
      const int stage = job & (STAGES - 1);
      const uint32_t phase = (uint32_t)(job / STAGES) & 1u;
      mbar_wait_warp(empty_bar(stage), phase ^ 1u);
      tc_fence_after();
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");  // this job's own mask words have landed
      const uint4 wq = *reinterpret_cast<const uint4 *>(ring + ((own % RING) * 128 + gt) * 4);
      const uint32_t wc[4] = {wq.x, wq.y, wq.z, wq.w};
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = nib4(wc[j >> 3], 4 * (j & 7));
      tc_st32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(ACOL + 32 * stage), v);
      tc_wait_st();
      tc_fence_before();
      // one arrive per warp: 128 lanes arriving on one mbarrier word serialise in the shared-memory pipe (ncu counted
      // ~1e9 LSU bank conflicts per E-step launch, the same pipe the bulk copies and the mask ring use)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      step(cp);
      job += NGRP;
      ++own;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == LOAD_WARP) {
    // ===================== digit-plane loader =====================
    // one thread streams the pre-swizzled digit-plane tiles with cp.async.bulk as soon as a stage is free,
    // independently of the mask expansion (completion counted in bytes on the stage's "full" mbarrier)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int z = tile / (mtiles * qgroups), rem = tile % (mtiles * qgroups);
        const int qt = NQT * (rem / mtiles);
        const int ks_begin = z * ks_per, ks_end = min(a.ksteps, ks_begin + ks_per);
        for (int ks = ks_begin; ks < ks_end; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sB = smem_base + stage * STAGE_BYTES;
          const int8_t *src = a.Bq + ((int64_t)ks * qtiles + qt) * (int64_t)B_BYTES;
          const uint32_t nbytes = (qt + 1 < qtiles ? 2u : 1u) * B_BYTES;  // tiles (ks, qt) and (ks, qt + 1) are adjacent
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(stage)), "r"(nbytes)
                       : "memory");
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sB),
              "l"(src), "r"(nbytes), "r"(full_bar(stage))
              : "memory");
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    int stage = 0, buf = 0;
    uint32_t phase = 0, tphase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qgroups), rem = tile % (mtiles * qgroups);
      const int ngroups = (NQT * (rem / mtiles) + 1 < qtiles) ? 2 : 1;  // the last pair may hold one q tile only
      const int ks_begin = z * ks_per, ks_end = min(a.ksteps, ks_begin + ks_per);
      mbar_wait_warp(tempty_bar(buf), tphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_c = tmem_base;
      if (ks_begin >= ks_end) {  // empty K slab (never produced by the host-side split): publish immediately
        if (lane == 0) tc_commit(tfull_bar(buf));
        __syncwarp();
      }
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        mbar_wait_warp(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sB = smem_base + stage * STAGE_BYTES;
          const uint32_t tmem_a = tmem_base + (uint32_t)(ACOL + 32 * stage);
          for (int g = 0; g < ngroups; ++g) {
            const uint64_t bd = smem_desc_sw128(sB + g * B_BYTES);
#pragma unroll
            for (int k = 0; k < BKB / 32; ++k)
              tc_mma_i8_ts(tmem_c + (uint32_t)(g * N), tmem_a + (uint32_t)(8 * k), bd + (uint64_t)(2 * k), IDESC,
                           (ks > ks_begin || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));
          if (ks == ks_end - 1) tc_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      tphase ^= 1u;  // one accumulator buffer: every tile flips the phase
    }
  } else {
    // ===================== epilogue =====================
    int buf = 0;
    uint32_t tphase = 0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access (hardware rule: warp id mod 4)
    const int cset = (warp - EPI_WARP0) >> 2;  // which NQ / NEPI column set of the tile this warp drains
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / (mtiles * qgroups), rem = tile % (mtiles * qgroups);
      const int qt0 = NQT * (rem / mtiles), mt = rem % mtiles;
      const bool empty_slab = z * ks_per >= min(a.ksteps, z * ks_per + ks_per);
      double *out;
      int64_t ldo;
      bool accumulate;
      if (a.splitk > 1) {
        out = a.partials + (int64_t)z * a.M * a.Nq;
        ldo = a.Nq;
        accumulate = a.defer_reduce != 0;
      } else {
        out = a.Out;
        ldo = a.ldo;
        accumulate = a.accumulate != 0;
      }
      mbar_wait_warp(tfull_bar(buf), tphase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < NQT; ++h) {  // the two q tiles of this tile
        const int qt = qt0 + h;
        if (qt >= qtiles) break;
        const int row = mt * BM + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * N);
#pragma unroll 1
        for (int qc = cset * (NQ / 8 / NEPI); qc < (cset + 1) * (NQ / 8 / NEPI); ++qc) {
          int rg[T][8];
#pragma unroll
          for (int t = 0; t < T; ++t) tc_ld8(taddr + (uint32_t)(t * NQ + 8 * qc), rg[t]);
          tc_wait_ld();
          const int q = qt * NQ + 8 * qc;
          if (row < a.M && q < a.Nq) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int d[T];
#pragma unroll
              for (int t = 0; t < T; ++t) d[t] = rg[t][j];
              v[j] = empty_slab ? 0.0 : combine_planes<T>(d) * (a.scale[q + j] * combine_scale<T>());
            }
            // 32-byte accesses: a lane's 8 outputs are two full L2 sectors (16-byte stores wrote every sector twice)
            double *p = out + (int64_t)row * ldo + q;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (accumulate) {
                double o0, o1, o2, o3;
                ld_global_256(p + 4 * j, o0, o1, o2, o3);
                v[4 * j] += o0;
                v[4 * j + 1] += o1;
                v[4 * j + 2] += o2;
                v[4 * j + 3] += o3;
              }
              st_global_256(p + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      tphase ^= 1u;  // one accumulator buffer: every tile flips the phase
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}


// ---------------------------------------------------------------------------------------------
// digit planes: [kstep][qtile] tiles, each the exact SWIZZLE_128B shared-memory image of (T NQ rows) x 128 bytes
// ---------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(128) slice_tc_kernel(const double *__restrict__ B, int64_t ldb, int K, int Nq,
                                                       int kblocks32, const unsigned long long *__restrict__ cm,
                                                       int8_t *out, double *scale) {
  // block = 32 columns x 4 consecutive 32-row blocks (one 128-byte K step)
  const int q = blockIdx.x * 32 + (threadIdx.x & 31);
  const int qt = q / tb::NQ, qi = q % tb::NQ;
  const int ks = blockIdx.y, kb = ks * 4 + (threadIdx.x >> 5);
  if (q >= Nq || kb >= kblocks32) return;
  const int qtiles = (Nq + tb::NQ - 1) / tb::NQ;
  const double m = __longlong_as_double((long long)cm[q]);
  int ex = 0;
  if (m > 0.0) frexp(m, &ex);
  if (kb == 0) scale[q] = ldexp(1.0, ex);
  const double inv = ldexp(1.0, 8 * T - 2 - ex);
  uint32_t words[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) words[t][j] = 0u;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    double xs[16];  // 16 independent loads in flight per thread before the first conversion
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int row = kb * 32 + 16 * half + r;
      xs[r] = (row < K) ? __ldg(B + (int64_t)row * ldb + q) : 0.0;
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const int r = 16 * half + rr;
      // rint(x / s 2^(8T-2)) as T balanced base-256 digits (see ibitgemm.cu): +128 below the top digit, then ^ 0x80
      constexpr unsigned long long BIAS = 0x0080808080808080ull & ((1ull << (8 * (T - 1))) - 1ull);
      const unsigned long long v = (unsigned long long)(__double2ll_rn(xs[rr] * inv) + (long long)BIAS) ^ BIAS;
#pragma unroll
      for (int t = 0; t < T; ++t)  // plane 0 = most significant digit
        words[t][r >> 2] |= (uint32_t)((v >> (8 * (T - 1 - t))) & 0xffull) << (8 * (r & 3));
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    // shared-memory image of the SWIZZLE_128B K-major tile: row n = t NQ + qi, 16-byte chunks 2 (kb & 3) and + 1
    int8_t *tile = out + ((int64_t)ks * qtiles + qt) * (int64_t)(T * tb::NQ * 128);
    const int n = t * tb::NQ + qi, c0 = 2 * (kb & 3);
    const int base = (n >> 3) * 1024 + (n & 7) * 128;
    // chunks c0 and c0 + 1 (c0 even) swizzle to the two halves of ONE aligned 32-byte sector, swapped when n is odd:
    // one 256-bit store instead of two half-sector stores
    const int sw = n & 7, lo = ((c0 ^ sw) & ~1) << 4;
    const bool swap = sw & 1;
    const uint32_t a0 = swap ? words[t][4] : words[t][0], a1 = swap ? words[t][5] : words[t][1];
    const uint32_t a2 = swap ? words[t][6] : words[t][2], a3 = swap ? words[t][7] : words[t][3];
    const uint32_t b0 = swap ? words[t][0] : words[t][4], b1 = swap ? words[t][1] : words[t][5];
    const uint32_t b2 = swap ? words[t][2] : words[t][6], b3 = swap ? words[t][3] : words[t][7];
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(tile + base + lo), "r"(a0), "r"(a1), "r"(a2),
                 "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3)
                 : "memory");
  }
}

// column maxima of |B|: 32 columns x 8 row lanes per block, coalesced along columns, 4 independent loads in flight
__global__ void __launch_bounds__(256) colmax_kernel_tc(const double *__restrict__ B, int64_t ldb, int K, int Nq,
                                                        unsigned long long *cm) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int q = blockIdx.x * 32 + tx;
  const int per = (K + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(K, lo + per);
  double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
  if (q < Nq) {
    int r = lo + ty;
    for (; r + 24 < hi; r += 32) {
      m0 = fmax(m0, fabs(B[(int64_t)r * ldb + q]));
      m1 = fmax(m1, fabs(B[(int64_t)(r + 8) * ldb + q]));
      m2 = fmax(m2, fabs(B[(int64_t)(r + 16) * ldb + q]));
      m3 = fmax(m3, fabs(B[(int64_t)(r + 24) * ldb + q]));
    }
    for (; r < hi; r += 8) m0 = fmax(m0, fabs(B[(int64_t)r * ldb + q]));
  }
  red[ty][tx] = fmax(fmax(m0, m1), fmax(m2, m3));
  __syncthreads();
  if (ty == 0 && q < Nq) {
    double m = red[0][tx];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmax(m, red[j][tx]);
    if (m > 0.0) atomicMax(cm + q, (unsigned long long)__double_as_longlong(m));  // order-independent
  }
}

size_t sliced_tc_bytes(int kblocks32, int Nq, int T) {
  const int64_t ks = (kblocks32 + 3) / 4, qt = (Nq + tb::NQ - 1) / tb::NQ;
  return (size_t)(ks * qt * T * tb::NQ * 128);
}

void launch_slice_tc(const Launcher &L, const double *Bmat, int64_t ldb, int K, int Nq, int kblocks32, int T, int8_t *q,
                     double *scale, unsigned long long *cm, bool have_colmax) {
  if (Nq <= 0 || kblocks32 <= 0) return;
  const int qtiles = (Nq + 31) / 32;  // 32-column thread blocks (independent of the MMA tile width)
  if (!have_colmax) {
    CUDA_CHECK(cudaMemsetAsync(cm, 0, sizeof(unsigned long long) * Nq, L.stream));
    int slabs = (4 * L.sms + qtiles - 1) / qtiles;
    if (slabs > (K + 63) / 64) slabs = (K + 63) / 64;
    if (slabs < 1) slabs = 1;
    colmax_kernel_tc<<<dim3(qtiles, slabs), 256, 0, L.stream>>>(Bmat, ldb, K, Nq, cm);
    CUDA_CHECK(cudaGetLastError());
    ++*L.launch_counter;
  }
  dim3 grid(qtiles, (kblocks32 + 3) / 4);
  if (T == 4) slice_tc_kernel<4><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, kblocks32, cm, q, scale);
  else if (T == 6) slice_tc_kernel<6><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, kblocks32, cm, q, scale);
  else if (T == 7) slice_tc_kernel<7><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, kblocks32, cm, q, scale);
  else if (T == 8) slice_tc_kernel<8><<<grid, 128, 0, L.stream>>>(Bmat, ldb, K, Nq, kblocks32, cm, q, scale);
  else PPCA_THROW(PPCA_ERR_INVALID, "tcgen05 int8 path: T must be 4, 6, 7 or 8 (got %d)", T);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

int tbitgemm_pick_splitk(int M, int Nq, int ksteps, int sms) {
  const int64_t tiles = round_up(M, tb::BM) / tb::BM * (round_up(Nq, tb::NQ) / tb::NQ);
  if (tiles >= 4 * sms) return 1;  // persistent kernel: enough tiles to balance
  // fewest K slabs (>= 4 K steps = 512 rows each) that keep every persistent CTA busy in the last round
  const int64_t max_s = ksteps / 4 > 0 ? ksteps / 4 : 1;
  int best = 1;
  double best_eff = 0.0;
  for (int64_t s = 1; s <= max_s; ++s) {
    const int64_t per = (ksteps + s - 1) / s;
    if ((s - 1) * per >= ksteps) continue;  // would leave an empty slab
    const int64_t ctas = tiles * s;
    if (ctas < sms && s < max_s) continue;
    const double eff = (double)ctas / (double)round_up(ctas, sms);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = (int)s;
    }
    if (eff >= 0.97) break;
  }
  return best;
}

template <int T>
static void launch_tb(const Launcher &L, const TBitGemmArgs &a) {
  const int64_t tiles = round_up(a.M, tb::BM) / tb::BM * (round_up(a.Nq, tb::NQ) / tb::NQ) * a.splitk;
  const int grid = (int)(tiles < L.sms ? tiles : L.sms);
  if constexpr (2 * T * tb::NQ + 32 * tb::STAGES <= 512) {
    static const bool smem_a = getenv("PPCA_B200_TC_SMEM_A") && atoi(getenv("PPCA_B200_TC_SMEM_A")) == 1;
    if (!smem_a) {  // A operand in tensor memory: shared memory carries only the digit planes
      constexpr size_t SMEM_ATM = (size_t)tb::STAGES * (T * tb::NQ * tb::BKB) + 1024 + 2 * tb::ATM_RING * 128 * 16;
      static PerDeviceOnce configured_atm;
      if (configured_atm.need()) {
        CUDA_CHECK(cudaFuncSetAttribute(tbitgemm_atm_kernel<T, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ATM));
        CUDA_CHECK(cudaFuncSetAttribute(tbitgemm_atm_kernel<T, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ATM));
      }
      static const int force = getenv("PPCA_B200_TC_ROLES") ? atoi(getenv("PPCA_B200_TC_ROLES")) : 0;
      const int ks_per = (a.ksteps + a.splitk - 1) / a.splitk;
      const bool drain = force ? force == 2 : ks_per <= 4;  // short K loops: epilogue-bound
      if constexpr (2 * T * tb::NQ + 32 * tb::STAGES <= 512) {
        // long K loops: two 32-column output tiles per expanded mask stage (tbitgemm_atm2_kernel; measured -13 % / -15 %
        // on the c3 E-/M-step contraction); PPCA_B200_TC_Q2=0 keeps the one-tile kernel
        static const bool q2 = !(getenv("PPCA_B200_TC_Q2") && atoi(getenv("PPCA_B200_TC_Q2")) == 0);
        // ... when the halved tile count still fills the persistent grid (c2's M-step has 6 x 29 = 174 paired tiles
        // for 148 CTAs: 59 % of the second round would idle, measured 0.35 vs 0.24 ms)
        const int64_t qg_ = (round_up(a.Nq, tb::NQ) / tb::NQ + 1) / 2;
        const int64_t t2_ = round_up(a.M, tb::BM) / tb::BM * qg_ * a.splitk;
        const bool balanced = (double)t2_ >= 0.85 * (double)(round_up(t2_, L.sms));
        if (q2 && !drain && a.Nq > tb::NQ && balanced) {
          constexpr size_t SMEM_Q2 = (size_t)tb::STAGES * 2 * (T * tb::NQ * tb::BKB) + 1024 + 2 * tb::ATM_RING * 128 * 16;
          static PerDeviceOnce configured_q2;
          if (configured_q2.need()) {
            CUDA_CHECK(cudaFuncSetAttribute(tbitgemm_atm2_kernel<T, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)SMEM_Q2));
          }
          const int64_t qg = (round_up(a.Nq, tb::NQ) / tb::NQ + 1) / 2;
          const int64_t tiles2 = round_up(a.M, tb::BM) / tb::BM * qg * a.splitk;
          const int grid2 = (int)(tiles2 < L.sms ? tiles2 : L.sms);
          tbitgemm_atm2_kernel<T, 2, 1><<<grid2, tb::THREADS_ATM, SMEM_Q2, L.stream>>>(a);
          CUDA_CHECK(cudaGetLastError());
          ++*L.launch_counter;
          L.count(V_TC_ATM2);
          return;
        }
      }
      if (drain) tbitgemm_atm_kernel<T, 1, 2><<<grid, tb::THREADS_ATM, SMEM_ATM, L.stream>>>(a);
      else tbitgemm_atm_kernel<T, 2, 1><<<grid, tb::THREADS_ATM, SMEM_ATM, L.stream>>>(a);
      CUDA_CHECK(cudaGetLastError());
      ++*L.launch_counter;
      L.count(drain ? V_TC_ATM_DRAIN : V_TC_ATM_FEED);
      return;
    }
  }
  constexpr size_t SMEM = (size_t)tb::STAGES * (tb::BM * tb::BKB + T * tb::NQ * tb::BKB) + 1024;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(tbitgemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  tbitgemm_kernel<T><<<grid, tb::THREADS, SMEM, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(V_TC_SMEM_A);
}

void launch_tbitgemm(const Launcher &L, const uint32_t *bits, int64_t ldbits, int nwords, const int8_t *Bq,
                     const double *scale, int T, double *Out, int64_t ldo, int M, int Nq, int ksteps, int accumulate,
                     double *partials, int splitk, int defer_reduce) {
  REQUIRE(Nq % 8 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(Out) & 31) == 0 &&
              (partials == nullptr || (reinterpret_cast<uintptr_t>(partials) & 31) == 0),
          "tbitgemm: Nq must be a multiple of 8, output pitch a multiple of 4, outputs 32-byte aligned");
  REQUIRE(splitk >= 1 && splitk <= (ksteps > 0 ? ksteps : 1), "tbitgemm: bad split-K");
  {  // every K slab must be non-empty
    const int per = (ksteps + splitk - 1) / splitk;
    REQUIRE((splitk - 1) * per < ksteps || splitk == 1, "tbitgemm: split-K leaves an empty slab");
  }
  REQUIRE(splitk == 1 || partials != nullptr, "tbitgemm: split-K needs a partials workspace");
  if (M <= 0 || Nq <= 0 || ksteps <= 0) return;
  TBitGemmArgs a;
  a.bits = bits; a.ldbits = ldbits; a.nwords = nwords; a.Bq = Bq; a.scale = scale; a.Out = Out; a.ldo = ldo;
  a.M = M; a.Nq = Nq; a.ksteps = ksteps; a.accumulate = accumulate; a.partials = partials; a.splitk = splitk;
  a.defer_reduce = defer_reduce;
  if (T == 4) launch_tb<4>(L, a);
  else if (T == 6) launch_tb<6>(L, a);
  else if (T == 7) launch_tb<7>(L, a);
  else if (T == 8) launch_tb<8>(L, a);
  else PPCA_THROW(PPCA_ERR_INVALID, "int8 path: T must be 6, 7 or 8 (got %d)", T);
  if (splitk > 1 && !defer_reduce) launch_bitgemm_reduce(L, partials, splitk, M, Nq, Out, ldo, accumulate);
}

}  // namespace ppca
