// bitgemm.cu — the masked-Gram contraction  Out[M x Nq] (+)= Bits[M x K] * Bmat[K x Nq]  in FP64.
//
// Both heavy contractions of the EM iteration are this one kernel:
//   E-step  Gs = Mask  * Ksym   (rows = samples,    K = output dims)  <- G_n = C_o^T C_o
//           reference: output_covariance.rs:57-59 inner_product after :123-131 masked, per sample
//   M-step  A += Mask^T * W     (rows = output dims, K = samples)     <- S_i = sum_{n: m_ni} w (z z^T + Sigma_n)
//           reference: ppca_model.rs:297-306, the dimension-parallel / sample-serial loop
// The left operand is the bit-packed mask (1 bit per element in HBM); each thread expands its bits to
// 0.0 / 1.0 in registers while building DMMA A-fragments, so the operand costs no shared-memory or HBM
// traffic for doubles.  The right operand streams through a cp.async multi-stage shared-memory pipeline
// (row pitch = 4 mod 16 doubles -> conflict-free B-fragment reads).  Math is mma.sync.m8n8k4.f64
// (DMMA.8x8x4); tcgen05 has no FP64 kind.
#include <cstdio>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

template <int MI_, int NI_, int WM_, int WN_, int STAGES_>
struct BgCfg {
  static constexpr int MI = MI_, NI = NI_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int BM = 8 * MI * WM;
  static constexpr int BN = 8 * NI * WN;
  static constexpr int THREADS = 32 * WM * WN;
  static constexpr int LDB = BN + ((20 - BN % 16) % 16);  // LDB % 16 == 4
  static constexpr int STAGE_DOUBLES = 32 * LDB;
  static constexpr int CH = BN / 2;  // 16-byte chunks per tile row
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) bitgemm_kernel(BitGemmArgs a) {
  constexpr int MI = Cfg::MI, NI = Cfg::NI, WM = Cfg::WM, STAGES = Cfg::STAGES;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, LDB = Cfg::LDB, THREADS = Cfg::THREADS, CH = Cfg::CH;
  extern __shared__ __align__(16) double smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int r = lane >> 2, c = lane & 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  const int kb_per = (a.kblocks + a.splitk - 1) / a.splitk;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(a.kblocks, kb_begin + kb_per);
  const int nkb = max(0, kb_end - kb_begin);

  auto load_stage = [&](int stage, int kb) {
    double *sB = smem + stage * Cfg::STAGE_DOUBLES;
    const double *g = a.Bmat + (int64_t)kb * 32 * a.ldb + n0;
#pragma unroll 4
    for (int idx = tid; idx < 32 * CH; idx += THREADS) {
      const int row = idx / CH, col = (idx % CH) * 2;
      const bool ok = (n0 + col) < a.Nq;
      const double *src = ok ? g + (int64_t)row * a.ldb + col : a.Bmat;
      cp_async16(sB + row * LDB + col, src, ok ? 16 : 0);
    }
  };

  // this thread's bit rows
  const uint32_t *brow[MI];
  bool brow_ok[MI];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int row = m0 + wm * (8 * MI) + 8 * mi + r;
    brow_ok[mi] = row < a.M;
    brow[mi] = a.bits + (int64_t)(brow_ok[mi] ? row : 0) * a.ldbits;
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, kb_begin + s);
    cp_async_commit();
  }
  uint32_t wcur[MI], wnext[MI];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    wcur[mi] = (nkb > 0 && brow_ok[mi]) ? __ldg(brow[mi] + kb_begin) : 0u;
    wnext[mi] = 0u;
  }

  for (int it = 0; it < nkb; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < nkb) load_stage(nxt % STAGES, kb_begin + nxt);
      cp_async_commit();
    }
    if (it + 1 < nkb) {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) wnext[mi] = brow_ok[mi] ? __ldg(brow[mi] + kb_begin + it + 1) : 0u;
    }
    const double *sB = smem + (it % STAGES) * Cfg::STAGE_DOUBLES + c * LDB + wn * (8 * NI) + r;
    uint32_t wsh[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) wsh[mi] = wcur[mi] >> c;
    const int smax = min(8, (a.kcols - 32 * (kb_begin + it) + 3) >> 2);  // trims the ragged last K block
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s >= smax) break;
      double b[NI];
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) b[ni] = sB[(4 * s) * LDB + 8 * ni];
      double av[MI];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) av[mi] = bit_to_double(wsh[mi], 4 * s);
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], b[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) wcur[mi] = wnext[mi];
  }
  cp_async_wait<0>();

  // epilogue
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
  for (int mi = 0; mi < MI; ++mi) {
    const int row = m0 + wm * (8 * MI) + 8 * mi + r;
    if (row >= a.M) continue;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      const int col = n0 + wn * (8 * NI) + 8 * ni + 2 * c;
      if (col >= a.Nq) continue;
      double2 *p = reinterpret_cast<double2 *>(out + (int64_t)row * ldo + col);
      double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      if (accumulate) {
        const double2 o = *p;
        v.x += o.x;
        v.y += o.y;
      }
      *p = v;
    }
  }
}

// Out[m][n] (+)= sum_z partials[z][m][n], z in fixed ascending order (run-to-run reproducible).
__global__ void bitgemm_reduce_kernel(const double *__restrict__ partials, int splitk, int M, int Nq, double *Out,
                                      int64_t ldo, int accumulate) {
  const int64_t total = (int64_t)M * (Nq / 2);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / (Nq / 2)), col = (int)(idx % (Nq / 2)) * 2;
    double2 s = make_double2(0.0, 0.0);
    const double2 *p = reinterpret_cast<const double2 *>(partials + (int64_t)row * Nq + col);
    for (int z = 0; z < splitk; ++z) {
      const double2 v = p[(int64_t)z * M * (Nq / 2)];
      s.x += v.x;
      s.y += v.y;
    }
    double2 *o = reinterpret_cast<double2 *>(Out + (int64_t)row * ldo + col);
    if (accumulate) {
      const double2 v = *o;
      s.x += v.x;
      s.y += v.y;
    }
    *o = s;
  }
}

// ---- configuration table ---------------------------------------------------------------------
using Cfg128 = BgCfg<4, 8, 4, 2, 4>;   // 128 x 128
using Cfg136 = BgCfg<2, 17, 8, 1, 4>;  // 128 x 136  (k = 16 : kk = 136)
using Cfg104 = BgCfg<2, 13, 8, 1, 4>;  // 128 x 104  (k = 64 : kk = 2080 = 20 * 104)
using Cfg64 = BgCfg<4, 8, 8, 1, 4>;    // 256 x 64
using Cfg32 = BgCfg<4, 4, 8, 1, 4>;    // 256 x 32
using Cfg16 = BgCfg<4, 2, 8, 1, 4>;    // 256 x 16
using Cfg8 = BgCfg<4, 1, 8, 1, 4>;     // 256 x 8

struct CfgInfo {
  int BM, BN;
};
static const CfgInfo kCfgs[] = {{Cfg128::BM, Cfg128::BN}, {Cfg136::BM, Cfg136::BN}, {Cfg104::BM, Cfg104::BN},
                                {Cfg64::BM, Cfg64::BN},   {Cfg32::BM, Cfg32::BN},   {Cfg16::BM, Cfg16::BN},
                                {Cfg8::BM, Cfg8::BN}};
static const int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);

// smallest padded column count wins; ties go to the wider tile
static int pick_cfg(int Nq) {
  int best = 0;
  int64_t best_cost = -1;
  for (int i = 0; i < kNumCfgs; ++i) {
    const int64_t cost = round_up(Nq, kCfgs[i].BN);
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && kCfgs[i].BN > kCfgs[best].BN)) {
      best = i;
      best_cost = cost;
    }
  }
  return best;
}

int bitgemm_pick_splitk(int M, int Nq, int kblocks, int sms) {
  const CfgInfo &c = kCfgs[pick_cfg(Nq)];
  const int64_t tiles = round_up(M, c.BM) / c.BM * (round_up(Nq, c.BN) / c.BN);
  const int64_t max_s = kblocks / 8 > 0 ? kblocks / 8 : 1;  // at least 8 K-blocks (256 rows) per slab
  if (tiles >= sms) {
    // one CTA per SM: pick the smallest split that keeps the last wave >= 95% full
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

size_t bitgemm_partials_len(int M, int Nq, int splitk) { return splitk > 1 ? (size_t)splitk * M * Nq : 0; }

template <class Cfg>
static void launch_cfg(const Launcher &L, const BitGemmArgs &a) {
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(bitgemm_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  dim3 grid((unsigned)(round_up(a.M, Cfg::BM) / Cfg::BM), (unsigned)(round_up(a.Nq, Cfg::BN) / Cfg::BN),
            (unsigned)a.splitk);
  bitgemm_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(V_DMMA);
}

void launch_bitgemm(const Launcher &L, const BitGemmArgs &a) {
  REQUIRE(a.Nq % 8 == 0 && a.ldb % 2 == 0 && a.ldo % 2 == 0, "bitgemm: Nq must be a multiple of 8, pitches even");
  REQUIRE(a.splitk >= 1, "bitgemm: splitk >= 1");
  REQUIRE(a.splitk == 1 || a.partials != nullptr, "bitgemm: split-K needs a partials workspace");
  if (a.M <= 0 || a.Nq <= 0) return;
  switch (pick_cfg(a.Nq)) {
    case 0: launch_cfg<Cfg128>(L, a); break;
    case 1: launch_cfg<Cfg136>(L, a); break;
    case 2: launch_cfg<Cfg104>(L, a); break;
    case 3: launch_cfg<Cfg64>(L, a); break;
    case 4: launch_cfg<Cfg32>(L, a); break;
    case 5: launch_cfg<Cfg16>(L, a); break;
    default: launch_cfg<Cfg8>(L, a); break;
  }
  if (a.splitk > 1 && !a.defer_reduce)
    launch_bitgemm_reduce(L, a.partials, a.splitk, a.M, a.Nq, a.Out, a.ldo, a.accumulate);
}

void launch_bitgemm_reduce(const Launcher &L, const double *partials, int splitk, int M, int Nq, double *Out,
                           int64_t ldo, int accumulate) {
  if (splitk <= 1 || M <= 0 || Nq <= 0) return;
  const int64_t total = (int64_t)M * (Nq / 2);
  const int blocks = (int)((total + 255) / 256 < 4 * L.sms ? (total + 255) / 256 : 4 * L.sms);
  bitgemm_reduce_kernel<<<blocks, 256, 0, L.stream>>>(partials, splitk, M, Nq, Out, ldo, accumulate);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
