// moments.cu — the sample-matrix side of the M-step and the reconstruction writers.
//
// cross_moment_kernel (the one pass over X of the M-step):
//   B      += Xc^T (w .* Z)                  total_cross_moment            (ppca_model.rs:281-293)
//   Sx_i   += sum_n w_n Xc_ni                (total_deviation = Sx - rowdot(C, Mask^T (w Z)), :338-347)
//   tot_i  += sum_n w_n m_ni                 totals                        (:348)
// The cross moment runs on DMMA from one centred shared-memory tile of X, (M,N,K) = (dims, k, samples); the column
// sums are plain FMAs folded into the centring loop.  Empty samples have all bits clear and contribute nothing,
// matching the reference's filter (:333).
//
// reconstruct_kernel: smoothed = C z + mu (ppca_model.rs:454-456); extrapolated = mask.choose(x, smoothed)
// (:460-463, utils.rs:137-153) — observed slots are copied, never recomputed.
#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(src_bytes));
}

template <int KT, int BS>
struct CrCfg {
  static constexpr int BD = 64, LDX = 68;
  static constexpr int KPP = 8 * KT;
  static constexpr int LDZ = KPP + ((20 - KPP % 16) % 16);
  static constexpr int STAGE = BS * LDX + BS * LDZ + BS + BS;  // X, WZ, w, mask words (2 u32 per row)
  static constexpr int FIXED = BD;                             // mu
  static constexpr size_t SMEM = (size_t)(FIXED + 2 * STAGE) * sizeof(double);
};

struct CrArgs {
  const double *X;
  int ldx;
  const uint32_t *mask;
  int dw;
  int64_t row0;
  int rows;
  int kp;
  int d32;
  const double *mupad;
  const double *WZ;  // chunk-local, row pitch kp
  const double *w;   // chunk-local weights
  int d64;           // dblocks * 64
  double *pB, *pT, *pO;  // per-slab partial slots, accumulated (+=) across chunks
};

// One pass over X per EM iteration (the ONLY one of the M-step):
//   B_i   += sum_n x~_ni (w z)_n                 total_cross_moment              (ppca_model.rs:281-293)
//   Sx_i  += sum_n w_n x~_ni                     first half of total_deviation   (:338-347)
//   tot_i += sum_n w_n m_ni                      totals                          (:348)
// with x~ = m ? x - mu : 0 (OLD mu; select, never multiply).  The residual R = m (x~ - C z) of the reference is never
// formed: sum_n w R_ni = Sx_i - c_i . (Mask^T (w Z))_i, whose second term rides on the M-step contraction, and |R_n|^2 is a
// per-sample scalar of the solve kernel (see SolveArgs::tn).  Round 1 ran a second DMMA GEMM (Z C^T) in this kernel for them.
// resident CTAs per SM: the kernel is a stream over X with two tiles in flight per CTA, so small states (few accumulator
// registers, small tiles) run more CTAs to keep more loads in flight
__host__ __device__ constexpr int cr_minb(int kt) { return kt <= 2 ? 4 : (kt <= 4 ? 3 : (kt <= 8 ? 2 : 1)); }

template <int KT, int BS>
__global__ void __launch_bounds__(256, cr_minb(KT)) cross_moment_kernel(CrArgs a) {
  using Cfg = CrCfg<KT, BS>;
  constexpr int LDX = Cfg::LDX, LDZ = Cfg::LDZ, KPP = Cfg::KPP;
  extern __shared__ __align__(16) double smem[];
  double *sMu = smem;
  double *stage0 = sMu + 64;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int xb = blockIdx.x;  // dimension block
  const int ntiles = (a.rows + BS - 1) / BS;

  auto load_tile = [&](int st, int t) {
    double *sX = stage0 + st * Cfg::STAGE;
    double *sWZ = sX + BS * LDX;
    double *sW = sWZ + BS * LDZ;
    uint32_t *sM = reinterpret_cast<uint32_t *>(sW + BS);
    const int base = t * BS;
    for (int idx = tid; idx < BS * 32; idx += 256) {
      const int row = idx >> 5, col = (idx & 31) * 2;
      const int gcol = 64 * xb + col;
      const bool ok = gcol < a.ldx;
      const double *src = ok ? a.X + (a.row0 + base + row) * a.ldx + gcol : a.X;
      cp_async16(sX + row * LDX + col, src, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BS * (KPP / 2); idx += 256) {
      const int row = idx / (KPP / 2), col = (idx % (KPP / 2)) * 2;
      const bool ok = col < a.kp;
      const int64_t off = (int64_t)(base + row) * a.kp + col;
      cp_async16(sWZ + row * LDZ + col, ok ? a.WZ + off : a.WZ, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BS; idx += 256) cp_async8(sW + idx, a.w + base + idx, 8);
    for (int idx = tid; idx < BS * 2; idx += 256) {
      const int row = idx >> 1, j = idx & 1;
      const int wj = 2 * xb + j;
      const bool ok = wj < a.dw;
      cp_async4(sM + idx, ok ? a.mask + (a.row0 + base + row) * a.dw + wj : a.mask, ok ? 4 : 0);
    }
  };

  if (tid < 64) sMu[tid] = (64 * xb + tid < a.d32) ? a.mupad[64 * xb + tid] : 0.0;

  double accB[2][KT][2];  // two independent accumulator sets (even / odd K steps) to shorten the DMMA chains
#pragma unroll
  for (int ni = 0; ni < KT; ++ni) accB[0][ni][0] = accB[0][ni][1] = accB[1][ni][0] = accB[1][ni][1] = 0.0;
  double sx = 0.0, tot = 0.0;  // this thread's column (tid & 63), rows (tid >> 6) + 4 it of every tile

  int t = blockIdx.y;
  if (t < ntiles) load_tile(0, t);
  cp_async_commit();
  int st = 0;
  for (; t < ntiles; t += gridDim.y, st ^= 1) {
    const int tn = t + gridDim.y;
    if (tn < ntiles) load_tile(st ^ 1, tn);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    double *sX = stage0 + st * Cfg::STAGE;
    const double *sWZ = sX + BS * LDX;
    const double *sW = sWZ + BS * LDZ;
    const uint32_t *sM = reinterpret_cast<const uint32_t *>(sW + BS);

    // centre + select (utils.rs:118-127 fillna semantics: select, never multiply), weighted column sums on the way
    {
      const int col = tid & 63;
      const double mu = sMu[col];
#pragma unroll
      for (int it = 0; it < BS * 64 / 256; ++it) {
        const int row = (tid >> 6) + 4 * it;
        const bool obs = (sM[row * 2 + (col >> 5)] >> (col & 31)) & 1u;
        double *p = sX + row * LDX + col;
        const double xc = obs ? (*p - mu) : 0.0;
        *p = xc;
        const double wn = sW[row];  // rows past the chunk carry weight 0 and mask 0
        sx = fma(wn, xc, sx);
        tot += obs ? wn : 0.0;
      }
    }
    __syncthreads();

    // accB[dims 8*warp.., k] += Xc^T (w z)
    {
      const double *pa = sX + c * LDX + 8 * warp + r;
      const double *pb = sWZ + c * LDZ + r;
#pragma unroll
      for (int s = 0; s < BS / 4; ++s) {
        const double av = pa[(4 * s) * LDX];
#pragma unroll
        for (int ni = 0; ni < KT; ++ni)
          dmma884(accB[s & 1][ni][0], accB[s & 1][ni][1], av, pb[(4 * s) * LDZ + 8 * ni]);
      }
    }
    __syncthreads();  // all reads of this stage done before it is refilled two iterations later
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- write this CTA's partials ----
  const int64_t slab = blockIdx.y;
  {
    double *pB = a.pB + (slab * a.d64 + 64 * xb + 8 * warp + r) * a.kp;
#pragma unroll
    for (int ni = 0; ni < KT; ++ni) {
      const int col = 8 * ni + 2 * c;
      if (col < a.kp) {
        double2 *o = reinterpret_cast<double2 *>(pB + col);
        const double2 prev = *o;
        *o = make_double2(prev.x + (accB[0][ni][0] + accB[1][ni][0]), prev.y + (accB[0][ni][1] + accB[1][ni][1]));
      }
    }
  }
  double *red = stage0;  // [2][4][64] scratch
  red[(tid >> 6) * 64 + (tid & 63)] = sx;
  red[256 + (tid >> 6) * 64 + (tid & 63)] = tot;
  __syncthreads();
  if (tid < 64) {
    const double v = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
    const double o = red[256 + tid] + red[320 + tid] + red[384 + tid] + red[448 + tid];
    a.pT[slab * a.d64 + 64 * xb + tid] += v;
    a.pO[slab * a.d64 + 64 * xb + tid] += o;
  }
}

// fixed-order reduction of the per-slab partials into the statistics buffer (accumulating across chunks);
// tdev_i = Sx_i - sum_a C[i][a] MZ[i][a] with MZ = Mask^T (w Z) from the M-step contraction
__global__ void cross_moment_reduce_kernel(int slabs, int d64, int d, int kp, const double *pB, const double *pT,
                                           const double *pO, const double *__restrict__ Cpad,
                                           const double *__restrict__ MZ, double *statB, double *statTdev,
                                           double *statTotals) {
  const int64_t totalB = (int64_t)d * kp;
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  // eight independent loads in flight per round, summed in slab order (fixed order: run-to-run reproducible); one
  // chain of `slabs` dependent global loads cost 81 us per step at c2
  for (int64_t idx = gtid; idx < totalB; idx += gsz) {
    double s = 0.0;
    for (int z0 = 0; z0 < slabs; z0 += 8) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (z0 + j < slabs) ? __ldg(pB + (int64_t)(z0 + j) * d64 * kp + idx) : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    statB[idx] += s;
  }
  for (int64_t idx = gtid; idx < d; idx += gsz) {
    double s = 0.0, o = 0.0;
    for (int z0 = 0; z0 < slabs; z0 += 8) {
      double v[8], u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = z0 + j < slabs;
        v[j] = ok ? __ldg(pT + (int64_t)(z0 + j) * d64 + idx) : 0.0;
        u[j] = ok ? __ldg(pO + (int64_t)(z0 + j) * d64 + idx) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[j];
        o += u[j];
      }
    }
    double cz = 0.0;
    for (int q = 0; q < kp; ++q) cz = fma(Cpad[idx * kp + q], MZ[idx * kp + q], cz);
    statTdev[idx] += s - cz;
    statTotals[idx] += o;
  }
}

static int cr_bs(int kp) { (void)kp; return 32; }
static int cr_ctas_per_sm(int kp) {
  static const int force = getenv("PPCA_B200_CROSS_MINB") ? atoi(getenv("PPCA_B200_CROSS_MINB")) : 0;
  return force > 0 ? force : cr_minb(kp / 8);
}

int cross_resid_slabs(int d, int k, int rows, int sms) {
  Shape s(d, k);
  const int dblocks = (d + 63) / 64;
  const int ntiles = (rows + cr_bs(s.kp) - 1) / cr_bs(s.kp);
  int slabs = sms * cr_ctas_per_sm(s.kp) / dblocks;
  if (slabs < 1) slabs = 1;
  if (slabs > ntiles) slabs = ntiles;
  return slabs < 1 ? 1 : slabs;
}

size_t cross_resid_partials_len(int d, int k, int slabs_alloc) {
  Shape s(d, k);
  const int dblocks = (d + 63) / 64, d64 = dblocks * 64;
  return (size_t)slabs_alloc * ((size_t)d64 * s.kp + 2 * (size_t)d64);
}

template <int KT, int BS>
static void launch_cr(const Launcher &L, CrArgs a, int dblocks, int slabs) {
  using Cfg = CrCfg<KT, BS>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(cross_moment_kernel<KT, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM));
  }
  cross_moment_kernel<KT, BS><<<dim3(dblocks, slabs), 256, Cfg::SMEM, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_cross_resid(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m,
                        const double *WZ, const double *w, double *partials, int slabs_alloc) {
  if (rows <= 0) return;
  const int d = m.s.d, kp = m.s.kp;
  REQUIRE(kp <= 128, "state_size %d > 128 is not supported by the cross-moment kernel", m.s.k);
  const int dblocks = (d + 63) / 64, d64 = dblocks * 64;
  int slabs = cross_resid_slabs(d, m.s.k, rows, L.sms);
  if (slabs > slabs_alloc) slabs = slabs_alloc;
  CrArgs a;
  a.X = st.X.p; a.ldx = st.ldx; a.mask = st.mask.p; a.dw = st.dw; a.row0 = row0; a.rows = rows;
  a.kp = kp; a.d32 = m.s.d32; a.mupad = m.mu; a.WZ = WZ; a.w = w; a.d64 = d64;
  a.pB = partials;
  a.pT = a.pB + (size_t)slabs_alloc * d64 * kp;
  a.pO = a.pT + (size_t)slabs_alloc * d64;
  const int kt = kp / 8;
  if (kt <= 1) launch_cr<1, 32>(L, a, dblocks, slabs);
  else if (kt <= 2) launch_cr<2, 32>(L, a, dblocks, slabs);
  else if (kt <= 4) launch_cr<4, 32>(L, a, dblocks, slabs);
  else if (kt <= 8) launch_cr<8, 32>(L, a, dblocks, slabs);
  else launch_cr<16, 32>(L, a, dblocks, slabs);
}

void launch_cross_resid_finish(const Launcher &L, int d, int k, const double *partials, int slabs_alloc,
                               const double *Cpad, const double *MZ, double *statB, double *statTdev,
                               double *statTotals) {
  Shape s(d, k);
  const int kp = s.kp;
  const int dblocks = (d + 63) / 64, d64 = dblocks * 64;
  const double *pB = partials;
  const double *pT = pB + (size_t)slabs_alloc * d64 * kp;
  const double *pO = pT + (size_t)slabs_alloc * d64;
  const int64_t total = (int64_t)d * kp;
  const int blocks = (int)((total + 255) / 256 < 2 * L.sms ? (total + 255) / 256 : 2 * L.sms);
  cross_moment_reduce_kernel<<<blocks, 256, 0, L.stream>>>(slabs_alloc, d64, d, kp, pB, pT, pO, Cpad, MZ, statB, statTdev,
                                                           statTotals);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// ---------------------------------------------------------------------------------------------
// Exact residual norms, on demand.  The noise update needs sum_n w_n |R_n|^2 with R_n = m (x~ - C z) (OLD C, OLD mu;
// ppca_model.rs:338-346).  The solve kernels get |R_n|^2 from the identity nx - y^T z - sigma^2 |z|^2, which subtracts
// numbers of size |x~|^2: fine until the residual is ~1e-5 of the signal (a feature in other units, sigma -> 0).
// resid_check_kernel (solve.cu) flags such a chunk; this kernel then forms R on DMMA tiles ((samples, dims, k) GEMM
// Z C^T against the centred X tile) and sums the squares.  Unflagged chunks: every CTA returns at once.
// ---------------------------------------------------------------------------------------------
template <int KT, int BS>
struct RxCfg {
  static constexpr int LDX = 68;
  static constexpr int KPP = 8 * KT;
  static constexpr int LDZ = KPP + ((20 - KPP % 16) % 16);
  static constexpr int STAGE = BS * LDX + BS * LDZ + BS + BS;  // X, Z, w, mask words
  static constexpr int FIXED = 64 * LDZ + 64;                  // C block, mu
  static constexpr size_t SMEM = (size_t)(FIXED + 2 * STAGE) * sizeof(double);
  static constexpr int MI = BS / 32;
};

struct RxArgs {
  const double *X;
  int ldx;
  const uint32_t *mask;
  int dw;
  int64_t row0;
  int rows;
  const double *Cpad;
  int kp;
  int d32;
  const double *mupad;
  const double *Z;
  const double *w;
  const int *flag;
  double *pD;  // [slab][dimension block]
};

template <int KT, int BS>
__global__ void __launch_bounds__(256, (KT <= 4 ? 2 : 1)) resid_exact_kernel(RxArgs a) {
  if (*a.flag == 0) return;
  using Cfg = RxCfg<KT, BS>;
  constexpr int LDX = Cfg::LDX, LDZ = Cfg::LDZ, KPP = Cfg::KPP, MI = Cfg::MI, NI = 4;
  extern __shared__ __align__(16) double smem[];
  double *sC = smem;
  double *sMu = sC + 64 * LDZ;
  double *stage0 = sMu + 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int xb = blockIdx.x;
  const int ntiles = (a.rows + BS - 1) / BS;

  auto load_tile = [&](int st, int t) {
    double *sX = stage0 + st * Cfg::STAGE;
    double *sZ = sX + BS * LDX;
    double *sW = sZ + BS * LDZ;
    uint32_t *sM = reinterpret_cast<uint32_t *>(sW + BS);
    const int base = t * BS;
    for (int idx = tid; idx < BS * 32; idx += 256) {
      const int row = idx >> 5, col = (idx & 31) * 2;
      const int gcol = 64 * xb + col;
      const bool ok = gcol < a.ldx;
      const double *src = ok ? a.X + (a.row0 + base + row) * a.ldx + gcol : a.X;
      cp_async16(sX + row * LDX + col, src, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BS * (KPP / 2); idx += 256) {
      const int row = idx / (KPP / 2), col = (idx % (KPP / 2)) * 2;
      const bool ok = col < a.kp;
      cp_async16(sZ + row * LDZ + col, ok ? a.Z + (int64_t)(base + row) * a.kp + col : a.Z, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BS; idx += 256) cp_async8(sW + idx, a.w + base + idx, 8);
    for (int idx = tid; idx < BS * 2; idx += 256) {
      const int row = idx >> 1, j = idx & 1;
      const int wj = 2 * xb + j;
      const bool ok = wj < a.dw;
      cp_async4(sM + idx, ok ? a.mask + (a.row0 + base + row) * a.dw + wj : a.mask, ok ? 4 : 0);
    }
  };
  for (int idx = tid; idx < 64 * KPP; idx += 256) {
    const int row = idx / KPP, col = idx % KPP;
    const int gi = 64 * xb + row;
    sC[row * LDZ + col] = (gi < a.d32 && col < a.kp) ? a.Cpad[(int64_t)gi * a.kp + col] : 0.0;
  }
  if (tid < 64) sMu[tid] = (64 * xb + tid < a.d32) ? a.mupad[64 * xb + tid] : 0.0;

  double dev2 = 0.0;
  const int mt0 = MI * (warp & 3), nt0 = NI * (warp >> 2);
  int t = blockIdx.y;
  if (t < ntiles) load_tile(0, t);
  cp_async_commit();
  int st = 0;
  for (; t < ntiles; t += gridDim.y, st ^= 1) {
    const int tn = t + gridDim.y;
    if (tn < ntiles) load_tile(st ^ 1, tn);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    double *sX = stage0 + st * Cfg::STAGE;
    const double *sZ = sX + BS * LDX;
    const double *sW = sZ + BS * LDZ;
    const uint32_t *sM = reinterpret_cast<const uint32_t *>(sW + BS);
#pragma unroll
    for (int it = 0; it < BS * 64 / 256; ++it) {  // centre + select
      const int idx = tid + 256 * it;
      const int row = idx >> 6, col = idx & 63;
      const uint32_t wbits = sM[row * 2 + (col >> 5)];
      double *p = sX + row * LDX + col;
      *p = ((wbits >> (col & 31)) & 1u) ? (*p - sMu[col]) : 0.0;
    }
    __syncthreads();
    {
      double acc[MI][NI][2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const double2 xv =
              *reinterpret_cast<const double2 *>(sX + (8 * (mt0 + mi) + r) * LDX + 8 * (nt0 + ni) + 2 * c);
          acc[mi][ni][0] = -xv.x;
          acc[mi][ni][1] = -xv.y;
        }
      const double *pa = sZ + (8 * mt0 + r) * LDZ + c;
      const double *pb = sC + (8 * nt0 + r) * LDZ + c;
#pragma unroll
      for (int s = 0; s < 2 * KT; ++s) {
        double av[MI], bv[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) av[mi] = pa[(8 * mi) * LDZ + 4 * s];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bv[ni] = pb[(8 * ni) * LDZ + 4 * s];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {  // acc = Z C^T - Xc = -R on observed slots
        const int row = 8 * (mt0 + mi) + r;
        const double wn = sW[row];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const int col = 8 * (nt0 + ni) + 2 * c;
          const uint32_t wbits = sM[row * 2 + (col >> 5)] >> (col & 31);
          const double r0 = (wbits & 1u) ? acc[mi][ni][0] : 0.0;
          const double r1 = (wbits & 2u) ? acc[mi][ni][1] : 0.0;
          dev2 = fma(wn * r0, r0, dev2);
          dev2 = fma(wn * r1, r1, dev2);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();
  double *red = stage0;
  dev2 = warp_sum(dev2);
  if (lane == 0) red[warp] = dev2;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    a.pD[(int64_t)blockIdx.y * gridDim.x + xb] += s;
  }
}

__global__ void resid_exact_reduce_kernel(int count, const double *__restrict__ pD, double *scalars) {
  if (threadIdx.x < 32) {  // fixed order: lane-strided partial sums, then a shuffle tree
    double s = 0.0;
    for (int z = threadIdx.x; z < count; z += 32) s += pD[z];
    s = warp_sum(s);
    if (threadIdx.x == 0) scalars[SC_DEV2] += s;
  }
}

size_t resid_exact_partials_len(int d, int slabs_alloc) { return (size_t)slabs_alloc * ((d + 63) / 64); }

template <int KT, int BS>
static void launch_rx(const Launcher &L, const RxArgs &a, int dblocks, int slabs) {
  using Cfg = RxCfg<KT, BS>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(resid_exact_kernel<KT, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)Cfg::SMEM));
  }
  resid_exact_kernel<KT, BS><<<dim3(dblocks, slabs), 256, Cfg::SMEM, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_resid_exact(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m, const double *Z,
                        const double *w, const int *flag, double *pD, int slabs_alloc) {
  if (rows <= 0) return;
  const int d = m.s.d, kp = m.s.kp;
  const int dblocks = (d + 63) / 64;
  int slabs = cross_resid_slabs(d, m.s.k, rows, L.sms);
  if (slabs > slabs_alloc) slabs = slabs_alloc;
  RxArgs a;
  a.X = st.X.p; a.ldx = st.ldx; a.mask = st.mask.p; a.dw = st.dw; a.row0 = row0; a.rows = rows;
  a.Cpad = m.C; a.kp = kp; a.d32 = m.s.d32; a.mupad = m.mu; a.Z = Z; a.w = w; a.flag = flag; a.pD = pD;
  const int kt = kp / 8;
  if (kt <= 1) launch_rx<1, 32>(L, a, dblocks, slabs);
  else if (kt <= 2) launch_rx<2, 32>(L, a, dblocks, slabs);
  else if (kt <= 4) launch_rx<4, 32>(L, a, dblocks, slabs);
  else if (kt <= 8) launch_rx<8, 32>(L, a, dblocks, slabs);
  else launch_rx<16, 32>(L, a, dblocks, slabs);
}

void launch_resid_exact_finish(const Launcher &L, int d, const double *pD, int slabs_alloc, double *scalars) {
  resid_exact_reduce_kernel<<<1, 32, 0, L.stream>>>((int)resid_exact_partials_len(d, slabs_alloc), pD, scalars);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// ---------------------------------------------------------------------------------------------
// reconstruction writer: 64 samples x 64 dims per CTA, K = k
// ---------------------------------------------------------------------------------------------
struct RecArgs {
  const double *X;
  int ldx;
  const uint32_t *mask;
  int dw;
  int64_t row0;
  int rows;
  int d;
  const double *Cpad;
  int kp;
  int d32;
  const double *mupad;
  const double *Z;
  int extrapolate;
  const double *scale;
  int64_t scale_ld;
  int accumulate;
  double *out;
  int64_t ldo;
};

template <int KT>
__global__ void __launch_bounds__(256) reconstruct_kernel(RecArgs a) {
  constexpr int KPP = 8 * KT;
  constexpr int LDZ = KPP + ((20 - KPP % 16) % 16);
  constexpr int LDO = 66;  // staging pitch of the 64 x 64 output tile
  extern __shared__ __align__(16) double smem[];
  double *sZ = smem;
  double *sC = smem + 64 * LDZ;
  double *sO = smem;  // Z and C tiles are dead after the MMA phase (the launch sizes smem for whichever is larger)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int xb = blockIdx.x, base = blockIdx.y * 64;
  for (int idx = tid; idx < 64 * KPP; idx += 256) {
    const int row = idx / KPP, col = idx % KPP;
    const int gi = 64 * xb + row;
    sC[row * LDZ + col] = (gi < a.d32 && col < a.kp) ? a.Cpad[(int64_t)gi * a.kp + col] : 0.0;
    sZ[row * LDZ + col] = (base + row < a.rows && col < a.kp) ? a.Z[(int64_t)(base + row) * a.kp + col] : 0.0;
  }
  __syncthreads();
  const int mt0 = 2 * (warp & 3), nt0 = 4 * (warp >> 2);
  double acc[2][4][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  const double *pa = sZ + (8 * mt0 + r) * LDZ + c;
  const double *pb = sC + (8 * nt0 + r) * LDZ + c;
#pragma unroll
  for (int s = 0; s < 2 * KT; ++s) {
    double av[2], bv[4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) av[mi] = pa[(8 * mi) * LDZ + 4 * s];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) bv[ni] = pb[(8 * ni) * LDZ + 4 * s];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
  }
  __syncthreads();  // all fragment reads of sZ / sC are done: reuse sZ as the output staging tile
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2 *>(sO + (8 * (mt0 + mi) + r) * LDO + 8 * (nt0 + ni) + 2 * c) =
          make_double2(acc[mi][ni][0], acc[mi][ni][1]);
  __syncthreads();
  // coalesced pass: warp w handles rows w, w + 8, ...; lanes walk the 64 columns of the dimension block
  for (int row = warp; row < 64; row += 8) {
    if (base + row >= a.rows) break;
    const int64_t grow = a.row0 + base + row;
    const double sc = a.scale ? a.scale[(int64_t)(base + row) * a.scale_ld] : 1.0;
    uint32_t w0 = 0u, w1 = 0u;
    if (a.extrapolate) {
      if (2 * xb < a.dw) w0 = a.mask[grow * a.dw + 2 * xb];
      if (2 * xb + 1 < a.dw) w1 = a.mask[grow * a.dw + 2 * xb + 1];
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = 64 * xb + 32 * hh + lane;
      if (col >= a.d) continue;
      double v = sO[row * LDO + 32 * hh + lane] + a.mupad[col];
      if (a.extrapolate && (((hh ? w1 : w0) >> lane) & 1u)) v = a.X[grow * a.ldx + col];  // observed: copy, never recompute
      if (a.scale) v *= sc;
      double *o = a.out + (int64_t)(base + row) * a.ldo + col;
      *o = a.accumulate ? (*o + v) : v;
    }
  }
}

template <int KT>
static void launch_rec(const Launcher &L, const RecArgs &a, dim3 grid) {
  constexpr int KPP = 8 * KT;
  constexpr int LDZ = KPP + ((20 - KPP % 16) % 16);
  constexpr size_t SMEM = (2 * 64 * LDZ > 64 * 66 ? 2 * 64 * LDZ : 64 * 66) * sizeof(double);
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(reconstruct_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  }
  reconstruct_kernel<KT><<<grid, 256, SMEM, L.stream>>>(a);
}

void launch_reconstruct(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m,
                        const double *Z, int extrapolate, const double *scale, int64_t scale_ld, int accumulate,
                        double *outX, int64_t ldo) {
  if (rows <= 0) return;
  REQUIRE(m.s.kp <= 128, "state_size %d > 128 is not supported by the reconstruction kernel", m.s.k);
  RecArgs a;
  a.X = st.X.p; a.ldx = st.ldx; a.mask = st.mask.p; a.dw = st.dw; a.row0 = row0; a.rows = rows; a.d = m.s.d;
  a.Cpad = m.C; a.kp = m.s.kp; a.d32 = m.s.d32; a.mupad = m.mu; a.Z = Z; a.extrapolate = extrapolate;
  a.scale = scale; a.scale_ld = scale_ld; a.accumulate = accumulate; a.out = outX; a.ldo = ldo;
  dim3 grid((m.s.d + 63) / 64, (rows + 63) / 64);
  const int kt = m.s.kp / 8;
  if (kt <= 1) launch_rec<1>(L, a, grid);
  else if (kt <= 2) launch_rec<2>(L, a, grid);
  else if (kt <= 4) launch_rec<4>(L, a, grid);
  else if (kt <= 8) launch_rec<8>(L, a, grid);
  else launch_rec<16>(L, a, grid);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

}  // namespace ppca
