// proj.cu — masked projection  Y = Xc * C  and  nx_n = |Xc_n|^2  with  Xc[n][i] = m_ni ? x_ni - mu_i : 0.
//
// Reference: per sample, Mask::mask(x - mean) (utils.rs:56-61, ppca_model.rs:131,200) and C_o^T x~ inside
// quadratic_form / estimator_transform (output_covariance.rs:135,90-94).  Here it is one skinny FP64 GEMM
// (samples x d x k) on DMMA.  The centred tile is SELECTED (never multiplied by the mask) in shared memory
// before the MMA phase, so non-finite inputs can never leak (masked slots are stored as 0.0 anyway).
#include "common.cuh"
#include "mma.cuh"

namespace ppca {

template <int NI>
struct ProjCfg {
  static constexpr int BM = 128, LDA = 36, STAGES = 2, THREADS = 256;
  static constexpr int BN = 8 * NI;
  static constexpr int LDC = BN + ((20 - BN % 16) % 16);
  static constexpr int X_STAGE = BM * LDA, C_STAGE = 32 * LDC;
  static constexpr size_t SMEM = (size_t)STAGES * (X_STAGE + C_STAGE) * sizeof(double);
};

template <int NI>
__global__ void __launch_bounds__(256, 2)
    proj_kernel(const double *__restrict__ X, int ldx, const uint32_t *__restrict__ mask, int dw, int64_t row0,
                const double *__restrict__ Cpad, int kp, const double *__restrict__ mupad, int nkb, int d,
                double *Y, double *nx) {
  using Cfg = ProjCfg<NI>;
  constexpr int BM = Cfg::BM, LDA = Cfg::LDA, LDC = Cfg::LDC, STAGES = Cfg::STAGES, BN = Cfg::BN;
  extern __shared__ __align__(16) double smem[];
  double *sXall = smem;
  double *sCall = smem + STAGES * Cfg::X_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  auto load_stage = [&](int stage, int kb) {
    double *sX = sXall + stage * Cfg::X_STAGE;
    double *sC = sCall + stage * Cfg::C_STAGE;
    for (int idx = tid; idx < BM * 16; idx += 256) {
      const int row = idx >> 4, col = (idx & 15) * 2;
      const int gcol = kb * 32 + col;
      const bool ok = gcol < ldx;
      const double *src = ok ? X + (row0 + m0 + row) * ldx + gcol : X;
      cp_async16(sX + row * LDA + col, src, ok ? 16 : 0);
    }
    for (int idx = tid; idx < 32 * (BN / 2); idx += 256) {
      const int row = idx / (BN / 2), col = (idx % (BN / 2)) * 2;
      const bool ok = (n0 + col) < kp;
      const double *src = ok ? Cpad + (int64_t)(kb * 32 + row) * kp + n0 + col : Cpad;
      cp_async16(sC + row * LDC + col, src, ok ? 16 : 0);
    }
  };

  double acc[2][NI][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  double nxacc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) nxacc[j] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }
  const uint32_t *mrow = mask + (row0 + m0 + 16 * warp) * dw;

  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kb + STAGES - 1;
      if (nxt < nkb) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    double *sX = sXall + (kb % STAGES) * Cfg::X_STAGE;
    const double *sC = sCall + (kb % STAGES) * Cfg::C_STAGE;
    // centre + select this warp's 16 rows (warp-local: only this warp reads them as A fragments)
    {
      const double mu = __ldg(mupad + kb * 32 + lane);
      uint32_t words[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) words[j] = __ldg(mrow + (int64_t)j * dw + kb);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double *p = sX + (16 * warp + j) * LDA + lane;
        const double xc = ((words[j] >> lane) & 1u) ? (*p - mu) : 0.0;
        *p = xc;
        nxacc[j] = fma(xc, xc, nxacc[j]);
      }
    }
    __syncwarp();
    const double *sA = sX + (16 * warp + r) * LDA + c;
    const double *sB = sC + c * LDC + r;
    const int smax = min(8, (d - 32 * kb + 3) >> 2);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s >= smax) break;
      double a[2], b[NI];
      a[0] = sA[4 * s];
      a[1] = sA[8 * LDA + 4 * s];
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) b[ni] = sB[(4 * s) * LDC + 8 * ni];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const int row = m0 + 16 * warp + 8 * mi + r;
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) {
      const int col = n0 + 8 * ni + 2 * c;
      if (col < kp) *reinterpret_cast<double2 *>(Y + (int64_t)row * kp + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
  }
  if (blockIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double s = warp_sum(nxacc[j]);
      if (lane == 0) nx[m0 + 16 * warp + j] = s;
    }
  }
}

template <int NI>
static void launch_proj_ni(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m,
                           double *Y, double *nx) {
  using Cfg = ProjCfg<NI>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(proj_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  dim3 grid((unsigned)((rows + Cfg::BM - 1) / Cfg::BM), (unsigned)((m.s.kp + Cfg::BN - 1) / Cfg::BN));
  proj_kernel<NI><<<grid, 256, Cfg::SMEM, L.stream>>>(st.X.p, st.ldx, st.mask.p, st.dw, row0, m.C, m.s.kp, m.mu,
                                                      m.s.d32 / 32, m.s.d, Y, nx);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

// Y: rows_pad x kp (rows up to the next multiple of 128 are written), nx: rows_pad
void launch_proj(const Launcher &L, const SampleStore &st, int64_t row0, int rows, const DevModel &m, double *Y,
                 double *nx) {
  if (rows <= 0) return;
  const int nt = m.s.kp / 8;
  if (nt <= 1) launch_proj_ni<1>(L, st, row0, rows, m, Y, nx);
  else if (nt <= 2) launch_proj_ni<2>(L, st, row0, rows, m, Y, nx);
  else if (nt <= 4) launch_proj_ni<4>(L, st, row0, rows, m, Y, nx);
  else launch_proj_ni<8>(L, st, row0, rows, m, Y, nx);
}


// Dense FP64 row GEMM on the same kernel:  Y[rows_pad x n8] = A[rows_pad x K] * Bt[K32 x n8]  (no masking, no centring).
// `ones` must hold K32 / 32 words of 0xffffffff (every row reads the same words: mask pitch 0), `zeros` K32 doubles of 0.
template <int NI>
static void launch_rowgemm_ni(const Launcher &L, const double *A, int lda, int rows_pad, int K, const double *Bt, int n8,
                              const uint32_t *ones, const double *zeros, double *Y, double *nx_scratch) {
  using Cfg = ProjCfg<NI>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(proj_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  dim3 grid((unsigned)(rows_pad / Cfg::BM), (unsigned)((n8 + Cfg::BN - 1) / Cfg::BN));
  proj_kernel<NI><<<grid, 256, Cfg::SMEM, L.stream>>>(A, lda, ones, 0, 0, Bt, n8, zeros, (K + 31) / 32, K, Y, nx_scratch);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_rowgemm(const Launcher &L, const double *A, int lda, int rows_pad, int K, const double *Bt, int n8,
                    const uint32_t *ones, const double *zeros, double *Y, double *nx_scratch) {
  REQUIRE(rows_pad % 128 == 0 && n8 % 8 == 0 && lda % 2 == 0, "rowgemm: rows must be padded to 128, columns to 8");
  if (rows_pad <= 0) return;
  const int nt = n8 / 8;
  if (nt <= 1) launch_rowgemm_ni<1>(L, A, lda, rows_pad, K, Bt, n8, ones, zeros, Y, nx_scratch);
  else if (nt <= 2) launch_rowgemm_ni<2>(L, A, lda, rows_pad, K, Bt, n8, ones, zeros, Y, nx_scratch);
  else if (nt <= 4) launch_rowgemm_ni<4>(L, A, lda, rows_pad, K, Bt, n8, ones, zeros, Y, nx_scratch);
  else launch_rowgemm_ni<8>(L, A, lda, rows_pad, K, Bt, n8, ones, zeros, Y, nx_scratch);
}

}  // namespace ppca
