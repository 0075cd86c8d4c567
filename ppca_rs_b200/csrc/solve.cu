// solve.cu — per-sample k x k posterior algebra, one warp per sample.
//
// Input per sample: G_n = C_o^T C_o (packed upper, from the masked-Gram contraction), y_n = C_o^T x~_n,
// nx_n = |x~_n|^2, d_n = #observed.  With M_n = sigma^2 I + G_n = L L^T (output_covariance.rs:61-64):
//   z_n     = M_n^{-1} y_n            == estimator_transform * x~   (output_covariance.rs:90-94, ppca_model.rs:205)
//   Sigma_n = sigma^2 M_n^{-1}        == I - T C_o                  (output_covariance.rs:98-101, ppca_model.rs:206)
//   llk_n   = -(nx - y^T M^{-1} y)/(2 sigma^2) - (ln det M + 2 ln sigma (d_n - k))/2 - ln(2 pi) d_n / 2
//                                                                   (ppca_model.rs:131-138, output_covariance.rs:115-142)
//   W_n     = w_n (z z^T + Sigma_n)   second moment                 (ppca_model.rs:303, :437-439)
//   t_n     = tr(Sigma_n G_n)         == (C_o Sigma_n).dot(C_o)     (ppca_model.rs:345)
// The reference uses an LU inverse and ln(det) by LU; M_n is SPD so the Cholesky route agrees to rounding
// and cannot overflow the determinant.  Empty samples: z = 0, Sigma = I, llk = 0 (ppca_model.rs:98-104,125-129).
//
// This is the generic (any k) shared-memory kernel: the (k+1) x k augmented matrix [M ; y^T] lives in shared
// memory, one per warp; L^{-1} is built in the unused strict upper triangle.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mma.cuh"

namespace ppca {

#define LN_2PI 1.8378770664093453

// 1 / x for a positive normal double: hardware seed (relative error <= 2^-23) and two Newton steps
// (2^-46, then rounding level); no special-case branches.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

struct SolveSmem {
  int ldk;      // odd pitch
  int per_warp; // doubles per warp
  __host__ __device__ SolveSmem(int k) {
    ldk = (k | 1);
    per_warp = (k + 1) * ldk + 2 * (k > 0 ? k : 1);
    per_warp = (per_warp + 1) & ~1;
  }
};

__global__ void __launch_bounds__(256) solve_kernel(SolveArgs a) {
  extern __shared__ double smem_solve[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int k = a.s.k, kk = a.s.kk, kkp = a.s.kkp, kp = a.s.kp;
  const SolveSmem lay(k);
  const int ldk = lay.ldk;
  double *M = smem_solve + (size_t)wi * lay.per_warp;  // (k+1) x ldk ; row k = y^T
  double *dinv = M + (k + 1) * ldk;                    // 1 / L_pp
  double *zv = dinv + k;
  const double sigma = a.sigma_dev ? *a.sigma_dev : a.sigma;  // device copy: the launch can be replayed from a CUDA graph
  const double s2 = sigma * sigma;
  const double ln_sigma = log(sigma);

  for (int row = blockIdx.x * warps + wi; row < a.rows_pad; row += gridDim.x * warps) {
    double *G = a.GW ? a.GW + (int64_t)row * kkp : nullptr;
    double *y = a.YZ + (int64_t)row * kp;
    const int dn = row < a.rows ? a.dn[row] : 0;
    if (dn == 0) {  // padding row or empty sample
      if (a.mode == 2) {
        for (int q = lane; q < kkp; q += 32) G[q] = 0.0;
        if (a.WZ)
          for (int q = lane; q < kp; q += 32) a.WZ[(int64_t)row * kp + q] = 0.0;
      }
      for (int q = lane; q < kp; q += 32) y[q] = 0.0;
      if (row < a.rows) {
        if (a.llk && lane == 0) a.llk[row] = 0.0;
        if (a.tn && lane == 0) a.tn[row] = 0.0;
        if (a.dv && lane == 0) a.dv[row] = 0.0;
        if (a.cov)
          for (int q = lane; q < k * k; q += 32) a.cov[(int64_t)row * k * k + q] = (q / k == q % k) ? 1.0 : 0.0;
      }
      continue;
    }
    const double w = a.w ? a.w[row] : 1.0;

    // 1. unpack [M ; y]
    for (int p = 0; p < k; ++p) {
      const int off = tri_row_off(p, k) - p;
      for (int b = p + lane; b < k; b += 32) M[b * ldk + p] = G[off + b] + (b == p ? s2 : 0.0);
    }
    for (int q = lane; q < k; q += 32) M[k * ldk + q] = y[q];
    __syncwarp();
    if (a.gscale) {  // precision guard of the int8-sliced Gram matrix (see SolveArgs)
      const double terms = guard_terms((double)dn) * a.guard_coef;
      for (int p = lane; p < k; p += 32)
        if (terms * a.gscale[tri_row_off(p, k)] > M[p * ldk + p]) atomicAdd(a.unsafe, 1u);
    }

    // 2. Cholesky of the augmented matrix: rows 0..k-1 -> L, row k -> u = L^{-1} y
    double logdet = 0.0;
    for (int p = 0; p < k; ++p) {
      const double dpp = M[p * ldk + p];
      const double inv = rsqrt(dpp);
      logdet += log(dpp);
      __syncwarp();
      for (int rr = p + 1 + lane; rr <= k; rr += 32) M[rr * ldk + p] *= inv;
      if (lane == 0) dinv[p] = inv;
      __syncwarp();
      for (int cc = p + 1; cc < k; ++cc) {
        const double lcp = M[cc * ldk + p];
        for (int rr = cc + lane; rr <= k; rr += 32) M[rr * ldk + cc] = fma(-M[rr * ldk + p], lcp, M[rr * ldk + cc]);
      }
      __syncwarp();
    }
    double quad = 0.0;  // y^T M^{-1} y = |u|^2
    for (int q = lane; q < k; q += 32) {
      const double u = M[k * ldk + q];
      quad = fma(u, u, quad);
    }
    quad = warp_sum(quad);
    if (a.llk) {
      const double llk = -0.5 * (a.nx[row] - quad) / s2 - 0.5 * (logdet + 2.0 * ln_sigma * (double)(dn - k)) -
                         0.5 * LN_2PI * (double)dn;
      if (lane == 0) a.llk[row] = llk;
    }
    if (a.mode == 0) {
      __syncwarp();
      continue;
    }

    // 3. X = L^{-1} into the strict upper triangle: X[r][j] (r > j) at M[j][r]; X[j][j] = dinv[j]
    for (int j = lane; j < ((k + 31) & ~31); j += 32) {
      const bool live = j < k;
      const double xjj = live ? dinv[j] : 0.0;
      for (int rr = 1; rr < k; ++rr) {
        double s = 0.0;
        for (int cc = 0; cc < rr; ++cc) {
          const double lrc = M[rr * ldk + cc];
          double x = 0.0;
          if (live && cc >= j) x = (cc == j) ? xjj : M[j * ldk + cc];
          s = fma(lrc, x, s);
        }
        if (live && rr > j) M[j * ldk + rr] = -s * dinv[rr];
      }
    }
    __syncwarp();

    // 4. z = X^T u
    for (int q = lane; q < k; q += 32) {
      double s = dinv[q] * M[k * ldk + q];
      for (int rr = q + 1; rr < k; ++rr) s = fma(M[q * ldk + rr], M[k * ldk + rr], s);
      zv[q] = s;
    }
    __syncwarp();
    for (int q = lane; q < kp; q += 32) {
      const double z = q < k ? zv[q] : 0.0;
      y[q] = z;
      if (a.WZ) a.WZ[(int64_t)row * kp + q] = w * z;
    }

    // 5. M^{-1} = X^T X (packed), W, t, optional full covariance
    if (a.mode == 2 || a.cov) {
      double tpart = 0.0;
      for (int p = 0; p < k; ++p) {
        const int off = tri_row_off(p, k) - p;
        const double zp = zv[p];
        for (int b0 = p; b0 < k; b0 += 32) {
          const int b = b0 + lane;
          const bool live = b < k;
          double s = 0.0;
          if (live) s = (b == p) ? dinv[p] * dinv[p] : M[p * ldk + b] * dinv[b];
          for (int rr = b0 + 1; rr < k; ++rr) {
            const double xp = M[p * ldk + rr];
            if (live && rr > b) s = fma(xp, M[b * ldk + rr], s);
          }
          if (live) {
            if (a.mode == 2) {
              const double g = G[off + b];
              tpart = fma((b == p ? 1.0 : 2.0) * s, g, tpart);
              G[off + b] = w * fma(zp, zv[b], s2 * s);
            }
            if (a.cov) {
              a.cov[(int64_t)row * k * k + p * k + b] = s2 * s;
              a.cov[(int64_t)row * k * k + b * k + p] = s2 * s;
            }
          }
        }
      }
      if (a.mode == 2) {
        for (int q = kk + lane; q < kkp; q += 32) G[q] = 0.0;
        double zz = 0.0;
        for (int q = lane; q < k; q += 32) zz = fma(zv[q], zv[q], zz);
        tpart = warp_sum(tpart);
        zz = warp_sum(zz);
        if (a.tn && lane == 0) a.tn[row] = s2 * tpart;
        if (a.dv && lane == 0) a.dv[row] = (a.nx[row] - quad) - s2 * zz;  // |R_n|^2, see SolveArgs::dv
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant for k <= 32: KP (8/16/32) lanes per sample, lane i owns row i of the symmetric
// matrix in registers.  M_n is inverted in place by the symmetric sweep operator in its square-root form:
//   pivot p (d = T[p][p], r = 1/sqrt(d)):   s_i = T[i][p] r ;  T[i][j] -= s_i s_j (i, j != p) ;  T[i][p] = s_i r ;  T[p][p] = -1/d
// after all pivots T = -M^-1.  The update product s_i s_j is commutative, so T[i][j] and T[j][i] stay BITWISE equal and
// the pivot row never has to be broadcast: every lane contributes its own column-p entry s_i through one shared-memory
// store per pivot and reads the k of them back (by symmetry that IS the pivot row).  The pivot row itself is left
// unscaled (stored = d x true) and scaled once at the end; a pivoted lane publishes stored x 1/d.
// Why this form: round 1 used the non-commutative product (T[i][p]/d) T[j][p]; the two triangles then follow different
// rounding paths, and on ill-conditioned M_n (one feature in other units: M = small + big v v^T) rebuilding the row from
// the column cost four digits of y^T M^-1 y against LU / Cholesky.  With the commutative product the elimination
// matches them (tools/solve_accuracy.py, profiles/r02_solve_accuracy.md).  The pivots are the Cholesky pivots squared,
// so ln det M = sum ln d_p (no determinant overflow, unlike determinant().ln() at output_covariance.rs:117).
// ---------------------------------------------------------------------------------------------
// 1 / sqrt(x) for a positive normal double: hardware seed (2^-22) and one third-order step (-> 2^-66, then rounding)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// PF: the packed rows of the NEXT group of samples are fetched with cp.async into a second staging buffer while this group
// is eliminated (ncu: long_scoreboard 2.0 of 10.9 stall cycles per instruction = the warp waiting for its own G loads).
template <int KP, int MINB, bool PF>
__global__ void __launch_bounds__(256, MINB) solve_reg_kernel(SolveArgs a) {
  constexpr int SPW = 32 / KP;
  constexpr int NST = PF ? 2 : 1;
  extern __shared__ __align__(16) double smem_reg[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int sub = lane / KP, li = lane % KP;
  const int k = a.s.k, kkp = a.s.kkp, kp = a.s.kp;
  const int per_warp = NST * SPW * kkp + 128;
  double *stage0 = smem_reg + (size_t)wi * per_warp;  // NST x SPW packed rows
  double *col = stage0 + NST * SPW * kkp;             // [2][32] pivot-column exchange
  double *yb = col + 64;                             // [32]
  double *zb = yb + 32;                              // [32]; zb[31 - ...] is never read beyond KP per sample
  // running max |W| per staged slot of this warp (column maxima for the int8 digit planes, fused here so the
  // M-step slicing does not need its own pass over W)
  double *cmw = smem_reg + (size_t)warps * per_warp + (size_t)wi * (SPW * kkp);
  if (a.colmax)
    for (int q = lane; q < SPW * kkp; q += 32) cmw[q] = 0.0;
  const double sigma = a.sigma_dev ? *a.sigma_dev : a.sigma;  // device copy: the launch can be replayed from a CUDA graph
  const double s2 = sigma * sigma;
  const double ln_sigma = log(sigma);
  const int groups = (a.rows_pad + SPW - 1) / SPW;
  // row li of the packed symmetric matrix: (li, j >= li) sits at up + j, (j < li, li) at j (2k - j - 1) / 2 + li.
  // The gather offsets do not depend on the sample: computed once (padding lanes / columns: -1).
  const int up = tri_row_off(li, k) - li;
  int gi[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j)
    gi[j] = (li < k && j < k) ? ((j >= li) ? up + j : ((j * (2 * k - j - 1)) >> 1) + li) : -1;

  const int gstride = gridDim.x * warps;
  auto fetch = [&](int g2, double *dst) {  // 16-byte asynchronous copies of one group's packed rows (PF only)
    if (g2 < groups) {
      const double *src = a.GW + (int64_t)g2 * SPW * kkp;
      for (int q = lane * 2; q < SPW * kkp; q += 64) cp_async16(dst + q, src + q, 16);
    }
    cp_async_commit();
  };
  int it = 0;
  if constexpr (PF) fetch(blockIdx.x * warps + wi, stage0);
  for (int g = blockIdx.x * warps + wi; g < groups; g += gstride, ++it) {
    const int row0 = g * SPW;
    const int row = row0 + sub;
    double *gsrc = a.GW + (int64_t)row0 * kkp;
    double *stage = stage0 + (PF ? (it & 1) * SPW * kkp : 0);
    if constexpr (PF) {
      fetch(g + gstride, stage0 + ((it + 1) & 1) * SPW * kkp);  // the other buffer's last reader finished before the warp sync that ended the previous iteration
    } else {
      for (int q = lane * 2; q < SPW * kkp; q += 64)
        *reinterpret_cast<double2 *>(stage + q) = *reinterpret_cast<const double2 *>(gsrc + q);
    }
    yb[lane] = (li < kp) ? a.YZ[(int64_t)row * kp + li] : 0.0;
    const int dn = row < a.rows ? a.dn[row] : 0;
    const bool empty = dn == 0;
    const double w = (a.w && row < a.rows) ? a.w[row] : (row < a.rows ? 1.0 : 0.0);
    if constexpr (PF) cp_async_wait<1>();
    __syncwarp();

    const double *st = stage + sub * kkp;
    const bool live = li < k && !empty;
    double A[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {  // M = sigma^2 I + G where live, the identity elsewhere
      const bool use = !empty && gi[j] >= 0;
      const double g = st[use ? gi[j] : 0];
      const double unit = (j == li) ? 1.0 : 0.0;
      A[j] = use ? fma(unit, s2, g) : unit;
    }
    if (a.gscale && live) {  // precision guard of the int8-sliced Gram matrix (see SolveArgs); lane li checks M_ii
      const int qd = up + li;
      if (guard_terms((double)dn) * a.guard_coef * a.gscale[qd] > s2 + st[qd]) atomicAdd(a.unsafe, 1u);
    }

    double mypiv = 1.0, sc = 1.0;  // d_i, and 1 / d_i once row li has been the pivot (1 before)
#pragma unroll
    for (int p = 0; p < KP; ++p) {
      const double dpp = __shfl_sync(0xffffffffu, A[p], sub * KP + p);  // current pivot, from the lane that owns row p
      const double rinv = fast_rsqrt(dpp);
      const double own = A[p] * rinv;  // s_i in this row's units
      double *cb = col + (p & 1) * 32;
      cb[lane] = own * sc;             // s_i in true units for everybody else
      __syncwarp();
      const bool isp = li == p;
      const double f = isp ? 0.0 : own;  // the pivot row is left as it is (its 1/d is applied at the end)
      const double2 *cs2 = reinterpret_cast<const double2 *>(cb + sub * KP);
#pragma unroll
      for (int j = 0; j < KP; j += 2) {
        const double2 cv = cs2[j >> 1];
        if (j != p) A[j] = fma(-f, cv.x, A[j]);
        if (j + 1 != p) A[j + 1] = fma(-f, cv.y, A[j + 1]);
      }
      A[p] = isp ? -1.0 : own * rinv;
      if (isp) {
        mypiv = dpp;
        sc = rinv * rinv;
      }
    }
    // M^-1[li][j] = -sc A[j]; the factor is folded into what follows instead of a pass over the row
    const double nsc = -sc;

    // z = M^{-1} y ; quad = y^T z ; ln det
    double zi = 0.0;
    {
      const double *ys = yb + sub * KP;
#pragma unroll
      for (int j = 0; j < KP; ++j) zi = fma(A[j], ys[j], zi);
    }
    zi = live ? zi * nsc : 0.0;
    double quad = yb[lane] * zi;
    double logdet = live ? log(mypiv) : 0.0;
#pragma unroll
    for (int o = KP / 2; o > 0; o >>= 1) {
      quad += __shfl_xor_sync(0xffffffffu, quad, o);
      logdet += __shfl_xor_sync(0xffffffffu, logdet, o);
    }
    if (a.llk && li == 0 && row < a.rows) {
      double llk = 0.0;
      if (!empty)
        llk = -0.5 * (a.nx[row] - quad) / s2 - 0.5 * (logdet + 2.0 * ln_sigma * (double)(dn - k)) -
              0.5 * LN_2PI * (double)dn;
      a.llk[row] = llk;
    }
    if (a.mode == 0) {
      __syncwarp();
      continue;
    }
    zb[lane] = zi;
    if (li < kp) {
      a.YZ[(int64_t)row * kp + li] = zi;
      if (a.WZ) a.WZ[(int64_t)row * kp + li] = w * zi;
    }
    if (a.cov && row < a.rows && li < k) {
      double *cv = a.cov + (int64_t)row * k * k + (int64_t)li * k;
      const double cs = s2 * nsc;
#pragma unroll
      for (int j = 0; j < KP; ++j)
        if (j < k) cv[j] = empty ? (j == li ? 1.0 : 0.0) : cs * A[j];
    }
    if (a.mode == 2) {
      // t = tr(Sigma G) = sigma^2 tr(M^{-1} (M - sigma^2 I)) = sigma^2 sum_i (1 - sigma^2 M^{-1}_ii)
      double tpart = 0.0;
      {
        double diag = 0.0;
#pragma unroll
        for (int j = 0; j < KP; ++j)
          if (j == li) diag = A[j];
        if (live) tpart = fma(-s2 * nsc, diag, 1.0);
      }
      double zz = zi * zi;
#pragma unroll
      for (int o = KP / 2; o > 0; o >>= 1) {
        tpart += __shfl_xor_sync(0xffffffffu, tpart, o);
        zz += __shfl_xor_sync(0xffffffffu, zz, o);
      }
      if (li == 0 && row < a.rows) {
        if (a.tn) a.tn[row] = empty ? 0.0 : s2 * tpart;
        if (a.dv) a.dv[row] = empty ? 0.0 : (a.nx[row] - quad) - s2 * zz;  // |R_n|^2, see SolveArgs::dv
      }
      __syncwarp();  // everyone is done reading G and zb is visible
      // W = w (z z^T + sigma^2 M^{-1}), upper triangle of row li, packed in place
      if (li < k) {
        double *so = stage + sub * kkp + up;
        const double *zs = zb + sub * KP;
        const double ws2 = empty ? 0.0 : w * s2 * nsc, wzi = empty ? 0.0 : w * zi;
#pragma unroll
        for (int j = 0; j < KP; ++j)
          if (j >= li && j < k) so[j] = fma(wzi, zs[j], ws2 * A[j]);
      }
      __syncwarp();
      for (int q = lane * 2; q < SPW * kkp; q += 64) {
        const double2 v = *reinterpret_cast<const double2 *>(stage + q);
        *reinterpret_cast<double2 *>(gsrc + q) = v;
        if (a.colmax) {
          double2 m = *reinterpret_cast<const double2 *>(cmw + q);
          m.x = fmax(m.x, fabs(v.x));
          m.y = fmax(m.y, fabs(v.y));
          *reinterpret_cast<double2 *>(cmw + q) = m;
        }
      }
    }
    __syncwarp();
  }
  if constexpr (PF) cp_async_wait<0>();
  if (a.colmax) {
    __syncthreads();
    const double *all = smem_reg + (size_t)warps * per_warp;
    for (int c = threadIdx.x; c < kkp; c += blockDim.x) {
      double m = 0.0;
      for (int j = 0; j < warps * SPW; ++j) m = fmax(m, all[(size_t)j * kkp + c]);
      if (m > 0.0) atomicMax(a.colmax + c, (unsigned long long)__double_as_longlong(m));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Register-tiled variant (16 < k <= 64, the default): thread (tr, tc) owns the TR x TC tile rows [TR tr, TR tr + TR),
// columns [TC tc, TC tc + TC) of the symmetric matrix.  Same elimination as above (square-root symmetric sweep, swept
// rows left in their own units), but every exchanged multiplier a thread reads from shared memory now feeds TR (column
// operands) or TC (row operands) FMAs instead of one: (TR + TC) 8-byte operands per TR TC FMAs (4 x 8: 0.375 per FMA,
// 6 x 6: 0.33) against 1.03 for the lane-owns-a-row layout.  ncu on the row layout (profiles/r02_c3s_step_ncu_details.txt)
// shows why that matters: shared-memory return bandwidth (128 B/clk/SM = 16 doubles) is a quarter of the FP64 FMA rate
// (64/clk/SM), so a kernel that fetches one operand per FMA cannot pass 25 % of the FP64 roof (L1/TEX 72 % busy,
// FP64 pipe 35 %).
//   KD = 32: 4 x 8 tiles,  32 threads per sample (one warp, __syncwarp between pivots)
//   KD = 48: 6 x 6 tiles,  64 threads per sample
//   KD = 64: 4 x 8 tiles, 128 threads per sample
// Per pivot p the threads of the column block that holds column p ("publishers", NTR consecutive threads) publish
// s_i = T[i][p] / sqrt(d_p) for their rows, in true units (column operands for everybody) and in the row's own units
// (row operands); the owner of (p+1, p+1) adds that diagonal entry, from which EVERY thread forms
// d_{p+1} = T[p+1][p+1] - s_{p+1}^2 right after the barrier, so 1/sqrt(d_{p+1}) overlaps the update: one barrier per pivot.
// Every thread tracks 1/d_i of its own rows in registers (swept rows stay in their own units, d_i x true).
// ---------------------------------------------------------------------------------------------
// Phase timing of solve_tile_kernel (development builds only: make NVCC_EXTRA=-DPPCA_SOLVE_TIMING): cycles per warp spent
// in [0] load + gather, [1] the publishers' panel section, [2] waiting at the panel barrier, [3] the rank-1 updates,
// [4] everything after the elimination; [5] = warp-samples counted.  Read with ppca_b200_debug_solve_timing (not in the ABI).
#ifdef PPCA_SOLVE_TIMING
__device__ unsigned long long g_solve_timing[8];
#define PT_DECL long long pt_last = clock64(), pt_acc[5] = {0, 0, 0, 0, 0}
#define PT_MARK(slot) { const long long pt_now = clock64(); pt_acc[slot] += pt_now - pt_last; pt_last = pt_now; }
#define PT_FLUSH if ((threadIdx.x & 31) == 0) { for (int pt_i = 0; pt_i < 5; ++pt_i) atomicAdd(&g_solve_timing[pt_i], (unsigned long long)pt_acc[pt_i]); atomicAdd(&g_solve_timing[5], 1ull); }
#else
#define PT_DECL
#define PT_MARK(slot)
#define PT_FLUSH
#endif

template <int N>
__device__ __forceinline__ void sample_sync(int id) {
  if constexpr (N <= 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

template <int KD, int TR, int TC>
struct TileLayout {
  static constexpr int NTR = KD / TR, NTC = KD / TC, TPS = NTR * NTC, SPC = 256 / TPS, WPS = TPS >= 32 ? TPS / 32 : 1;
  static constexpr int SR = TR == 2 ? 2 : 6;  // pitch (doubles) of one thread's row group in the own-units exchange: 2 x odd
                                // keeps the 128-bit reads of eight consecutive row groups on disjoint banks (TR = 2: contiguous)
  // PW pivots per exchange (a "panel"): its columns sit inside one thread tile, its rows in one thread (TR >= PW) or in
  // PW / TR consecutive threads of the column block (TR = 2)
  static constexpr int PW = (TC % 4 == 0 && (TR % 4 == 0 || 4 % TR == 0)) ? 4 : 3;
  static constexpr int PB = TR > TC ? TR : TC;          // pivots per unrolled block: static indices repeat with this period
  // one exchange buffer: PW x (true units [KD] | own units, padded [NTR SR]) | PW 1/d | pad
  static constexpr int EXQ = KD + NTR * SR;
  static constexpr int EX = PW * EXQ + 4;
  // Samples narrower than a warp (KD = 16: 8 threads, four samples per warp) synchronise with __syncwarp, which costs
  // next to nothing: ONE exchange buffer and a second sync after the reads instead of two buffers (32 samples per CTA have to fit).
  static constexpr bool SUBWARP = TPS < 32;
  static constexpr int NBUF = SUBWARP ? 1 : 2;
  // two staging buffers per sample: the next sample's packed G arrives by cp.async during this sample's elimination.  Same-box
  // A/B (tools/gpu_r2ae.sh): k = 48 solve 47.5 -> 45.0 ms, k = 64 27.15 -> 27.54 ms: on for KD = 48 only.  One array of column
  // maxima per CTA (shared-memory atomicMax) instead of one per sample slot pays for the second buffer.
  static constexpr int NSTAGE = KD == 48 ? 2 : 1;
  // per sample: stage[kkp] | exch[NBUF][EX] (reused for the partial z sums after the elimination) | yb[KD] | zb[KD] | piv[KD] | red[16]
  static constexpr int FIXED = NBUF * EX + 3 * KD + 16;
  static_assert(KD % TR == 0 && KD % TC == 0 && TR % 2 == 0 && TC % 2 == 0 && TR <= SR, "tile shape");
  static_assert((TPS % 32 == 0 || 32 % TPS == 0) && 256 % TPS == 0 && 32 % NTR == 0, "thread layout");
  static_assert((TR % PW == 0 || PW % TR == 0) && TC % PW == 0 && PB % TR == 0 && PB % TC == 0 && (PB / PW) % 2 == 0 && KD % PB == 0, "panels");
  static_assert(NBUF * EX >= NTC * KD, "the partial z sums reuse the exchange buffers");
  static size_t smem_doubles(int kkp, bool colmax) {
    return (size_t)SPC * (NSTAGE * kkp + FIXED) + (colmax ? (size_t)kkp : 0);
  }
};

// sum over the TPS lanes of one sample (TPS <= 32: inside the warp segment; every lane of the segment gets the sum)
template <int TPS>
__device__ __forceinline__ double sample_seg_sum(double v) {
#pragma unroll
  for (int o = (TPS < 32 ? TPS : 32) / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KD, int TR, int TC, int MINB>
__global__ void __launch_bounds__(256, MINB) solve_tile_kernel(SolveArgs a) {
  using LY = TileLayout<KD, TR, TC>;
  constexpr int NTR = LY::NTR, NTC = LY::NTC, TPS = LY::TPS, SPC = LY::SPC, WPS = LY::WPS, SR = LY::SR, EX = LY::EX;
  constexpr int PW = LY::PW, PB = LY::PB, EXQ = LY::EXQ;
  extern __shared__ __align__(16) double smem_reg[];
  const int smp = threadIdx.x / TPS, t = threadIdx.x % TPS;
  const int tr = t % NTR, tc = t / NTR;
  const int r0 = TR * tr, c0 = TC * tc;
  const int lane = threadIdx.x & 31, wis = t >> 5;  // warp within the sample
  const int bar_id = smp + 1;
  const int k = a.s.k, kkp = a.s.kkp, kp = a.s.kp;
  // rows / columns past k are identity padding whose pivots change nothing: whole pivot blocks past k are skipped
  const int nblk = (k + PB - 1) / PB;
  const int kpiv = nblk * PB;
  const int per_smp = LY::NSTAGE * kkp + LY::FIXED;
  double *stage0 = smem_reg + (size_t)smp * per_smp;  // NSTAGE packed rows: G in, W out
  double *exch = stage0 + LY::NSTAGE * kkp;
  double *yb = exch + LY::NBUF * EX;
  double *zb = yb + KD;
  double *piv = zb + KD;
  double *red = piv + KD;
  double *zpart = exch;  // [NTC][KD], after the elimination
  // running max |W|: one array per CTA, updated with shared-memory atomics during the write-back
  double *cmw = smem_reg + (size_t)SPC * per_smp;
  if (a.colmax) {
    for (int q = threadIdx.x; q < kkp; q += blockDim.x) cmw[q] = 0.0;
    __syncthreads();
  }
  const double sigma = a.sigma_dev ? *a.sigma_dev : a.sigma;  // device copy: the launch can be replayed from a CUDA graph
  const double s2 = sigma * sigma;
  const double ln_sigma = log(sigma);

  auto fetch = [&](int row2, double *dst) {  // asynchronous 16-byte copies of one sample's packed G (NSTAGE == 2)
    if (row2 < a.rows_pad) {
      const double *src = a.GW + (int64_t)row2 * kkp;
      for (int q = t * 2; q < kkp; q += 2 * TPS) cp_async16(dst + q, src + q, 16);
    }
    cp_async_commit();
  };
  const int rstride = gridDim.x * SPC;
  int it = 0;
  if constexpr (LY::NSTAGE == 2) fetch(blockIdx.x * SPC + smp, stage0);
  for (int row = blockIdx.x * SPC + smp; row < a.rows_pad; row += rstride, ++it) {
    PT_DECL;
    double *gsrc = a.GW + (int64_t)row * kkp;
    double *stage = stage0 + (LY::NSTAGE == 2 ? (it & 1) * kkp : 0);
    if constexpr (LY::NSTAGE == 2) {
      fetch(row + rstride, stage0 + ((it + 1) & 1) * kkp);  // its last reader (the previous sample's write-back) is behind a barrier
      cp_async_wait<1>();
    } else {
      for (int q = t * 2; q < kkp; q += 2 * TPS)
        *reinterpret_cast<double2 *>(stage + q) = *reinterpret_cast<const double2 *>(gsrc + q);
    }
    for (int r = t; r < KD; r += TPS) yb[r] = (r < kp) ? a.YZ[(int64_t)row * kp + r] : 0.0;
    const int dn = row < a.rows ? a.dn[row] : 0;
    const bool empty = dn == 0;
    const double w = (a.w && row < a.rows) ? a.w[row] : (row < a.rows ? 1.0 : 0.0);
    sample_sync<TPS>(bar_id);

    double A[TR][TC];
#pragma unroll
    for (int ia = 0; ia < TR; ++ia) {
      const int i = r0 + ia;
      const int upi = tri_row_off(i, k) - i;
#pragma unroll
      for (int jb = 0; jb < TC; ++jb) {
        const int j = c0 + jb;
        const bool use = !empty && i < k && j < k;
        int idx = (j >= i) ? upi + j : ((j * (2 * k - j - 1)) >> 1) + i;
        idx = use ? idx : 0;
        const double g = stage[idx];
        const double unit = (i == j) ? 1.0 : 0.0;
        A[ia][jb] = use ? fma(unit, s2, g) : unit;
      }
    }
    if (a.gscale && !empty) {  // precision guard (see SolveArgs): M_rr for the rows r = t, t + TPS, ...
      for (int r = t; r < k; r += TPS) {
        const int qd = tri_row_off(r, k);
        if (guard_terms((double)dn) * a.guard_coef * a.gscale[qd] > s2 + stage[qd]) atomicAdd(a.unsafe, 1u);
      }
    }

    // PANEL-BLOCKED elimination: PW pivots per exchange.  The threads that hold the panel's columns ("publishers", one
    // column block = NTR consecutive lanes) run the PW pivots of the panel among themselves — pivot values and the
    // multipliers of the panel's own rows travel by warp shuffle from the lane that holds the panel's diagonal block —
    // and publish all PW multiplier vectors at once; after ONE barrier every thread applies the PW rank-1 updates to its
    // tile.  The arithmetic per element is the pivot-by-pivot sweep's, FMA for FMA (same order), so the results are bitwise
    // those of the unblocked loop; what changes is the serial chain per sample: one (barrier, shared-memory round trip)
    // per PW pivots instead of per pivot.  The loop is unrolled over one block of PB pivots (static register indices need
    // p % TR, p % TC and the buffer parity at compile time, all periodic in PB; fully unrolled, 64 pivot bodies were 140 KB
    // of SASS against a 32 KB instruction cache).  Measured history and the ncu analysis: profiles/r02_solve_tile.md — the
    // kernel is bound by the publishers' serial chain (~400 cycles per pivot: shuffle, 1/sqrt, two multiplies, shuffle,
    // FMA) with four samples resident per SM; a look-ahead variant (next panel's chain before the publishers' own
    // remaining updates) was slower, because the publishers' warp stays the critical path and only gets more to do.
    double sc[TR];  // 1/d_i once row i has been swept (1 before): swept rows stay in their own units (d_i x true)
#pragma unroll
    for (int ia = 0; ia < TR; ++ia) sc[ia] = 1.0;
    const int gbase = lane - tr;  // first lane of this thread's column block within its warp
    const unsigned gmask = NTR == 32 ? 0xffffffffu : (((1u << NTR) - 1u) << gbase);
    for (int pb = 0; pb < nblk; ++pb) {
#pragma unroll
      for (int ps = 0; ps < PB / PW; ++ps) {
        const int pp0 = ps * PW;             // first pivot of the panel within the block (compile time after unrolling)
        const int p0 = pb * PB + pp0;
        const int cl0 = pp0 % TC;            // the panel's columns inside the tile that holds them
        // row p0 + q of the panel: register row RA(q) of the thread with tr == trp + RT(q) (TR >= PW: one thread holds them all)
#define RA(q) ((pp0 + (q)) % TR)
#define RT(q) ((pp0 % TR + (q)) / TR)
        const int trp = pb * (PB / TR) + pp0 / TR;
        double *ex = exch + ((ps & 1) % LY::NBUF) * EX;
        const bool pub = tc == pb * (PB / TC) + pp0 / TC;   // this thread holds the panel's columns
        PT_MARK(ps == 0 && pb == 0 ? 0 : 3);
        if (pub) {
          const int dlane = gbase + trp;  // lane of the thread with the panel's first diagonal entry
          double dq = __shfl_sync(gmask, A[RA(0)][cl0], dlane);
#pragma unroll
          for (int q = 0; q < PW; ++q) {
            const bool prow = tr == trp + RT(q);  // this thread holds row p0 + q
            const double dthis = dq;
            const double rinv = fast_rsqrt(dthis);
            double own[TR], tv[TR];
#pragma unroll
            for (int ia = 0; ia < TR; ++ia) own[ia] = A[ia][cl0 + q] * rinv;
            if (q + 1 < PW) {
              // next pivot d' = T[p+1][p+1] - s_{p+1}^2 straight from the lane that holds it (row p + 1 is not swept yet:
              // its own units are the true units), ahead of the panel's other updates: bitwise what the loop below leaves
              const int n = q + 1 < PW ? q + 1 : q;
              const double dn = fma(-own[RA(n)], own[RA(n)], A[RA(n)][cl0 + n]);
              dq = __shfl_sync(gmask, dn, dlane + RT(n));
            }
#pragma unroll
            for (int ia = 0; ia < TR; ++ia) tv[ia] = own[ia] * sc[ia];
#pragma unroll
            for (int ia = 0; ia < TR; ia += 2) {
              *reinterpret_cast<double2 *>(ex + q * EXQ + r0 + ia) = make_double2(tv[ia], tv[ia + 1]);
              *reinterpret_cast<double2 *>(ex + q * EXQ + KD + tr * SR + ia) = make_double2(own[ia], own[ia + 1]);
            }
            if (prow) {
              piv[p0 + q] = dthis;
              ex[PW * EXQ + q] = rinv * rinv;
            }
            // the other columns of the panel (swept or not) take this pivot's update now: their column operands are the
            // true-unit multipliers of the panel's own rows, which the diagonal lanes hold
#pragma unroll
            for (int q2 = 0; q2 < PW; ++q2) {
              if (q2 == q) continue;
              const double sq = __shfl_sync(gmask, tv[RA(q2)], dlane + RT(q2));
#pragma unroll
              for (int ia = 0; ia < TR; ++ia) {
                const double f = (prow && ia == RA(q)) ? 0.0 : own[ia];  // the pivot row is left as it is
                A[ia][cl0 + q2] = fma(-f, sq, A[ia][cl0 + q2]);
              }
            }
#pragma unroll
            for (int ia = 0; ia < TR; ++ia) A[ia][cl0 + q] = own[ia] * rinv;  // T[i][p] = s_i / sqrt(d)
            if (prow) {
              A[RA(q)][cl0 + q] = -1.0;
              sc[RA(q)] = rinv * rinv;
            }
          }
        }
        PT_MARK(1);
        sample_sync<TPS>(bar_id);
        PT_MARK(2);
#pragma unroll
        for (int q = 0; q < PW; ++q) {
          const bool prow = tr == trp + RT(q);
          double f[TR], cv[TC];
#pragma unroll
          for (int ia = 0; ia < TR; ia += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(ex + q * EXQ + KD + tr * SR + ia);
            f[ia] = v.x;
            f[ia + 1] = v.y;
          }
#pragma unroll
          for (int jb = 0; jb < TC; jb += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(ex + q * EXQ + c0 + jb);
            cv[jb] = v.x;
            cv[jb + 1] = v.y;
          }
          f[RA(q)] = prow ? 0.0 : f[RA(q)];  // the pivot row is left as it is (its 1/d is applied at the end)
#pragma unroll
          for (int jq = 0; jq < PW; ++jq) cv[cl0 + jq] = pub ? 0.0 : cv[cl0 + jq];  // the publishers' panel columns are done
#pragma unroll
          for (int ia = 0; ia < TR; ++ia)
#pragma unroll
            for (int jb = 0; jb < TC; ++jb) A[ia][jb] = fma(-f[ia], cv[jb], A[ia][jb]);
          if (prow && !pub) sc[RA(q)] = ex[PW * EXQ + q];
        }
#undef RA
#undef RT
        if constexpr (LY::NBUF == 1) sample_sync<TPS>(bar_id);  // single exchange buffer: reads done before the next panel
      }
    }

    PT_MARK(3);
    sample_sync<TPS>(bar_id);  // the partial z sums below reuse the exchange buffers

    // M^{-1}[i][j] = -(1/d_i) A[i][j]
    bool rlive[TR];
#pragma unroll
    for (int ia = 0; ia < TR; ++ia) {
      const int i = r0 + ia;
      rlive[ia] = i < k && !empty;
      const double nsc = -sc[ia];
#pragma unroll
      for (int jb = 0; jb < TC; ++jb) A[ia][jb] *= nsc;
    }
    {
      double yv[TC];
#pragma unroll
      for (int jb = 0; jb < TC; jb += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(yb + c0 + jb);
        yv[jb] = v.x;
        yv[jb + 1] = v.y;
      }
#pragma unroll
      for (int ia = 0; ia < TR; ++ia) {
        double zp = 0.0;
#pragma unroll
        for (int jb = 0; jb < TC; ++jb) zp = fma(A[ia][jb], yv[jb], zp);
        zpart[tc * KD + r0 + ia] = rlive[ia] ? zp : 0.0;
      }
    }
    sample_sync<TPS>(bar_id);
    double r_ld = 0.0, r_quad = 0.0, r_zz = 0.0, r_tp = 0.0;
    for (int r = t; r < KD; r += TPS) {  // row r of z: sum of the column blocks' partial products
      double zi = 0.0;
#pragma unroll
      for (int c = 0; c < NTC; ++c) zi += zpart[c * KD + r];
      zb[r] = zi;
      r_quad = fma(yb[r], zi, r_quad);
      r_zz = fma(zi, zi, r_zz);
      if (r < kpiv) r_ld += log(piv[r]);
    }
    if (a.mode == 2) {  // t = sigma^2 sum_i (1 - sigma^2 M^{-1}_ii): the diagonal entries this tile holds
#pragma unroll
      for (int ia = 0; ia < TR; ++ia)
#pragma unroll
        for (int jb = 0; jb < TC; ++jb)
          if (r0 + ia == c0 + jb && rlive[ia]) r_tp += fma(-s2, A[ia][jb], 1.0);
    }
    r_ld = sample_seg_sum<TPS>(r_ld);
    r_quad = sample_seg_sum<TPS>(r_quad);
    if (a.mode == 2) {
      r_zz = sample_seg_sum<TPS>(r_zz);
      r_tp = sample_seg_sum<TPS>(r_tp);
    }
    if ((t & 31) == 0) {
      red[wis * 4 + 0] = r_ld;
      red[wis * 4 + 1] = r_quad;
      red[wis * 4 + 2] = r_zz;
      red[wis * 4 + 3] = r_tp;
    }
    sample_sync<TPS>(bar_id);
    if (t == 0 && row < a.rows) {
      double ld = 0.0, quad = 0.0, zz = 0.0, tp = 0.0;
#pragma unroll
      for (int u = 0; u < WPS; ++u) {
        ld += red[u * 4 + 0];
        quad += red[u * 4 + 1];
        zz += red[u * 4 + 2];
        tp += red[u * 4 + 3];
      }
      if (a.llk) {
        double llk = 0.0;
        if (!empty)
          llk = -0.5 * (a.nx[row] - quad) / s2 - 0.5 * (ld + 2.0 * ln_sigma * (double)(dn - k)) - 0.5 * LN_2PI * (double)dn;
        a.llk[row] = llk;
      }
      if (a.mode == 2 && a.tn) a.tn[row] = empty ? 0.0 : s2 * tp;
      if (a.mode == 2 && a.dv) a.dv[row] = empty ? 0.0 : (a.nx[row] - quad) - s2 * zz;  // |R_n|^2, see SolveArgs::dv
    }
    if (a.mode != 0) {
      for (int r = t; r < kp; r += TPS) {
        const double zi = r < KD ? zb[r] : 0.0;
        a.YZ[(int64_t)row * kp + r] = zi;
        if (a.WZ) a.WZ[(int64_t)row * kp + r] = w * zi;
      }
      if (a.cov && row < a.rows) {
#pragma unroll
        for (int ia = 0; ia < TR; ++ia) {
          const int i = r0 + ia;
          if (i < k) {
            double *cvp = a.cov + (int64_t)row * k * k + (int64_t)i * k;
#pragma unroll
            for (int jb = 0; jb < TC; ++jb)
              if (c0 + jb < k) cvp[c0 + jb] = empty ? (c0 + jb == i ? 1.0 : 0.0) : s2 * A[ia][jb];
          }
        }
      }
    }
    if (a.mode == 2) {
      // all reads of G happened before the elimination barriers; W overwrites it in place (upper triangle, packed)
      double zc[TC];
#pragma unroll
      for (int jb = 0; jb < TC; jb += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(zb + c0 + jb);
        zc[jb] = v.x;
        zc[jb + 1] = v.y;
      }
      const double ws2 = empty ? 0.0 : w * s2;
#pragma unroll
      for (int ia = 0; ia < TR; ++ia) {
        const int i = r0 + ia;
        if (i < k) {
          double *so = stage + (tri_row_off(i, k) - i);
          const double wzi = empty ? 0.0 : w * zb[i];
#pragma unroll
          for (int jb = 0; jb < TC; ++jb) {
            const int j = c0 + jb;
            if (j >= i && j < k) so[j] = fma(wzi, zc[jb], ws2 * A[ia][jb]);
          }
        }
      }
      sample_sync<TPS>(bar_id);
      for (int q = t * 2; q < kkp; q += 2 * TPS) {
        const double2 v = *reinterpret_cast<const double2 *>(stage + q);
        *reinterpret_cast<double2 *>(gsrc + q) = v;
        if (a.colmax) {  // one array per CTA: non-negative doubles order like their bit patterns
          unsigned long long *cm = reinterpret_cast<unsigned long long *>(cmw + q);
          const unsigned long long bx = (unsigned long long)__double_as_longlong(fabs(v.x));
          const unsigned long long by = (unsigned long long)__double_as_longlong(fabs(v.y));
          if (bx > cm[0]) atomicMax(cm, bx);
          if (by > cm[1]) atomicMax(cm + 1, by);
        }
      }
    }
    sample_sync<TPS>(bar_id);
    PT_MARK(4);
    PT_FLUSH;
  }
  if constexpr (LY::NSTAGE == 2) cp_async_wait<0>();
  if (a.colmax) {
    __syncthreads();
    for (int c = threadIdx.x; c < kkp; c += blockDim.x) {
      const double m = cmw[c];
      if (m > 0.0) atomicMax(a.colmax + c, (unsigned long long)__double_as_longlong(m));
    }
  }
}

template <int KD, int TR, int TC, int MINB>
static void launch_solve_tile(const Launcher &L, const SolveArgs &a) {
  using LY = TileLayout<KD, TR, TC>;
  const size_t smem = LY::smem_doubles(a.s.kkp, a.colmax != nullptr) * sizeof(double);
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(solve_tile_kernel<KD, TR, TC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    110 * 1024));
  }
  REQUIRE(smem <= 110 * 1024, "solve_tile: shared memory %zu", smem);
  int64_t blocks = (a.rows_pad + LY::SPC - 1) / LY::SPC;
  if (blocks > MINB * (int64_t)L.sms) blocks = MINB * (int64_t)L.sms;
  solve_tile_kernel<KD, TR, TC, MINB><<<(unsigned)blocks, 256, smem, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(V_SOLVE_TILE);
}

template <int KP, int MINB, bool PF>
static void launch_solve_reg_pf(const Launcher &L, const SolveArgs &a) {
  constexpr int SPW = 32 / KP;
  const int warps = 8;
  const size_t smem = (size_t)warps * ((PF ? 2 : 1) * SPW * a.s.kkp + 128 + (a.colmax ? SPW * a.s.kkp : 0)) * sizeof(double);
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(solve_reg_kernel<KP, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  }
  const int groups = (a.rows_pad + SPW - 1) / SPW;
  int64_t blocks = (groups + warps - 1) / warps;
  // exactly one resident wave (the kernel's __launch_bounds__ occupancy): the grid-stride loop over sample groups
  // balances to < 1 %, where a 2.67-wave grid left the last third of the SMs idle for a whole CTA lifetime
  const int64_t cap = (int64_t)L.sms * MINB;
  if (blocks > cap) blocks = cap;
  solve_reg_kernel<KP, MINB, PF><<<(unsigned)blocks, warps * 32, smem, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(KP == 8 ? V_SOLVE_REG8 : KP == 16 ? V_SOLVE_REG16 : V_SOLVE_REG32);
}

template <int KP, int MINB>
static void launch_solve_reg_b(const Launcher &L, const SolveArgs &a) {
  // measured (tools/gpu_r2ac.sh): k = 32 (c4) 45.8 -> 43.0 ms with the prefetch, k = 16 (c2) 1.53 -> 1.63 ms: on for KP = 32 only
  static const char *env = getenv("PPCA_B200_SOLVE_PREFETCH");
  const bool prefetch = env ? strcmp(env, "0") != 0 : KP == 32;
  if (prefetch) launch_solve_reg_pf<KP, MINB, true>(L, a);
  else launch_solve_reg_pf<KP, MINB, false>(L, a);
}

template <int KP>
static void launch_solve_reg(const Launcher &L, const SolveArgs &a) {
  // resident CTAs per SM the kernel is compiled for (register budget 65536 / (256 MINB)); PPCA_B200_SOLVE_MINB
  // overrides the default for k <= 16 (experiments)
  static const int minb16 = getenv("PPCA_B200_SOLVE_MINB") ? atoi(getenv("PPCA_B200_SOLVE_MINB")) : 3;
  if constexpr (KP == 32) launch_solve_reg_b<32, 2>(L, a);
  else if constexpr (KP == 16) {
    if (minb16 == 2) launch_solve_reg_b<16, 2>(L, a);
    else if (minb16 == 4) launch_solve_reg_b<16, 4>(L, a);
    else launch_solve_reg_b<16, 3>(L, a);
  } else launch_solve_reg_b<8, 4>(L, a);
}

// Per-sample scalars of one chunk -> SOLVE_SLOTS partial slots (block b always owns slot b and rows
// [b R, (b+1) R) of the chunk, so the summation order is fixed run to run); launch_solve_finish sums the slots.
// Per-sample scalars of one chunk -> SOLVE_SLOTS partial slots (block b always owns slot b and rows [b R, (b+1) R) of the
// chunk, so the summation order is fixed run to run); launch_solve_finish sums the slots.
// The residual norms dv come from the identity |R_n|^2 = nx - y^T z - sigma^2 |z|^2, which subtracts numbers of size nx:
// its absolute error is ~2 eps nx.  Every block also sums w nx and w (t + dv); the LAST block to finish adds the block
// totals in fixed order, raises *resid_flag when 2 eps sum(w nx) could move the noise update by more than 1e-11 (the chunk
// then takes its residual norms from resid_exact_kernel, moments.cu), and only then folds the block totals into `part`,
// with dv left out when flagged.
__global__ void __launch_bounds__(256) solve_reduce_kernel(int rows, const double *__restrict__ llk,
                                                           const double *__restrict__ tn, const double *__restrict__ dv,
                                                           const double *__restrict__ nx, const int *__restrict__ dn,
                                                           const double *__restrict__ w, double *part, double *scratch,
                                                           int *resid_flag) {
  __shared__ double sh[6][8];
  __shared__ int last_flag[2];
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(rows, lo + per);
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // w t, w llk, w, #non-empty, w dv, w nx
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const double wi = w ? w[i] : 1.0;
    if (tn) v[0] = fma(wi, tn[i], v[0]);
    if (llk) v[1] = fma(wi, llk[i], v[1]);
    v[2] += wi;
    v[3] += dn[i] > 0 ? 1.0 : 0.0;
    if (dv) {
      v[4] = fma(wi, dv[i], v[4]);
      v[5] = fma(wi, nx[i], v[5]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j] = warp_sum(v[j]);
    if (lane == 0) sh[j][wid] = v[j];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sh[threadIdx.x][i];
    scratch[blockIdx.x * 6 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  unsigned int *counter = reinterpret_cast<unsigned int *>(scratch + (size_t)gridDim.x * 6);
  if (threadIdx.x == 0) last_flag[0] = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_flag[0]) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double e = 0.0, s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) {
      e += scratch[b * 6 + 5];
      s += scratch[b * 6 + 0] + scratch[b * 6 + 4];
    }
    e = warp_sum(e);
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      const int flag = (dv != nullptr && 4.5e-16 * e > 1e-11 * s) ? 1 : 0;
      last_flag[1] = flag;
      if (resid_flag) *resid_flag = flag;
      *counter = 0u;  // ready for the next launch
    }
  }
  __syncthreads();
  const bool use_dv = dv != nullptr && !last_flag[1];
  for (int idx = threadIdx.x; idx < (int)gridDim.x * 4; idx += 256) {
    const int b = idx >> 2, j = idx & 3;
    part[idx] += scratch[b * 6 + j] + ((j == 0 && use_dv) ? scratch[b * 6 + 4] : 0.0);
  }
}

__global__ void solve_finish_kernel(const double *__restrict__ part, double *scalars) {
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int b = 0; b < SOLVE_SLOTS; ++b) s += part[b * 4 + threadIdx.x];
    const int j = threadIdx.x;
    const int slot = j == 0 ? SC_SQERR : j == 1 ? SC_LLK : j == 2 ? SC_SUMW : SC_NONEMPTY;
    scalars[slot] += s;
  }
}

void launch_solve_reduce(const Launcher &L, int rows, const double *llk, const double *tn, const double *dv,
                         const double *nx, const int *dn, const double *w, double *part, double *scratch, int *resid_flag) {
  if (rows <= 0) return;
  solve_reduce_kernel<<<SOLVE_SLOTS, 256, 0, L.stream>>>(rows, llk, tn, dv, nx, dn, w, part, scratch, resid_flag);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

void launch_solve_finish(const Launcher &L, const double *part, double *scalars) {
  solve_finish_kernel<<<1, 32, 0, L.stream>>>(part, scalars);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
}

static void launch_solve_generic(const Launcher &L, const SolveArgs &a) {
  const SolveSmem lay(a.s.k);
  const size_t per_warp = (size_t)lay.per_warp * sizeof(double);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  REQUIRE(per_warp * warps <= 227 * 1024, "state_size %d too large for the per-sample solve kernel", a.s.k);
  const size_t smem = per_warp * warps;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm * warps > 48) ctas_per_sm = 48 / warps > 0 ? 48 / warps : 1;
  int64_t blocks = (a.rows_pad + warps - 1) / warps;
  const int64_t cap = (int64_t)L.sms * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  solve_kernel<<<(unsigned)blocks, warps * 32, smem, L.stream>>>(a);
  CUDA_CHECK(cudaGetLastError());
  ++*L.launch_counter;
  L.count(V_SOLVE_GENERIC);
}

static bool tile32_fits(const SolveArgs &a) {  // eight samples per CTA: the widest states with column maxima do not fit
  using LY = TileLayout<32, 4, 8>;
  return LY::smem_doubles(a.s.kkp, a.colmax != nullptr) * sizeof(double) <= 110 * 1024;
}

#ifdef PPCA_SOLVE_TIMING
}  // namespace ppca
extern "C" __attribute__((visibility("default"))) int ppca_b200_debug_solve_timing(unsigned long long *out8, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out8, ppca::g_solve_timing, sizeof(unsigned long long) * 8) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(ppca::g_solve_timing, z, sizeof(z));
  }
  return 0;
}
namespace ppca {
#endif

void launch_solve(const Launcher &L, const SolveArgs &a) {
  if (a.rows_pad <= 0) return;
  REQUIRE(a.mode == 0 || a.GW != nullptr, "solve: missing Gram buffer");
  // Removed in round 2: a DMMA-blocked 8 x 8 Gauss-Jordan kernel and a 64-lane variant (the blocked elimination was not
  // symmetric-consistent — tools/solve_accuracy.py: four digits worse on ill-conditioned M_n — and never beat the scalar
  // kernels, profiles/r01_solve_blk_ncu.txt) and the 128-thread row kernel for 32 < k <= 64 (superseded by the tiled one).
  // k <= 32: lane-owns-a-row kernels; 32 < k <= 64: register-tiled, panel-blocked sweep (measured against the round-1/2
  // 128-thread row kernel it replaced: k = 48 -22 %, k = 64 -11 %; profiles/r02_solve_tile.md).  The tiled kernel also runs
  // k <= 32 (PPCA_B200_SOLVE=tile: k = 32 +16 %) and k <= 16 (PPCA_B200_SOLVE16=tile: +36 %): slower there, kept as switches.
  static const bool tile32 = getenv("PPCA_B200_SOLVE") && !strcmp(getenv("PPCA_B200_SOLVE"), "tile");
  static const bool tile16 = getenv("PPCA_B200_SOLVE16") && !strcmp(getenv("PPCA_B200_SOLVE16"), "tile");
  if (a.s.k <= 8) launch_solve_reg<8>(L, a);
  else if (a.s.k <= 16 && tile16) launch_solve_tile<16, 4, 8, 2>(L, a);
  else if (a.s.k <= 16) launch_solve_reg<16>(L, a);
  else if (a.s.k <= 32 && tile32 && tile32_fits(a)) launch_solve_tile<32, 4, 8, 2>(L, a);
  else if (a.s.k <= 32) launch_solve_reg<32>(L, a);
  else if (a.s.k <= 48) launch_solve_tile<48, 6, 6, 2>(L, a);
  else if (a.s.k <= 64) {
    static const bool wide = getenv("PPCA_B200_SOLVE64") && !strcmp(getenv("PPCA_B200_SOLVE64"), "2x16");
    if (wide) launch_solve_tile<64, 2, 16, 2>(L, a);
    else launch_solve_tile<64, 4, 8, 2>(L, a);
  }
  else {
    REQUIRE(a.colmax == nullptr, "solve: the generic kernel (state_size > 64) does not produce column maxima");
    launch_solve_generic(L, a);
  }
  if (a.part) launch_solve_reduce(L, a.rows, a.llk, a.tn, nullptr, nullptr, a.dn, a.w, a.part, a.rscratch, nullptr);
}

}  // namespace ppca
