// k2_suffstats.cuh -- K2: weighted sufficient statistics of the proposal update, float64, sm_100a.
//
// Replaces (reference loops, /root/reference/pypmc):
//   gaussian_pmc   einsum('n,nk->k'), einsum('n,nk,ni->ki'), per-k einsum('n,n,ni,nj->ij')   mix_adapt/pmc.pyx:191-222
//   student_t_pmc  the same with gamma_nk                                                 mix_adapt/pmc.pyx:612-650
//   GaussianInference._update_N_comp / _update_x_mean_comp / _update_S [+ _weighted]      mix_adapt/variational.pyx:699-709, 806-932
//
// For every component k it accumulates, over the samples n of this rank,
//     A_k = sum_n w_n rho_nk                      B_k = sum_n v_nk          (v_nk = w_n rho_nk gamma_nk)
//     m_k = sum_n v_nk y_n                        R_k = sum_n v_nk y_n y_n^T  (lower triangle)
// with y_n = x_n - shift (one shift vector for all components, chosen by the host near the bulk of the
// mixture so that the raw moments do not cancel badly).  The host turns (A, B, m, R) into the reference's
// two-pass quantities: delta = m/B, mean = shift + delta, cov = (R - B delta delta^T) / A.
//
// Shape of the computation: Out[k, f] = sum_n V[n, k] * Phi[n, f], with the feature vector
// Phi_n = [1, y_n, tril(y_n y_n^T)] of length F = 1 + D + D(D+1)/2 -- a (K x N) x (N x F) product whose
// B operand is built on the fly in shared memory (D(D+1)/2 multiplies per sample, shared by all K
// components).  K F FP64 FMAs per sample against 8 (D + K) bytes: DFMA-pipe bound like K1.
//
// Mapping: a warp owns a 16 (components) x 128 (features) block of Out, each lane a 16 x 4 register tile
// (64 accumulators).  Per sample a lane issues 8 broadcast LDS.128 (16 v's) + 2 LDS.128 (4 phi's) for 64
// DFMAs.  CTAs are persistent over sample tiles; each CTA writes one partial Out block, and a second
// tiny kernel adds the partials in CTA order -- no floating-point atomics, so results are reproducible
// run to run for a given grid.
#pragma once

#include "pmc_common.cuh"

namespace pmc {

constexpr int K2_TK = 16;    // components per warp block
constexpr int K2_TF = 128;   // features per warp block
constexpr int K2_MAX_WARPS = 8;

struct StatsArgs {
  const double* x;      // [n, ldx]
  int64_t n;
  int64_t ldx;
  int d;
  const double* shift;  // [d]
  const double* rho;    // [n, ld_rho]
  const double* gamma;  // [n, ld_rho] or null
  const double* sw;     // [n] or null
  int k;                // number of components (columns used)
  int ld_rho;
  int F;                // 1 + d + d(d+1)/2
  int units_k, units_f; // ceil(k/16), ceil(F/128)
  int wk, wf;           // warp grid of one CTA (wk * wf warps)
  int tn;               // samples per tile
  double* partial;      // [gridDim.x, k, F+2]   (column 0 = A, column 1+f = Out[k,f], column F+1 = sum w rho ln gamma)
};

__global__ void __launch_bounds__(K2_MAX_WARPS * 32, 1) k2_suffstats(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nwarps = a.wk * a.wf;
  const int nthreads = nwarps * 32;
  const int KC = a.wk * K2_TK;          // components covered by this CTA (padded)
  const int FC = a.wf * K2_TF;          // features covered by this CTA (padded)
  const int chunk_k = blockIdx.y % ((a.units_k + a.wk - 1) / a.wk);
  const int chunk_f = blockIdx.y / ((a.units_k + a.wk - 1) / a.wk);
  const int k0 = chunk_k * KC, f0 = chunk_f * FC;
  const int D = a.d, TN = a.tn;

  double* ytile = reinterpret_cast<double*>(smem_raw);     // [TN][D]
  double* vtile = ytile + ((TN * D + 1) & ~1);             // [TN][KC]   w rho gamma
  double* atile = vtile + TN * KC;                         // [TN][KC]   w rho        (only if gamma)
  double* ltile = atile + (a.gamma ? TN * KC : 0);         // [TN][KC]   w rho ln(gamma) (only if gamma)
  double* phi = ltile + (a.gamma ? TN * KC : 0);           // [TN][FC]
  short2* fmap = reinterpret_cast<short2*>(phi + size_t(TN) * FC);  // [FC] feature -> (i, j)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wkid = warp % a.wk, wfid = warp / a.wk;

  // feature map: f = 0 -> (-1,-1) [constant 1];  1..D -> (i,-1) [y_i];  then (i,j), j<=i;  beyond F -> (-2,-2) [0]
  for (int fl = tid; fl < FC; fl += nthreads) {
    const int f = f0 + fl;
    short2 ij;
    if (f == 0) ij = make_short2(-1, -1);
    else if (f <= D) ij = make_short2(short(f - 1), -1);
    else if (f < a.F) {
      const int t = f - 1 - D;
      int i = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      while (i * (i + 1) / 2 > t) --i;
      ij = make_short2(short(i), short(t - i * (i + 1) / 2));
    } else ij = make_short2(-2, -2);
    fmap[fl] = ij;
  }

  double acc[K2_TK][4];
#pragma unroll
  for (int i = 0; i < K2_TK; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double acc_a = 0.0, acc_l = 0.0;  // A_k and sum w rho ln(gamma) for thread tid < KC (chunk_f == 0, gamma != null)

  const int64_t num_tiles = (a.n + TN - 1) / TN;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TN;
    __syncthreads();  // previous tile fully consumed
    // ---- y tile (coalesced over the contiguous rows) ----
    for (int e = tid; e < TN * D; e += nthreads) {
      const int r = e / D, j = e - r * D;
      const int64_t row = row0 + r;
      ytile[e] = (row < a.n) ? (__ldg(a.x + row * a.ldx + j) - __ldg(a.shift + j)) : 0.0;
    }
    // ---- v tile ----
    for (int e = tid; e < TN * KC; e += nthreads) {
      const int r = e / KC, kk = e - r * KC;
      const int64_t row = row0 + r;
      const int k = k0 + kk;
      double v = 0.0, va = 0.0, vl = 0.0;
      if (row < a.n && k < a.k) {
        va = __ldg(a.rho + row * a.ld_rho + k);
        if (a.sw) va *= __ldg(a.sw + row);
        v = va;
        if (a.gamma) {
          const double g = __ldg(a.gamma + row * a.ld_rho + k);
          v = va * g;
          vl = (va != 0.0) ? va * log(g) : 0.0;   // feeds the dof condition, pmc.pyx:672-679
        }
      }
      vtile[e] = v;
      if (a.gamma) { atile[e] = va; ltile[e] = vl; }
    }
    __syncthreads();
    // ---- feature tile ----
    for (int e = tid; e < TN * FC; e += nthreads) {
      const int r = e / FC, fl = e - r * FC;
      const short2 ij = fmap[fl];
      double p;
      if (ij.x == -1) p = 1.0;
      else if (ij.x == -2) p = 0.0;
      else if (ij.y == -1) p = ytile[r * D + ij.x];
      else p = ytile[r * D + ij.x] * ytile[r * D + ij.y];
      phi[e] = p;
    }
    __syncthreads();
    // ---- rank-TN update of the register tiles ----
    const double* vp = vtile + wkid * K2_TK;
    const double* pp = phi + wfid * K2_TF + 2 * lane;
#pragma unroll 2
    for (int r = 0; r < TN; ++r) {
      const double2 p0 = *reinterpret_cast<const double2*>(pp + size_t(r) * FC);
      const double2 p1 = *reinterpret_cast<const double2*>(pp + size_t(r) * FC + 64);
#pragma unroll
      for (int kk = 0; kk < K2_TK / 2; ++kk) {
        const double2 v = *reinterpret_cast<const double2*>(vp + r * KC + 2 * kk);
        acc[2 * kk][0] = fma(v.x, p0.x, acc[2 * kk][0]);
        acc[2 * kk][1] = fma(v.x, p0.y, acc[2 * kk][1]);
        acc[2 * kk][2] = fma(v.x, p1.x, acc[2 * kk][2]);
        acc[2 * kk][3] = fma(v.x, p1.y, acc[2 * kk][3]);
        acc[2 * kk + 1][0] = fma(v.y, p0.x, acc[2 * kk + 1][0]);
        acc[2 * kk + 1][1] = fma(v.y, p0.y, acc[2 * kk + 1][1]);
        acc[2 * kk + 1][2] = fma(v.y, p1.x, acc[2 * kk + 1][2]);
        acc[2 * kk + 1][3] = fma(v.y, p1.y, acc[2 * kk + 1][3]);
      }
    }
    if (a.gamma && chunk_f == 0 && tid < KC) {
      for (int r = 0; r < TN; ++r) { acc_a += atile[r * KC + tid]; acc_l += ltile[r * KC + tid]; }
    }
  }

  // ---- write this CTA's partial block ----
  const int ldp = a.F + 2;
  double* out = a.partial + size_t(blockIdx.x) * a.k * ldp;
#pragma unroll
  for (int i = 0; i < K2_TK; ++i) {
    const int k = k0 + wkid * K2_TK + i;
    if (k >= a.k) continue;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int f = f0 + wfid * K2_TF + 64 * c + 2 * lane + e;
        if (f < a.F) {
          out[size_t(k) * ldp + 1 + f] = acc[i][2 * c + e];
          if (f == 0 && !a.gamma) out[size_t(k) * ldp] = acc[i][2 * c + e];  // A == B without gamma
        }
      }
  }
  if (chunk_f == 0 && tid < KC && k0 + tid < a.k) {
    if (a.gamma) out[size_t(k0 + tid) * ldp] = acc_a;
    out[size_t(k0 + tid) * ldp + a.F + 1] = acc_l;
  }
}

// out[e] = sum_b partial[b][e], b ascending (fixed order)
__global__ void k2_reduce_partials(const double* __restrict__ partial, int nblocks, int64_t len, double* __restrict__ out) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= len) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[size_t(b) * len + e];
  out[e] = s;
}

}  // namespace pmc
