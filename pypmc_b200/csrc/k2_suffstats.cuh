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
// Shape of the computation: K symmetric rank-N updates sharing one y (plus the linear terms sum v y, sum v):
// (D+1)(D+2)/2 FP64 FMAs per sample-component against 8 (D + K) bytes per sample: DFMA-pipe bound like K1.
//
// Mapping (round 1 profile: the first K2 staged a [samples x features] product table in shared memory behind
// three block-wide barriers per 44-sample tile and reached 18 % of the DFMA peak): a thread owns, for 16
// components (64 accumulators), either one 2x2 block (row pair r, column pair p <= r) of the lower triangle of
// y y^T, or four consecutive entries of the linear row [y, 1] (m_k and B_k).  Per sample it reads the 16 v's
// (8 broadcast LDS.128) and two pairs of the sample row (2 LDS.128), forms its 4 features in registers (4 DMUL
// + selects) and issues 64 DFMAs; no product table exists anywhere.  Lane tiles are numbered
// t = kgroup * (Bq + Lq) + tile (Bq = P0(P0+1)/2 blocks, P0 = ceil(D/2); Lq = ceil((D+1)/4) quads) and dealt to
// 8 warps per CTA -- two per SM sub-partition, the only warp count that both leaves 255 registers per thread
// and loads the four schedulers evenly; K=32, D=30 gives exactly 256 lane tiles.  If there are more,
// gridDim.y CTAs share the same samples.  Samples stream through a two-stage shared-memory pipeline:
// raw rho / gamma / x / w rows arrive with cp.async (LDGSTS) one tile ahead, a short in-place pass turns them
// into v and yh, and the block-wide barriers are per 128-sample tile (~37k clk of DFMA work).
// Each CTA writes one partial block; a second tiny kernel adds the partials in CTA order -- no floating-point
// atomics, so results are reproducible run to run for a given grid.
#pragma once

#include "pmc_common.cuh"

namespace pmc {

constexpr int K2_TK = 16;          // components per lane tile
constexpr int K2_MAX_THREADS = 256;  // 8 warps, two per SM sub-partition

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
  int P0;               // ceil(d/2): pairs of y
  int Bq;               // P0(P0+1)/2 quadratic blocks per component group
  int Lq;               // ceil((d+1)/4) quads of the linear row [y, 1]
  int DP4;              // d+1 rounded up to a multiple of 4: row length of the staged samples
  int KP;               // k rounded up to a multiple of 16
  int LT;               // lane tiles = (KP/16) * (Bq + Lq)
  int tn;               // samples per tile (multiple of 2)
  double* partial;      // [gridDim.x, k, F+2]   (column 0 = A, column 1+f = Out[k,f], column F+1 = sum w rho ln gamma)
};

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(K2_MAX_THREADS, 1) k2_suffstats(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int D = a.d, KP = a.KP, DP2 = a.DP4, TN = a.tn;
  const bool has_g = a.gamma != nullptr;

  // stage layout (doubles): V [TN][KP] | G [TN][KP] (gamma only) | Y [TN][DP2] | W [TN]
  const int stage_len = TN * KP * (has_g ? 2 : 1) + TN * DP2 + TN;
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  double* shift_s = stage0 + 2 * stage_len;                       // [DP2]
  double* colsum = shift_s + DP2;                                 // gamma only: [nwarps][2][KP] column sums of w rho, w rho ln(gamma)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  for (int j = tid; j < DP2; j += nthreads) shift_s[j] = (j < D) ? a.shift[j] : 0.0;
  if (has_g)
    for (int e = tid; e < nwarps * 2 * KP; e += nthreads) colsum[e] = 0.0;

  // ---- this thread's lane tile ----
  const int t = blockIdx.y * nthreads + tid;
  const bool active = t < a.LT;
  const int tt = active ? t : 0;
  const int per_group = a.Bq + a.Lq;
  const int kg = tt / per_group, b = tt - kg * per_group;
  const bool lin = b >= a.Bq;                     // linear tile: entries 4q .. 4q+3 of [y, 1]
  int r = 0, p = 0;
  if (!lin) {
    r = int((sqrt(8.0 * b + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= b) ++r;
    while (r * (r + 1) / 2 > b) --r;
    p = b - r * (r + 1) / 2;
  }
  const int off_a = lin ? 4 * (b - a.Bq) + 2 : 2 * r;    // second pair (linear) / row pair (quadratic)
  const int off_b = lin ? 4 * (b - a.Bq) : 2 * p;        // first pair (linear) / column pair (quadratic)

  double acc[K2_TK][4];
#pragma unroll
  for (int i = 0; i < K2_TK; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  const int64_t num_tiles = (a.n + TN - 1) / TN;

  auto issue_loads = [&](int64_t tile, double* st) {
    const int64_t row0 = tile * TN;
    const int rows = int((a.n - row0 < TN) ? (a.n - row0) : TN);
    double* Vs = st;
    double* Gs = st + TN * KP;
    double* Ys = st + TN * KP * (has_g ? 2 : 1);
    double* Ws = Ys + TN * DP2;
    for (int rr = warp; rr < rows; rr += nwarps) {          // one warp per sample row, lanes over columns
      const double* rp = a.rho + (row0 + rr) * a.ld_rho;
      for (int kk = lane; kk < a.k; kk += 32) cp_async8(Vs + rr * KP + kk, rp + kk);
      if (has_g) {
        const double* gp = a.gamma + (row0 + rr) * a.ld_rho;
        for (int kk = lane; kk < a.k; kk += 32) cp_async8(Gs + rr * KP + kk, gp + kk);
      }
      const double* xp = a.x + (row0 + rr) * a.ldx;
      for (int jj = lane; jj < D; jj += 32) cp_async8(Ys + rr * DP2 + jj, xp + jj);
      if (a.sw && lane == 0) cp_async8(Ws + rr, a.sw + row0 + rr);
    }
  };

  int s = 0;
  if (int64_t(blockIdx.x) < num_tiles) issue_loads(blockIdx.x, stage0);
  cp_async_commit();

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, s ^= 1) {
    double* st = stage0 + s * stage_len;
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles) issue_loads(next, stage0 + (s ^ 1) * stage_len);
    cp_async_commit();
    cp_async_wait<1>();            // everything but the group just committed has landed
    __syncthreads();

    const int64_t row0 = tile * TN;
    const int rows = int((a.n - row0 < TN) ? (a.n - row0) : TN);
    double* Vs = st;
    double* Gs = st + TN * KP;
    double* Ys = st + TN * KP * (has_g ? 2 : 1);
    double* Ws = Ys + TN * DP2;

    // ---- in-place transform: v = w rho gamma (zero outside the data), yh = [x - shift, 1, 0] ----
    double* cs_a = colsum + warp * 2 * KP;                  // this warp's column sums (gamma only)
    for (int rr = warp; rr < TN; rr += nwarps) {
      const bool in = rr < rows;
      const double w = (in && a.sw) ? Ws[rr] : 1.0;
      for (int kk = lane; kk < KP; kk += 32) {
        double v = 0.0;
        if (in && kk < a.k) {
          v = Vs[rr * KP + kk] * w;
          if (has_g) {
            const double g = Gs[rr * KP + kk];
            cs_a[kk] += v;                                    // same thread every time: ordered, no race
            cs_a[KP + kk] += (v != 0.0) ? v * log(g) : 0.0;   // feeds the dof condition, pmc.pyx:672-679
            v *= g;
          }
        }
        Vs[rr * KP + kk] = v;
      }
      for (int jj = lane; jj < DP2; jj += 32) {
        double y = 0.0;
        if (in) y = (jj < D) ? (Ys[rr * DP2 + jj] - shift_s[jj]) : ((jj == D) ? 1.0 : 0.0);
        Ys[rr * DP2 + jj] = y;
      }
    }
    __syncthreads();

    // ---- rank-TN update of the register tile ----
    const double* vp = Vs + kg * K2_TK;
    const double* yr = Ys + off_a;
    const double* yp = Ys + off_b;
#pragma unroll 2
    for (int n = 0; n < TN; ++n) {
      const double2 ya = *reinterpret_cast<const double2*>(yr + n * DP2);
      const double2 yb = *reinterpret_cast<const double2*>(yp + n * DP2);
      const double f0 = lin ? yb.x : ya.x * yb.x, f1 = lin ? yb.y : ya.x * yb.y;
      const double f2 = lin ? ya.x : ya.y * yb.x, f3 = lin ? ya.y : ya.y * yb.y;
#pragma unroll
      for (int c = 0; c < K2_TK / 2; ++c) {
        const double2 v = *reinterpret_cast<const double2*>(vp + n * KP + 2 * c);
        acc[2 * c][0] = fma(v.x, f0, acc[2 * c][0]);
        acc[2 * c][1] = fma(v.x, f1, acc[2 * c][1]);
        acc[2 * c][2] = fma(v.x, f2, acc[2 * c][2]);
        acc[2 * c][3] = fma(v.x, f3, acc[2 * c][3]);
        acc[2 * c + 1][0] = fma(v.y, f0, acc[2 * c + 1][0]);
        acc[2 * c + 1][1] = fma(v.y, f1, acc[2 * c + 1][1]);
        acc[2 * c + 1][2] = fma(v.y, f2, acc[2 * c + 1][2]);
        acc[2 * c + 1][3] = fma(v.y, f3, acc[2 * c + 1][3]);
      }
    }
    __syncthreads();   // the stage may be refilled by the loads issued at the top of the next iteration
  }
  cp_async_wait<0>();

  // ---- write this CTA's partial block ----
  const int ldp = a.F + 2;
  double* out = a.partial + size_t(blockIdx.x) * a.k * ldp;
  if (active) {
#pragma unroll
    for (int c = 0; c < K2_TK; ++c) {
      const int k = kg * K2_TK + c;
      if (k >= a.k) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int f;
        if (lin) {
          const int j = off_b + q;                    // entry of [y, 1]
          if (j > D) continue;                        // padding
          f = (j == D) ? 0 : 1 + j;                   // B_k / m_k
        } else {
          const int i = 2 * r + (q >> 1), j = 2 * p + (q & 1);
          if (j > i || i >= D) continue;              // duplicate above the diagonal / padding (odd D)
          f = 1 + D + i * (i + 1) / 2 + j;            // second moments, lower triangle row-major
        }
        out[size_t(k) * ldp + 1 + f] = acc[c][q];
        if (f == 0 && !has_g) out[size_t(k) * ldp] = acc[c][q];   // A == B without gamma
      }
    }
  }
  if (blockIdx.y == 0) {
    if (has_g) {   // column sums: add the per-warp slots in warp order
      __syncthreads();
      for (int k = tid; k < a.k; k += nthreads) {
        double sa = 0.0, sl = 0.0;
        for (int wv = 0; wv < nwarps; ++wv) { sa += colsum[wv * 2 * KP + k]; sl += colsum[wv * 2 * KP + KP + k]; }
        out[size_t(k) * ldp] = sa;
        out[size_t(k) * ldp + a.F + 1] = sl;
      }
    } else {
      for (int k = tid; k < a.k; k += nthreads) out[size_t(k) * ldp + a.F + 1] = 0.0;
    }
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
