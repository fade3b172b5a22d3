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
// gridDim.y CTAs share the same samples.  The CTA is warp-specialised: a third warpgroup (setmaxnreg.dec to 88
// registers, the consumers setmaxnreg.inc to 208) reads rho / gamma / x / w rows from global memory, forms
// v = w rho gamma and [x - shift, 1] and fills a three-stage shared-memory ring; stages are handed over with
// mbarriers (full / empty), so the consumers never wait on global memory and there is no block-wide barrier
// in the sample loop.
// Each CTA writes one partial block; a second tiny kernel adds the partials in CTA order -- no floating-point
// atomics, so results are reproducible run to run for a given grid.
#pragma once

#include "pmc_common.cuh"

namespace pmc {

constexpr int K2_TK = 16;          // components per lane tile

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

constexpr int K2_CONSUMERS = 256;       // 8 consumer warps (two warpgroups)
constexpr int K2_PRODUCERS = 128;       // 1 producer warpgroup
constexpr int K2_THREADS = K2_CONSUMERS + K2_PRODUCERS;
constexpr int K2_STAGES = 3;
constexpr int K2_PR = 4;                // rows a producer warp keeps in flight
// setmaxnreg budget: the CTA owns 384 x 168 = 64512 registers (launch bound); 256 x 208 + 128 x 88 = 64512.
static_assert(K2_CONSUMERS * 208 + K2_PRODUCERS * 88 <= K2_THREADS * 168, "setmaxnreg.inc would wait forever");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Stage layout (doubles): V [TN][KP] | Y [TN][DP4].  The producer warpgroup (88 registers) reads rho / gamma /
// x / w rows from global memory, forms v = w rho gamma and yh = [x - shift, 1, 0...] and stores them; the two
// consumer warpgroups (208 registers, setmaxnreg) only ever touch shared memory and the FP64 pipe.
__global__ void __launch_bounds__(K2_THREADS, 1) k2_suffstats(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int D = a.d, KP = a.KP, DP4 = a.DP4, TN = a.tn;
  const bool has_g = a.gamma != nullptr;
  const int stage_len = TN * (KP + DP4);
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  double* shift_s = stage0 + K2_STAGES * stage_len;               // [DP4]
  double* colsum = shift_s + DP4;                                 // gamma only: [4 producer warps][2][KP]
  uint64_t* full = reinterpret_cast<uint64_t*>(colsum + (has_g ? 4 * 2 * KP : 0));
  uint64_t* empty = full + K2_STAGES;

  for (int j = tid; j < DP4; j += K2_THREADS) shift_s[j] = (j < D) ? a.shift[j] : 0.0;
  if (has_g)
    for (int e = tid; e < 4 * 2 * KP; e += K2_THREADS) colsum[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < K2_STAGES; ++s) {
      mbar_init(&full[s], K2_PRODUCERS);
      mbar_init(&empty[s], K2_CONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int64_t num_tiles = (a.n + TN - 1) / TN;
  const int ldp = a.F + 2;
  double* out = a.partial + size_t(blockIdx.x) * a.k * ldp;

  if (tid >= K2_CONSUMERS) {
    // =============================== producer warpgroup ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const int pw = (tid - K2_CONSUMERS) >> 5, lane = tid & 31;
    double* cs_a = colsum + pw * 2 * KP;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it % K2_STAGES;
      mbar_wait(&empty[s], uint32_t(((it / K2_STAGES) & 1) ^ 1));
      double* Vs = stage0 + s * stage_len;
      double* Ys = Vs + TN * KP;
      const int64_t row0 = tile * TN;
      const int rows = int((a.n - row0 < TN) ? (a.n - row0) : TN);
      if (KP <= 64 && DP4 <= 64) {
        // common sizes: every global load of a 4-row step is issued before the first use (about 20 in flight per
        // thread), so a step costs one memory latency instead of one per column block
        for (int rb = pw * K2_PR; rb < TN; rb += 4 * K2_PR) {
          double w[K2_PR], rv[K2_PR][2], gv[K2_PR][2], xv[K2_PR][2];
#pragma unroll
          for (int u = 0; u < K2_PR; ++u) {
            const bool rin = rb + u < rows;
            const int64_t row = row0 + rb + u;
            w[u] = (a.sw && rin) ? __ldg(a.sw + row) : 1.0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int kk = lane + 32 * c;
              const bool in = rin && kk < a.k;
              rv[u][c] = in ? __ldg(a.rho + row * a.ld_rho + kk) : 0.0;
              gv[u][c] = (in && has_g) ? __ldg(a.gamma + row * a.ld_rho + kk) : 1.0;
              xv[u][c] = (rin && kk < D) ? __ldg(a.x + row * a.ldx + kk) : 0.0;
            }
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int kk = lane + 32 * c;                         // column of V and of Y handled by this lane
            if (kk < KP) {
              double sa = 0.0, sl = 0.0;
#pragma unroll
              for (int u = 0; u < K2_PR; ++u) {
                double v = rv[u][c] * w[u];
                if (has_g) {
                  sa += v;
                  sl += (v != 0.0) ? v * log(gv[u][c]) : 0.0;     // feeds the dof condition, pmc.pyx:672-679
                  v *= gv[u][c];
                }
                if (rb + u < TN) Vs[(rb + u) * KP + kk] = v;
              }
              if (has_g) { cs_a[kk] += sa; cs_a[KP + kk] += sl; }  // same thread every time: ordered, no race
            }
            if (kk < DP4) {
              const double sh = shift_s[kk];
#pragma unroll
              for (int u = 0; u < K2_PR; ++u) {
                double y = 0.0;
                if (rb + u < rows) y = (kk < D) ? (xv[u][c] - sh) : ((kk == D) ? 1.0 : 0.0);
                if (rb + u < TN) Ys[(rb + u) * DP4 + kk] = y;
              }
            }
          }
        }
      } else {
        for (int rb = pw * K2_PR; rb < TN; rb += 4 * K2_PR) {       // K2_PR rows in flight per warp
          double w[K2_PR];
  #pragma unroll
          for (int u = 0; u < K2_PR; ++u) w[u] = (a.sw && rb + u < rows) ? __ldg(a.sw + row0 + rb + u) : 1.0;
          for (int kk = lane; kk < KP; kk += 32) {
            double rv[K2_PR], gv[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              const bool in = (rb + u < rows) && (kk < a.k);
              rv[u] = in ? __ldg(a.rho + (row0 + rb + u) * a.ld_rho + kk) : 0.0;
              gv[u] = (in && has_g) ? __ldg(a.gamma + (row0 + rb + u) * a.ld_rho + kk) : 1.0;
            }
            double sa = 0.0, sl = 0.0;
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              double v = rv[u] * w[u];
              if (has_g) {
                sa += v;
                sl += (v != 0.0) ? v * log(gv[u]) : 0.0;        // feeds the dof condition, pmc.pyx:672-679
                v *= gv[u];
              }
              if (rb + u < TN) Vs[(rb + u) * KP + kk] = v;
            }
            if (has_g) { cs_a[kk] += sa; cs_a[KP + kk] += sl; }  // same thread every time: ordered, no race
          }
          for (int jj = lane; jj < DP4; jj += 32) {
            double xv[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u)
              xv[u] = (rb + u < rows && jj < D) ? __ldg(a.x + (row0 + rb + u) * a.ldx + jj) : 0.0;
            const double sh = shift_s[jj];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              double y = 0.0;
              if (rb + u < rows) y = (jj < D) ? (xv[u] - sh) : ((jj == D) ? 1.0 : 0.0);
              if (rb + u < TN) Ys[(rb + u) * DP4 + jj] = y;
            }
          }
        }
      }
      mbar_arrive(&full[s]);                                   // release: this thread's stores are visible to waiters
    }
    // ---- column sums (gamma) / zero column (no gamma), CTA row 0 only ----
    if (blockIdx.y == 0) {
      asm volatile("bar.sync 1, %0;" ::"n"(K2_PRODUCERS) : "memory");   // producers only
      const int ptid = tid - K2_CONSUMERS;
      for (int k = ptid; k < a.k; k += K2_PRODUCERS) {
        if (has_g) {
          double sa = 0.0, sl = 0.0;
          for (int wv = 0; wv < 4; ++wv) { sa += colsum[wv * 2 * KP + k]; sl += colsum[wv * 2 * KP + KP + k]; }
          out[size_t(k) * ldp] = sa;
          out[size_t(k) * ldp + a.F + 1] = sl;
        } else {
          out[size_t(k) * ldp + a.F + 1] = 0.0;
        }
      }
    }
    return;
  }

  // ================================= consumer warpgroups =================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
  // ---- this thread's lane tile ----
  const int t = blockIdx.y * K2_CONSUMERS + tid;
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

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int s = it % K2_STAGES;
    mbar_wait(&full[s], uint32_t((it / K2_STAGES) & 1));
    const double* Vs = stage0 + s * stage_len;
    const double* Ys = Vs + TN * KP;
    // ---- rank-TN update of the register tile ----
    const double* vp = Vs + kg * K2_TK;
    const double* yr = Ys + off_a;
    const double* yp = Ys + off_b;
#pragma unroll 2
    for (int n = 0; n < TN; ++n) {
      const double2 ya = *reinterpret_cast<const double2*>(yr + n * DP4);
      const double2 yb = *reinterpret_cast<const double2*>(yp + n * DP4);
      const double f0 = lin ? yb.x : ya.x * yb.x, f1 = lin ? yb.y : ya.x * yb.y;
      const double f2 = lin ? ya.x : ya.y * yb.x, f3 = lin ? ya.y : ya.y * yb.y;
#pragma unroll
      for (int c = 0; c < K2_TK / 2; ++c) {
        const double2 v = *reinterpret_cast<const double2*>(vp + n * KP + 2 * c);
        acc[2 * c][0] = fma(v.x, f0, acc[2 * c][0]);
        acc[2 * c][1] = fma(v.x, f1, acc[2 * c][1]);
        acc[2 * c][2] = fma(v.x, f2, acc[2 * c][2]);
        acc[2 * c][3] = fma(v.x, f3, acc[2 * c][3]);
        // snake order: consecutive DFMAs share v or f alternately, so each needs one new register operand plus its
        // accumulator -- the register file delivers ~1 64-bit warp operand per clock (scripts/ubench/dfma_snake.cu)
        acc[2 * c + 1][3] = fma(v.y, f3, acc[2 * c + 1][3]);
        acc[2 * c + 1][2] = fma(v.y, f2, acc[2 * c + 1][2]);
        acc[2 * c + 1][1] = fma(v.y, f1, acc[2 * c + 1][1]);
        acc[2 * c + 1][0] = fma(v.y, f0, acc[2 * c + 1][0]);
      }
    }
    mbar_arrive(&empty[s]);                                     // the producer may refill this stage
  }

  // ---- write this CTA's partial block ----
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
