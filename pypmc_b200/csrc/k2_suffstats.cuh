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
// Round 2: the matrix-instruction consumers read a TRANSPOSED, swizzled stage (8-sample blocks, one LDS.128 per operand
// and two 4-sample steps -- k2_stage / k2_swz below), the producers pull their next tile into L2 ahead of time and
// issue no FP64 instruction that is not needed (their DMULs queue behind the consumers' DMMAs), and the feature
// blocks are dealt evenly over the warps where that lowers the busiest scheduler's count.
// Each CTA writes one partial block; a second tiny kernel adds the partials in CTA order -- no floating-point
// atomics, so results are reproducible run to run for a given grid.  The default consumer form issues the update as
// FP64 matrix instructions (DMMA, see the template note below); the DFMA register tile described here is kept
// behind PMCB200_K2_FORM=dfma for comparison (13.4 ms vs 10.4 ms at N=1e7, K=32, D=30).
#pragma once

#include "pmc_common.cuh"

#include <type_traits>

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
  int tn;               // samples per tile (multiple of 8)
  int VS, YS;           // row strides (doubles) of the staged v and [y, 1] rows
  int nFB;              // MMA form: feature blocks of 8 = ceil(F / 8)
  int fchunks;          // MMA form: CTAs (gridDim.y direction) sharing the feature blocks
  int fbw;              // MMA form: feature blocks per warp (= template FB)
  int chunked;          // MMA form, several CTAs share the component blocks: each stages only its own columns of v
  int kchunk, cb8;      // ... kchunk columns staged (multiple of 16), first column = (blockIdx.y / fchunks) * cb8
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

// Producer step for KP <= 32 CV and DP4 <= 64: PR rows per warp pass, every global load of the pass issued before its
// first use.  Lane l handles columns l + 32 c of V (c < CV) and of Y (c < 2).
// TR (the matrix-instruction consumers): the stage is TRANSPOSED in blocks of 8 samples, V [TN/8][KW][8] | Y [TN/8][DP4][8],
// so that a consumer lane fetches the operands of two 4-sample steps (samples 2 tq, 2 tq + 1 of the block) with ONE
// LDS.128 -- an LDS instruction of either width costs the sub-partition about 2.8 clk of FP64 issue
// (scripts/ubench/k2_feed.cu), and halving their number is worth 5 (C2) to 8 (C4) points of the pipe.
// Inside a column's 64 bytes the four 16-byte sample pairs are XOR-swizzled with bits 1..2 of the column index: the
// producers (lane = column, one pair per store) then hit eight different 16-byte bank groups per quarter-warp, and a
// consumer quarter-warp (two columns x four pairs) still covers 128 contiguous bytes.
__device__ __forceinline__ int k2_swz(int col, int pair) { return pair ^ ((col >> 1) & 3); }
template <bool TR, int PR>
__device__ __forceinline__ void k2_store_rows(double* base, int rb, int col, int ncols, const double (&v)[PR]) {
  if constexpr (TR) {
    static_assert(PR == 2 || PR == 4, "rows of a pass share one 8-sample block");
    double* blk = base + (rb >> 3) * (ncols * 8) + col * 8;
#pragma unroll
    for (int u = 0; u < PR; u += 2)
      *reinterpret_cast<double2*>(blk + 2 * k2_swz(col, ((rb & 7) + u) >> 1)) = make_double2(v[u], v[u + 1]);
  } else {
    double* dst = base + rb * ncols + col;
#pragma unroll
    for (int u = 0; u < PR; ++u) dst[u * ncols] = v[u];
  }
}

template <int PR, int CV, bool TR>
__device__ __forceinline__ void k2_fill_stage(const StatsArgs& a, double* Vs, double* Ys, const double* shift_s,
                                              int64_t row0, int rows, int pw, int lane, bool has_g, int KP,
                                              const double* __restrict__ rho, const double* __restrict__ gamma, int kvalid) {
  // rho / gamma point at this CTA's first staged column; kvalid of the KP staged columns hold components
  const int D = a.d, DP4 = a.DP4, TN = a.tn, VS = a.VS, YS = a.YS;   // TR: VS / YS = columns per block
  const bool has_w = a.sw != nullptr;
  for (int rb = pw * PR; rb < TN; rb += 4 * PR) {
    double w[PR], rv[PR][CV], gv[PR][CV], xv[PR][2];
#pragma unroll
    for (int u = 0; u < PR; ++u) {
      const bool rin = rb + u < rows;
      const int64_t row = row0 + rb + u;
      w[u] = (a.sw && rin) ? __ldg(a.sw + row) : 1.0;
#pragma unroll
      for (int c = 0; c < CV; ++c) {
        const int kk = lane + 32 * c;
        const bool in = rin && kk < kvalid;
        rv[u][c] = in ? __ldg(rho + row * a.ld_rho + kk) : 0.0;
        gv[u][c] = (in && has_g) ? __ldg(gamma + row * a.ld_rho + kk) : 1.0;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int kk = lane + 32 * c;
        xv[u][c] = (rin && kk < D) ? __ldg(a.x + row * a.ldx + kk) : 0.0;
      }
    }
    // (TN is a multiple of 8 and rb of PR, so the PR rows of a pass lie inside the stage and inside one 8-sample block)
#pragma unroll
    for (int c = 0; c < CV; ++c) {
      const int kk = lane + 32 * c;                         // column of V handled by this lane
      if (kk < KP) {
        double v[PR];
#pragma unroll
        for (int u = 0; u < PR; ++u) {
          // (the producers' FP64 instructions queue behind the consumers' DMMAs: none that is not needed)
          v[u] = rv[u][c];
          if (has_w) v[u] *= w[u];
          if (has_g) v[u] *= gv[u][c];
        }
        k2_store_rows<TR, PR>(Vs, rb, kk, VS, v);
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int kk = lane + 32 * c;                         // column of Y handled by this lane
      if (kk < DP4) {
        const double sh = shift_s[kk];
        double y[PR];
#pragma unroll
        for (int u = 0; u < PR; ++u) {
          y[u] = 0.0;
          if (rb + u < rows) y[u] = (kk < D) ? (xv[u][c] - sh) : ((kk == D) ? 1.0 : 0.0);
        }
        k2_store_rows<TR, PR>(Ys, rb, kk, YS, y);
      }
    }
  }
}

// Stage layout (doubles): V [TN][KP] | Y [TN][DP4].  The producer warpgroup (88 registers) reads rho / gamma /
// x / w rows from global memory, forms v = w rho gamma and yh = [x - shift, 1, 0...] and stores them; the two
// consumer warpgroups (208 registers, setmaxnreg) only ever touch shared memory and the FP64 pipe.
//
// Consumer forms (template): CB == 0 -- the DFMA register tile described above.  CB > 0 -- the same rank-N update
// issued as FP64 matrix instructions (mma.sync.m8n8k4.f64, SASS DMMA): a warp owns CB x FB tiles of 8 components x
// 8 features, per 4 samples it loads CB A-fragments (v) and forms FB B-fragments (products of two [y, 1] entries)
// and issues CB x FB DMMAs.  Why: a DFMA reads three 64-bit register operands and the register file delivers about
// one per clock, which caps the DFMA form at 66-83 % of the pipe's peak (profiles/r01_operand_delivery.md); a
// DMMA moves 256 FMAs with four operand registers per thread and runs at the full 37 TFLOP/s.
template <int CB, int FB>
__global__ void __launch_bounds__(K2_THREADS, 1) k2_suffstats(const StatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool TR = CB > 0;                       // stage layout: transposed 8-sample blocks for the matrix-instruction form
  const int tid = threadIdx.x;
  const int D = a.d, KP = a.KP, DP4 = a.DP4, TN = a.tn, VS = a.VS, YS = a.YS;
  const bool has_g = a.gamma != nullptr;
  // columns of v this CTA stages: all KP, or (several CTAs sharing the component blocks) its own kchunk from koff on
  const int KW = a.chunked ? a.kchunk : KP, koff = a.chunked ? int(blockIdx.y / a.fchunks) * a.cb8 : 0;
  const int kvalid = min(KW, a.k - koff);
  const double* const rho_c = a.rho + koff;
  const double* const gamma_c = has_g ? a.gamma + koff : nullptr;
  const int stage_len = TN * (VS + YS);
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  double* shift_s = stage0 + K2_STAGES * stage_len;               // [DP4]
  uint64_t* full = reinterpret_cast<uint64_t*>(shift_s + DP4);
  uint64_t* empty = full + K2_STAGES;

  for (int j = tid; j < DP4; j += K2_THREADS) shift_s[j] = (j < D) ? a.shift[j] : 0.0;
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
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it % K2_STAGES;
      mbar_wait(&empty[s], uint32_t(((it / K2_STAGES) & 1) ^ 1));
      double* Vs = stage0 + s * stage_len;
      double* Ys = Vs + TN * VS;
      const int64_t row0 = tile * TN;
      const int rows = int((a.n - row0 < TN) ? (a.n - row0) : TN);
      // The producers are latency-bound (a pass = issue every load of 4 rows, wait, store): pull the rows of this CTA's
      // NEXT tile into L2 now, so that the loads of the next stage cost an L2 round trip instead of an HBM one.  With few
      // FMAs per staged byte (K = 40, D = 20: consumers 20 % of their time on `full`) that is what bounds the kernel.
      {
        const int64_t nrow = row0 + int64_t(gridDim.x) * TN + (tid - K2_CONSUMERS);
        if (tid - K2_CONSUMERS < TN && nrow < a.n) {
          const char* pr = reinterpret_cast<const char*>(rho_c + nrow * a.ld_rho);
          for (int b = 0; b < kvalid * 8; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + b));
          if (has_g) {
            const char* pg = reinterpret_cast<const char*>(gamma_c + nrow * a.ld_rho);
            for (int b = 0; b < kvalid * 8; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pg + b));
          }
          const char* px = reinterpret_cast<const char*>(a.x + nrow * a.ldx);
          for (int b = 0; b < D * 8; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + b));
          if (a.sw && ((tid - K2_CONSUMERS) & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sw + nrow));
        }
      }
      if (KW <= 64 && DP4 <= 64) {
        // common sizes: every global load of a 4-row step is issued before the first use (about 20 in flight per
        // thread), so a step costs one memory latency instead of one per column block
        k2_fill_stage<K2_PR, 2, TR>(a, Vs, Ys, shift_s, row0, rows, pw, lane, has_g, KW, rho_c, gamma_c, kvalid);
      } else if (KW <= 128 && DP4 <= 64) {
        // 65..128 components: the same with two rows in flight and four column blocks per lane (the register budget of
        // the producer warps, 88, holds 2 x (4 rho + 4 gamma + 2 x) values)
        k2_fill_stage<2, 4, TR>(a, Vs, Ys, shift_s, row0, rows, pw, lane, has_g, KW, rho_c, gamma_c, kvalid);
      } else {
        for (int rb = pw * K2_PR; rb < TN; rb += 4 * K2_PR) {       // K2_PR rows in flight per warp
          double w[K2_PR];
  #pragma unroll
          for (int u = 0; u < K2_PR; ++u) w[u] = (a.sw && rb + u < rows) ? __ldg(a.sw + row0 + rb + u) : 1.0;
          for (int kk = lane; kk < KW; kk += 32) {
            double rv[K2_PR], gv[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              const bool in = (rb + u < rows) && (kk < kvalid);
              rv[u] = in ? __ldg(rho_c + (row0 + rb + u) * a.ld_rho + kk) : 0.0;
              gv[u] = (in && has_g) ? __ldg(gamma_c + (row0 + rb + u) * a.ld_rho + kk) : 1.0;
            }
            double v[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              v[u] = rv[u];
              if (a.sw) v[u] *= w[u];
              if (has_g) v[u] *= gv[u];
            }
            k2_store_rows<TR, K2_PR>(Vs, rb, kk, VS, v);
          }
          for (int jj = lane; jj < DP4; jj += 32) {
            double xv[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u)
              xv[u] = (rb + u < rows && jj < D) ? __ldg(a.x + (row0 + rb + u) * a.ldx + jj) : 0.0;
            const double sh = shift_s[jj];
            double y[K2_PR];
  #pragma unroll
            for (int u = 0; u < K2_PR; ++u) {
              y[u] = 0.0;
              if (rb + u < rows) y[u] = (jj < D) ? (xv[u] - sh) : ((jj == D) ? 1.0 : 0.0);
            }
            k2_store_rows<TR, K2_PR>(Ys, rb, jj, YS, y);
          }
        }
      }
      mbar_arrive(&full[s]);                                   // release: this thread's stores are visible to waiters
    }
    // column 0 (A) with gamma and column F+1 (L) are written by k2_colsums / k2_colsums_final
    return;
  }

  // ================================= consumer warpgroups =================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
  if constexpr (CB > 0) {
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;            // fragment coordinates: A row / B col / C row = g, k index = tq
    const int chunk_f = blockIdx.y % a.fchunks, chunk_c = blockIdx.y / a.fchunks;
    // Feature blocks of this CTA: [cta_first, cta_first + n_cta).  Two ways to deal them to the 8 warps:
    //   in order -- FB blocks each until they run out (the last warps may hold fewer or none);
    //   evenly   -- n_cta / 8 each, the first `rem` warps one more; those with FB - 1 run a second copy of the loop.
    // Warps w and w + 4 share a scheduler, so what counts is the busiest scheduler's total: D = 40 has 108 blocks, in order
    // 7 x 14 + 10 (schedulers 28, 28, 28, 24), evenly 4 x 14 + 4 x 13 (27 each).  The even deal is taken only where it lowers
    // that peak (at C2 and C3 both deals peak at 16 and 8, and the in-order one measured the same or better), only for
    // FB >= 8 (a block less is not worth a second loop copy on small tiles) and only when every warp then holds FB or FB - 1
    // blocks (the instantiated FB values are coarse for some CB).
    const int per_cta = (a.nFB + a.fchunks - 1) / a.fchunks, cta_first = chunk_f * per_cta;
    const int n_cta = max(0, min(per_cta, a.nFB - cta_first));
    constexpr int NWC = K2_CONSUMERS / 32;
    const int fb_base = n_cta / NWC, fb_rem = n_cta - fb_base * NWC;
    int peak_in_order = 0, peak_even = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int a0 = max(0, min(FB, n_cta - w * FB)), a1 = max(0, min(FB, n_cta - (w + 4) * FB));
      peak_in_order = max(peak_in_order, a0 + a1);
      peak_even = max(peak_even, 2 * fb_base + (w < fb_rem ? 1 : 0) + (w + 4 < fb_rem ? 1 : 0));
    }
    constexpr bool K2_EVEN_DEAL = FB >= 8;
    const bool even_deal = K2_EVEN_DEAL && n_cta >= NWC * (FB - 1) && peak_even < peak_in_order;
    const int fb_first = cta_first + (even_deal ? warp * fb_base + min(warp, fb_rem) : warp * FB);   // first block of this warp
    const int cb_first = chunk_c * CB, cb_total = KP / 8;
    const int nfb_w = even_deal ? min(FB, fb_base + (warp < fb_rem ? 1 : 0))   // (the host picks FB >= ceil(per_cta / 8))
                                : max(0, min(FB, n_cta - warp * FB));
    // feature f -> the two entries of [y, 1, 0] whose product it is: f = 0: 1*1 (B_k); 1..D: y_i * 1 (m_k); then the
    // lower triangle of y y^T row-major; beyond F: the zero column D+1
    int off_i[FB], off_j[FB];
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) {
      const int f = (fb_first + fb) * 8 + g;
      int oi = D + 1, oj = D + 1;
      if (fb >= nfb_w) { }                             // not this warp's block: the zero column
      else if (f == 0) { oi = D; oj = D; }
      else if (f <= D) { oi = f - 1; oj = D; }
      else if (f < a.F) {
        const int t = f - 1 - D;
        int r = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= t) ++r;
        while (r * (r + 1) / 2 > t) --r;
        oi = r; oj = t - r * (r + 1) / 2;
      }
      off_i[fb] = oi * 64 + 16 * k2_swz(oi, tq);          // byte offsets inside an 8-sample block: column, swizzled pair tq
      off_j[fb] = oj * 64 + 16 * k2_swz(oj, tq);
    }
    // staged column of each of this warp's component blocks: with `chunked` the CTA staged only its own columns (from
    // 0), otherwise all of them; blocks beyond KP/8 (cb_total not a multiple of CB) re-read column block 0 and are
    // dropped at the end
    int cb_off[CB];
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
    {
      const int col = ((cb_first + cb < cb_total) ? (a.chunked ? 0 : cb_first * 8) + cb * 8 : 0) + g;
      cb_off[cb] = col * 64 + 16 * k2_swz(col, tq);        // bytes
    }
    double acc[CB][FB][2];
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
      for (int fb = 0; fb < FB; ++fb) { acc[cb][fb][0] = 0.0; acc[cb][fb][1] = 0.0; }

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it % K2_STAGES;
      mbar_wait(&full[s], uint32_t((it / K2_STAGES) & 1));
      const double* Vs = stage0 + s * stage_len;
      const double* Ys = Vs + TN * VS;
      // no branch inside the step: every load of a pass can be issued before its first DMMA and the DMMAs go back to back.
      // One pass = an 8-sample block = two 4-sample steps (samples 2 tq and 2 tq + 1 of the block are the k index of the
      // first / second step): CB + 2 NF LDS.128, 2 NF DMUL, 2 CB NF DMMA.  NF = the warp's own block count (FB or FB - 1:
      // no issued work on blocks it does not have); any other count runs all FB slots against the zero column.
      auto consume = [&](auto nf_tag) {
        constexpr int NF = decltype(nf_tag)::value;
        const char* vblk = reinterpret_cast<const char*>(Vs);
        const char* yblk = reinterpret_cast<const char*>(Ys);
        for (int n0 = 0; n0 < TN; n0 += 8, vblk += VS * 64, yblk += YS * 64) {
          double2 av[CB];
          double b0[NF], b1[NF];
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) av[cb] = *reinterpret_cast<const double2*>(vblk + cb_off[cb]);
#pragma unroll
          for (int fb = 0; fb < NF; ++fb) {
            const double2 yi = *reinterpret_cast<const double2*>(yblk + off_i[fb]);
            const double2 yj = *reinterpret_cast<const double2*>(yblk + off_j[fb]);
            b0[fb] = yi.x * yj.x;
            b1[fb] = yi.y * yj.y;
          }
#pragma unroll
          for (int fb = 0; fb < NF; ++fb)
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(acc[cb][fb][0]), "+d"(acc[cb][fb][1])
                           : "d"(av[cb].x), "d"(b0[fb]));
#pragma unroll
          for (int fb = 0; fb < NF; ++fb)
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(acc[cb][fb][0]), "+d"(acc[cb][fb][1])
                           : "d"(av[cb].y), "d"(b1[fb]));
        }
      };
      if (K2_EVEN_DEAL && nfb_w == FB - 1) consume(std::integral_constant<int, (K2_EVEN_DEAL ? FB - 1 : FB)>());
      else if (nfb_w > 0) consume(std::integral_constant<int, FB>());
      mbar_arrive(&empty[s]);                                   // the producer may refill this stage
    }
    // ---- write this CTA's partial block: lane holds Out[8 cb + g][8 fb + 2 tq + e] ----
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      const int k = (cb_first + cb) * 8 + g;
      if (k >= a.k) continue;
#pragma unroll
      for (int fb = 0; fb < FB; ++fb) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int f = (fb_first + fb) * 8 + 2 * tq + e;
          if (fb >= nfb_w || f >= a.F) continue;
          out[size_t(k) * ldp + 1 + f] = acc[cb][fb][e];
          if (f == 0 && !has_g) out[size_t(k) * ldp] = acc[cb][fb][e];   // A == B without gamma
        }
      }
    }
    return;
  } else {
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
      const double* Ys = Vs + TN * VS;
      // ---- rank-TN update of the register tile ----
      const double* vp = Vs + kg * K2_TK;
      const double* yr = Ys + off_a;
      const double* yp = Ys + off_b;
  #pragma unroll 2
      for (int n = 0; n < TN; ++n) {
        const double2 ya = *reinterpret_cast<const double2*>(yr + n * YS);
        const double2 yb = *reinterpret_cast<const double2*>(yp + n * YS);
        const double f0 = lin ? yb.x : ya.x * yb.x, f1 = lin ? yb.y : ya.x * yb.y;
        const double f2 = lin ? ya.x : ya.y * yb.x, f3 = lin ? ya.y : ya.y * yb.y;
  #pragma unroll
        for (int c = 0; c < K2_TK / 2; ++c) {
          const double2 v = *reinterpret_cast<const double2*>(vp + n * VS + 2 * c);
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
}

#ifndef PMC_K2_TEMPLATE_ONLY   // (k2_inst.cu instantiates more tile shapes of the template above and needs only that)
}  // namespace pmc
#include "k1_exp_table.cuh"
namespace pmc {
// ---------------------------------------------------------------------------------------------
// Column sums that go with gamma (Student-t): A_k = sum_n w_n rho_nk and L_k = sum_n w_n rho_nk ln(gamma_nk)
// (the N-sized part of the dof condition, pmc.pyx:654-691).  A streaming pass of its own: in the producer warps of
// k2_suffstats the logarithm cost 2 ms of 8 at C4 (4 warps per SM, latency-bound); here every SM runs 8 full CTAs.
// Thread t owns column t % kc and rows t / kc, t / kc + 256 / kc, ...; per-CTA partials, summed in CTA order.
// ---------------------------------------------------------------------------------------------
// ln(x) for any positive normal x (else the library function): x = 2^e m, m in [1, 2), c_j = 1 + (j + 1/2)/128 for the top
// seven mantissa bits, r = m / c_j - 1, ln x = e ln2 + ln c_j + (r - r^2/2 + ... - r^6/6) -- the table logarithm of K1's
// Student-t epilogue (k1_mma_eval.cuh: log_tab) with the exponent term formed by two FMAs, so arguments below 1 (gamma
// < 1 for samples beyond the component's bulk) need no table entry.  10 FP64 instructions; absolute error <= 2e-16 (1 + |ln x|).
__device__ __forceinline__ double k2_log_pos(double x, const double* __restrict__ tab /* [1/c_j | ln c_j] */) {
  const int hi = __double2hiint(x);
  if (unsigned(hi) - 0x00100000u >= 0x7fe00000u) return log(x);        // zero, subnormal, negative, inf, nan
  const int j = (hi >> 13) & 127;
  const double e = double((hi >> 20) - 1023);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, tab[j], -1.0);
  double u = fma(r, -1.66666666666666657e-01, 2.00000000000000011e-01);
  u = fma(r, u, -0.25);
  u = fma(r, u, 3.33333333333333315e-01);
  u = fma(r, u, -0.5);
  const double p = fma(r * r, u, r);
  return fma(e, 0x1.62e42fefa38p-1, (p + tab[128 + j]) + e * 0x1.ef35793c7673p-45);   // ln 2 = hi (43 bits: e hi exact) + lo
}

__global__ void __launch_bounds__(256) k2_colsums(const double* __restrict__ rho, const double* __restrict__ gamma,
                                                  const double* __restrict__ sw, int64_t n, int k, int ld_rho, int kc,
                                                  double* __restrict__ partial /* [grid][2][k] */) {
  __shared__ double red[2][256];
  __shared__ double ltab[256];
  const int tid = threadIdx.x, rows_per_pass = 256 / kc, r_off = tid / kc;
  if (tid < 128) { ltab[tid] = kLogInvC[tid]; ltab[128 + tid] = kLogC[tid]; }
  __syncthreads();
  constexpr int U = 4;                                               // rows in flight per thread
  for (int k0 = 0; k0 < k; k0 += kc) {
    const int kk = k0 + tid % kc;
    double sa = 0.0, sl = 0.0;
    if (kk < k) {
      const int64_t stride = int64_t(gridDim.x) * rows_per_pass;
      for (int64_t r = int64_t(blockIdx.x) * rows_per_pass + r_off; r < n; r += U * stride) {
        double v[U], gm[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t ru = r + u * stride;
          const bool in = ru < n;
          v[u] = in ? __ldg(rho + ru * ld_rho + kk) : 0.0;
          gm[u] = in ? __ldg(gamma + ru * ld_rho + kk) : 1.0;
          if (sw && in) v[u] *= __ldg(sw + ru);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {                                 // (row order as before: r, r + stride, ...)
          sa += v[u];
          sl += (v[u] != 0.0) ? v[u] * k2_log_pos(gm[u], ltab) : 0.0;
        }
      }
    }
    red[0][tid] = sa;
    red[1][tid] = sl;
    __syncthreads();
    if (tid < kc && kk < k) {
      double a0 = 0.0, l0 = 0.0;
      for (int u = tid; u < 256; u += kc) { a0 += red[0][u]; l0 += red[1][u]; }     // fixed order
      partial[(size_t(blockIdx.x) * 2 + 0) * k + kk] = a0;
      partial[(size_t(blockIdx.x) * 2 + 1) * k + kk] = l0;
    }
    __syncthreads();
  }
}

// out[k][0] = A_k, out[k][F+1] = L_k from the per-CTA partials; L_k = 0 without gamma.  One CTA per component: thread t
// adds the partials of CTAs t, t + 256, ... in ascending order, then a fixed tree over the threads -- reproducible, and
// microseconds where one thread per component walking all ~1200 partials took 0.2 ms.
__global__ void __launch_bounds__(256) k2_colsums_final(const double* __restrict__ partial, int nblocks, int k, int ldp,
                                                        int F, int has_gamma, double* __restrict__ out) {
  __shared__ double red[2][256];
  const int kk = blockIdx.x, tid = threadIdx.x;
  if (kk >= k) return;
  if (!has_gamma) {
    if (tid == 0) out[size_t(kk) * ldp + F + 1] = 0.0;
    return;
  }
  double a0 = 0.0, l0 = 0.0;
  for (int b = tid; b < nblocks; b += 256) {
    a0 += partial[(size_t(b) * 2 + 0) * k + kk];
    l0 += partial[(size_t(b) * 2 + 1) * k + kk];
  }
  red[0][tid] = a0;
  red[1][tid] = l0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { red[0][tid] += red[0][tid + o]; red[1][tid] += red[1][tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    out[size_t(kk) * ldp] = red[0][0];
    out[size_t(kk) * ldp + F + 1] = red[1][0];
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

#endif  // PMC_K2_TEMPLATE_ONLY

// k2_suffstats<CB, FB> for CB = 3, 5, 6, 7 (k2_inst.cu); returns 2 when (cb, fb) is not instantiated
int k2_launch_extra(int cb, int fb, const StatsArgs& a, dim3 grid, size_t smem, cudaStream_t stream);

}  // namespace pmc
