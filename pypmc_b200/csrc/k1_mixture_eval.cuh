// k1_mixture_eval.cuh -- K1, exact-difference form: fused per-sample x per-component log-pdf + mixture
// log-sum-exp + responsibilities, float64, sm_100a.  This is the FALLBACK of K1: the form that normally runs is
// k1_fast_eval.cuh; this one takes over (device-side flag from k1_prepare, no host synchronisation) when a component
// lies so far from the common shift that the fast form's rounding bound D eps max|b| would exceed the contract.
// It forms y = x - mu_k per component exactly like the reference does and was the first K1 of round 1
// (16.2 ms at N=1e7, K=32, D=30 against 12.9 ms for the fast form).
//
// Replaces (reference loops, /root/reference/pypmc):
//   Gauss.multi_evaluate            density/gauss.pyx:146-151      (bilinear_sym tools/_linalg.pyx:10-39)
//   StudentT.multi_evaluate         density/student_t.pyx:154-164
//   MixtureDensity.multi_evaluate   density/mixture.pyx:112-156    (logsumexp2D tools/_regularize.pyx:57-83)
//   calculate_rho_rb                mix_adapt/pmc.pyx:23-43
//   Student-t gamma_nk              mix_adapt/pmc.pyx:602-610
//   VB expectation_gauss_exponent / log_rho / r   mix_adapt/variational.pyx:774-798, 675-691, 711-757
//   PMC.log_likelihood reduction    mix_adapt/pmc.pyx:371-391 ;  VB E[log q(Z)]  variational.pyx:1003-1013
//
// Shape of the computation: q_nk = || T_k (x_n - c_k) ||^2 with T_k lower triangular
// (T_k^T T_k = Sigma_k^-1), i.e. D(D+1)/2 + D FP64 FMAs per (sample, component) pair against 8 D
// bytes per sample: ~65 flop/byte at K=32, D=30, so the DFMA pipe is the roof, not HBM.
//
// Mapping: one thread owns S samples; their differences y = x - c_k live in registers (2 S DP
// registers), so every T element fetched from shared memory (one broadcast LDS.128 = two elements)
// feeds 2 S DFMAs.  A CTA is persistent (one per SM); each warp owns a private 32 S-row slice of
// the CTA's sample tile in shared memory and re-reads it once per component.  Component records
// (T_k, c_k, scalars; ~4 KB) are streamed through a 3-stage shared-memory ring by the TMA engine
// (cp.async.bulk + mbarrier complete_tx); the warp that releases a stage last re-arms it, so no
// warp ever blocks on an "empty" barrier and there is no block-wide barrier in the steady state.
#pragma once

#include "pmc_common.cuh"

namespace pmc {

struct EvalArgs {
  const double* x;        // [n, ldx] samples (device)
  int64_t n;
  int64_t ldx;
  int d;                  // true dimension (DP-1 or DP)
  const double* records;  // [kl, record_len(DP)] packed components to evaluate
  const int* cols;        // [kl] output column of each record
  int kl;
  int k_out;              // row stride (number of columns) of the N x K outputs
  int mode;               // Mode
  double max_init;        // -DBL_MAX, or 0 when dead columns (value 0) take part in the max
  double* logq;           // [n] or null
  double* lp_out;         // [n, k_out] or null : individual (mixture) / log_rho (VB)
  double* resp_out;       // [n, k_out] or null : rho (PMC) / r (VB)
  double* aux_out;        // [n, k_out] or null : gamma (Student-t) / expectation_gauss_exponent (VB)
  const double* sw;       // [n] sample weights or null
  double* partials;       // [grid * PMC_MAX_WARPS * 2] or null : per-warp (sum w*logq | sum w r log r , sum w)
  const int* flag;        // null: always run.  else the exact-difference kernel runs iff *flag != 0 (k1_fast_eval.cuh)
};

template <int DP>
struct EvalCfg {
  // samples per thread / warps per CTA, chosen so that y (2*S*DP registers) fits the register file
  // and the sample tile (S*32*NW rows) fits shared memory beside the record ring.
  static constexpr int S = (DP <= 20) ? 4 : (DP <= 40) ? 2 : 1;
  static constexpr int NW = (DP <= 12) ? 12 : (DP <= 20) ? 8 : (DP <= 32) ? 12 : 8;
  static constexpr int NS = 3;                                         // record ring depth
  static constexpr int XS = ((DP / 2) % 2 == 1) ? DP : DP + 2;         // x row stride: XS/2 odd => LDS.128 conflict-free
  static constexpr int ROWS_PER_WARP = 32 * S;
  static constexpr int TS = ROWS_PER_WARP * NW;                        // samples per CTA tile
  static constexpr int RL = record_len(DP);
  static constexpr size_t SMEM_BYTES =
      sizeof(double) * (size_t(NS) * RL + size_t(TS) * XS) + NS * sizeof(uint64_t) + NS * sizeof(int) + 16;
};

template <int DP>
__global__ void __launch_bounds__(EvalCfg<DP>::NW * 32, 1) k1_mixture_eval(const EvalArgs a) {
  using C = EvalCfg<DP>;
  constexpr int S = C::S, NW = C::NW, NS = C::NS, XS = C::XS, RL = C::RL, H = DP / 2;
  if (a.flag != nullptr && *a.flag == 0) return;                       // the fast form handled this launch
  constexpr int NT = tri_len(DP);
  constexpr uint32_t REC_BYTES = RL * sizeof(double);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);                  // [NS][RL]
  double* xs_all = ring + NS * RL;                                     // [TS][XS]
  uint64_t* full = reinterpret_cast<uint64_t*>(xs_all + size_t(C::TS) * XS);
  int* empty_cnt = reinterpret_cast<int*>(full + NS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* xs = xs_all + size_t(warp) * C::ROWS_PER_WARP * XS;          // this warp's private slice

  const int64_t num_tiles = (a.n + C::TS - 1) / C::TS;
  const int64_t my_tiles = (int64_t(blockIdx.x) < num_tiles) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total_steps = my_tiles * a.kl;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      empty_cnt[s] = 0;
    }
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS && s < total_steps; ++s) {
      mbar_arrive_expect_tx(&full[s], REC_BYTES);
      bulk_g2s(ring + s * RL, a.records + size_t(s % a.kl) * RL, REC_BYTES, &full[s]);
    }
  }

  double part_a = 0.0, part_w = 0.0;  // per-thread partial sums (fixed order => deterministic)
  double* const scratch = a.lp_out ? a.lp_out : a.resp_out;  // where lp_nk waits for the second pass
  const bool second_pass = (a.resp_out != nullptr) || (a.mode == MODE_VB && a.lp_out != nullptr);

  int64_t step = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * int64_t(gridDim.x);
    const int64_t row0 = tile * C::TS + int64_t(warp) * C::ROWS_PER_WARP;

    // ---- stage this warp's rows: coalesced global reads -> padded row-major shared slice ----
    {
      const int d = a.d;
      int r = 0, j = lane;
      while (j >= d) { j -= d; ++r; }
      const int dr = 32 / d, dj = 32 % d;
      for (; r < C::ROWS_PER_WARP;) {
        const int64_t row = row0 + r;
        xs[r * XS + j] = (row < a.n) ? __ldg(a.x + row * a.ldx + j) : 0.0;
        r += dr; j += dj;
        if (j >= d) { j -= d; ++r; }
      }
      if (d < DP) {  // zero the padding column of an odd dimension
        for (int rr = lane; rr < C::ROWS_PER_WARP; rr += 32) xs[rr * XS + d] = 0.0;
      }
      __syncwarp();
    }

    double run_max[S], run_sum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { run_max[s] = a.max_init; run_sum[s] = 0.0; }

    // ---- pass 1: all evaluated components ----
    for (int kk = 0; kk < a.kl; ++kk, ++step) {
      const int stage = int(step % NS);
      const uint32_t parity = uint32_t((step / NS) & 1);
      mbar_wait(&full[stage], parity);
      const double* rec = ring + stage * RL;
      const double* mu = rec + NT;
      const double* sc = rec + NT + DP;

      double y[S][DP];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const double* xr = xs + (lane + 32 * s) * XS;
#pragma unroll
        for (int p = 0; p < H; ++p) {
          const double2 xv = *reinterpret_cast<const double2*>(xr + 2 * p);
          const double2 mv = *reinterpret_cast<const double2*>(mu + 2 * p);
          y[s][2 * p] = xv.x - mv.x;
          y[s][2 * p + 1] = xv.y - mv.y;
        }
      }

      double q[S];
#pragma unroll
      for (int s = 0; s < S; ++s) q[s] = 0.0;
#pragma unroll
      for (int r = 0; r < H; ++r) {
        double z0[S], z1[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { z0[s] = 0.0; z1[s] = 0.0; }
#pragma unroll
        for (int p = 0; p <= r; ++p) {
          const double2 t0 = *reinterpret_cast<const double2*>(rec + 2 * r * (r + 1) + 4 * p);
          const double2 t1 = *reinterpret_cast<const double2*>(rec + 2 * r * (r + 1) + 4 * p + 2);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            z0[s] = fma(t0.x, y[s][2 * p], z0[s]);
            if (p < r) z0[s] = fma(t0.y, y[s][2 * p + 1], z0[s]);  // T[2r][2r+1] == 0
            z1[s] = fma(t1.x, y[s][2 * p], z1[s]);
            z1[s] = fma(t1.y, y[s][2 * p + 1], z1[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          q[s] = fma(z0[s], z0[s], q[s]);
          q[s] = fma(z1[s], z1[s], q[s]);
        }
      }

      const double c0 = sc[S0], c1 = sc[S1], c2 = sc[S2], c3 = sc[S3], c4 = sc[S4], wk = sc[S_WEIGHT];
      const int col = __ldg(a.cols + kk);
      __syncwarp();
      // ---- release the ring stage; the last warp to leave re-arms it (TMA refill) ----
      if (lane == 0) {
        __threadfence_block();
        const int old = atomicAdd(&empty_cnt[stage], 1);
        if (old == NW - 1) {
          empty_cnt[stage] = 0;
          const int64_t nxt = step + NS;
          if (nxt < total_steps) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&full[stage], REC_BYTES);
            bulk_g2s(ring + stage * RL, a.records + size_t(nxt % a.kl) * RL, REC_BYTES, &full[stage]);
          }
        }
      }

#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int64_t row = row0 + lane + 32 * s;
        double lp, aux;
        if (a.mode == MODE_GAUSS) {
          lp = c0 - 0.5 * q[s];                                   // gauss.pyx:151
          aux = 0.0;
        } else if (a.mode == MODE_STUDENT_T) {
          double t = q[s] * c2;                                   // student_t.pyx:159-164
          t += 1.0;
          t = log(t);
          t *= c1;
          lp = t + c0;
          aux = c4 / (c3 + q[s]);                                 // gamma_nk, pmc.pyx:610
        } else {
          aux = c3 + c4 * q[s];                                   // variational.pyx:798
          lp = c0 + 0.5 * (c1 - c2 - aux);                        // variational.pyx:691
        }
        // online weighted log-sum-exp (same value as _regularize.pyx:72-81 up to rounding)
        if (lp > run_max[s]) {
          run_sum[s] = run_sum[s] * exp(run_max[s] - lp) + wk;
          run_max[s] = lp;
        } else {
          run_sum[s] += wk * exp(lp - run_max[s]);
        }
        if (row < a.n) {
          if (scratch) scratch[row * a.k_out + col] = lp;
          if (a.aux_out) a.aux_out[row * a.k_out + col] = aux;
        }
      }
    }

    // ---- per-sample results + pass 2 (responsibilities) ----
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t row = row0 + lane + 32 * s;
      if (row >= a.n) continue;
      const double w_n = a.sw ? __ldg(a.sw + row) : 1.0;
      part_w += w_n;
      if (a.mode != MODE_VB) {
        const double lq = log(run_sum[s]) + run_max[s];           // _regularize.pyx:81
        if (a.logq) a.logq[row] = lq;
        part_a += w_n * lq;                                       // pmc.pyx:388-391
        if (second_pass) {
          const double den = exp(lq) + kTiny;                     // pmc.pyx:39-41
          for (int kk = 0; kk < a.kl; ++kk) {
            const int col = __ldg(a.cols + kk);
            const double wk = a.records[size_t(kk) * RL + NT + DP + S_WEIGHT];
            double v = exp(scratch[row * a.k_out + col]) * wk;
            v /= den;
            a.resp_out[row * a.k_out + col] = v;
          }
        }
      } else {
        if (a.logq) a.logq[row] = log(run_sum[s]) + run_max[s];
        if (second_pass) {
          const double norm_inv = 1.0 / run_sum[s];               // variational.pyx:728-755
          const double log_norm_inv = log(norm_inv);
          double acc = 0.0;
          for (int kk = 0; kk < a.kl; ++kk) {
            const int col = __ldg(a.cols + kk);
            const double lr = scratch[row * a.k_out + col] - run_max[s];
            double r = exp(lr) * norm_inv;
            if (r == 0.0) r = kTiny;
            const double lrn = lr + log_norm_inv;
            if (a.resp_out) a.resp_out[row * a.k_out + col] = r;
            if (a.lp_out) a.lp_out[row * a.k_out + col] = lrn;
            acc = fma(r, lrn, acc);                               // variational.pyx:1003-1013
          }
          part_a += w_n * acc;
        }
      }
    }
    __syncwarp();  // everyone is done with xs before the next tile overwrites it
  }

  if (a.partials) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      part_a += __shfl_xor_sync(0xffffffffu, part_a, o);
      part_w += __shfl_xor_sync(0xffffffffu, part_w, o);
    }
    if (lane == 0) {
      a.partials[(size_t(blockIdx.x) * PMC_MAX_WARPS + warp) * 2 + 0] = part_a;
      a.partials[(size_t(blockIdx.x) * PMC_MAX_WARPS + warp) * 2 + 1] = part_w;
    }
  }
}

}  // namespace pmc
