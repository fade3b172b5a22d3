// k1_fast_eval.cuh -- K1 (fast form): fused per-sample x per-component log-pdf + mixture log-sum-exp +
// responsibilities, float64, sm_100a.  Same outputs and reference citations as k1_mixture_eval.cuh (the
// "exact-difference" form); this is the form that runs unless the prepare kernel flags the inputs (below).
//
// What differs from k1_mixture_eval.cuh, and why (ncu, profiles/r01_*): there every component re-read the
// sample from shared memory to form y = x - mu_k (30 non-broadcast LDS.128 + 60 DADD per thread-component)
// and the LSU pipe (68 % busy) throttled the DFMA pipe (60 %).  Here
//     q_nk = || T_k (x_n - mu_k) ||^2 = || T_k x'_n - b_k ||^2,   x' = x - c,  b_k = T_k (mu_k - c)
// with one shift c for the whole mixture (weighted mean of the centres, k1_prepare), so that
//   * x' is loaded ONCE per tile straight into registers (2 S DP registers) and stays there for all K
//     components: no shared-memory sample tile at all;
//   * per component the thread only fetches T (broadcast LDS.128, two elements each, feeding 2 S DFMAs)
//     and starts its accumulators at -b_k: D(D+1)/2 + D DFMAs per sample-component, nothing else on the
//     FP64 pipe but the epilogue's exp/log.
// Rounding: the partial sums are of size |b| instead of |x - mu|, so the relative error of q grows from
// ~D eps to ~D eps max|b|.  k1_prepare computes max|b| and raises a flag above 1e4 (or for non-finite
// parameters); the fast kernel then returns at once and the exact-difference kernel, launched right
// behind it on the same stream, does the work instead (and vice versa) -- no host synchronisation.
//
// Mapping: persistent CTAs (one per SM), NW warps, each thread owns S samples.  Component records
// (T | -b | scalars) stream through a 3-stage shared-memory ring filled by the TMA engine
// (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP); the warp that leaves a stage last re-arms it.
// Samples come from HBM with per-lane 128-bit loads (each lane reads its own rows; the two halves of a
// 32-byte sector are consumed by consecutive instructions, so L1 serves the second half), and the next
// tile is prefetched into L2 while the current one is being worked on.
//
// N x K outputs: log-pdfs are staged per warp in shared memory ([KC components][rows], conflict-free) and
// written row-major with lanes running over components, so the global stores are full 32-byte sectors;
// the second pass (rho = exp(lp) w_k / (exp(log q) + tiny), or the VB soft-max) is a separate streaming kernel
// (k1_finish, k1_prepare.cuh) that runs at HBM speed; inside this kernel it cost 3.5 ms of 16.4 at C2 and
// 14 ms of 27 at C3 because its dependent load -> exp -> store chains had only two warps per scheduler to hide behind.
#pragma once

#include "k1_mixture_eval.cuh"

namespace pmc {

constexpr double kFastMaxBias = 1.0e4;   // above this |b| the exact-difference kernel runs instead

struct FastArgs {
  EvalArgs e;             // e.records = derived records (centre slot holds -b_k)
  const double* shift;    // [DP] c
  const int* flag;        // flag[0] != 0: exact-difference form runs; else flag[1] != 0: k1_mma_eval runs; else this one
  double* rowstat;        // [n, 2] per-row (running max, 1/denominator) for k1_finish, or null when no second pass follows
};

template <int DP>
struct FastCfg {
  static constexpr int S = (DP <= 20) ? 4 : (DP <= 32) ? 3 : (DP <= 40) ? 2 : 1;
  static constexpr int NW = (DP <= 12) ? 12 : 8;
  static constexpr int NS = 3;
  static constexpr int KC = (S >= 4) ? 8 : 16;             // components per output staging chunk
  static constexpr int ROWS_PER_WARP = 32 * S;
  static constexpr int RWP = ROWS_PER_WARP + 1;            // padded row count of the staging tile
  static constexpr int TS = ROWS_PER_WARP * NW;
  static constexpr int RL = record_len(DP);
  // staging: per warp 2 tiles (lp, aux) of KC x RWP doubles
  static constexpr size_t STAGE_DOUBLES = size_t(NW) * 2 * KC * RWP;
  static constexpr size_t SMEM_BASE = sizeof(double) * (size_t(NS) * RL + DP) + NS * sizeof(uint64_t) + NS * sizeof(int) + 16;
  static constexpr size_t SMEM_STAGED = SMEM_BASE + sizeof(double) * STAGE_DOUBLES;
};

template <int DP>
__global__ void __launch_bounds__(FastCfg<DP>::NW * 32, 1) k1_fast_eval(const FastArgs fa) {
  using C = FastCfg<DP>;
  constexpr int S = C::S, NW = C::NW, NS = C::NS, RL = C::RL, H = DP / 2, KC = C::KC, RWP = C::RWP;
  constexpr int NT = tri_len(DP);
  constexpr uint32_t REC_BYTES = RL * sizeof(double);
  const EvalArgs& a = fa.e;
  if (fa.flag[0] != 0 || fa.flag[1] != 0) return;                      // another form takes over

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);                  // [NS][RL]
  double* cs = ring + NS * RL;                                         // [DP] shift
  uint64_t* full = reinterpret_cast<uint64_t*>(cs + DP);
  int* empty_cnt = reinterpret_cast<int*>(full + NS);
  double* stage_all = reinterpret_cast<double*>(smem_raw + ((C::SMEM_BASE + 15) & ~size_t(15)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool staged = (a.lp_out != nullptr) || (a.resp_out != nullptr) || (a.aux_out != nullptr);
  double* st_lp = stage_all + size_t(warp) * 2 * KC * RWP;             // [KC][RWP]
  double* st_aux = st_lp + KC * RWP;

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
  for (int j = threadIdx.x; j < DP; j += blockDim.x) cs[j] = (j < a.d) ? fa.shift[j] : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS && s < total_steps; ++s) {
      mbar_arrive_expect_tx(&full[s], REC_BYTES);
      bulk_g2s(ring + s * RL, a.records + size_t(s % a.kl) * RL, REC_BYTES, &full[s]);
    }
  }

  double part_a = 0.0, part_w = 0.0;
  double* const scratch = a.lp_out ? a.lp_out : a.resp_out;
  const bool vec_ok = (a.d == DP) && ((a.ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);

  int64_t step = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * int64_t(gridDim.x);
    const int64_t row0 = tile * C::TS + int64_t(warp) * C::ROWS_PER_WARP;

    // ---- this thread's S samples, shifted, into registers ----
    double xr[S][DP];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t row = row0 + lane + 32 * s;
      if (row < a.n) {
        const double* xp = a.x + row * a.ldx;
        if (vec_ok) {
#pragma unroll
          for (int p = 0; p < H; ++p) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(xp) + p);
            const double2 c = *reinterpret_cast<const double2*>(cs + 2 * p);
            xr[s][2 * p] = v.x - c.x;
            xr[s][2 * p + 1] = v.y - c.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < DP; ++j) xr[s][j] = (j < a.d) ? (__ldg(xp + j) - cs[j]) : 0.0;
        }
      } else {
#pragma unroll
        for (int j = 0; j < DP; ++j) xr[s][j] = 0.0;
      }
    }
    // ---- L2 prefetch of this warp's slice of the CTA's next tile ----
    if (it + 1 < my_tiles) {
      const int64_t nrow0 = row0 + int64_t(gridDim.x) * C::TS;
      if (nrow0 < a.n) {
        const int64_t rows = (a.n - nrow0 < C::ROWS_PER_WARP) ? (a.n - nrow0) : C::ROWS_PER_WARP;
        const char* base = reinterpret_cast<const char*>(a.x + nrow0 * a.ldx);
        const int64_t bytes = rows * a.ldx * int64_t(sizeof(double));
        for (int64_t off = int64_t(lane) * 128; off < bytes; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
      }
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
      const double* nb = rec + NT;            // -b_k
      const double* sc = rec + NT + DP;

      double q[S];
#pragma unroll
      for (int s = 0; s < S; ++s) q[s] = 0.0;
#pragma unroll
      for (int r = 0; r < H; ++r) {
        const double2 b2 = *reinterpret_cast<const double2*>(nb + 2 * r);
        double z0[S], z1[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { z0[s] = b2.x; z1[s] = b2.y; }
#pragma unroll
        for (int p = 0; p <= r; ++p) {
          const double2 t0 = *reinterpret_cast<const double2*>(rec + 2 * r * (r + 1) + 4 * p);
          const double2 t1 = *reinterpret_cast<const double2*>(rec + 2 * r * (r + 1) + 4 * p + 2);
          // snake order over the 2 x S outer product of each column: consecutive DFMAs share T or x alternately, so
          // each needs one new register operand plus its accumulator (the register file delivers ~1 64-bit warp
          // operand per clock per sub-partition, a DFMA wants three: scripts/ubench/dfma_rf.cu, dfma_snake.cu)
#pragma unroll
          for (int s = 0; s < S; ++s) z0[s] = fma(t0.x, xr[s][2 * p], z0[s]);
#pragma unroll
          for (int s = S - 1; s >= 0; --s) z1[s] = fma(t1.x, xr[s][2 * p], z1[s]);
          if (p < r) {                                                 // T[2r][2r+1] == 0
#pragma unroll
            for (int s = 0; s < S; ++s) z0[s] = fma(t0.y, xr[s][2 * p + 1], z0[s]);
#pragma unroll
            for (int s = S - 1; s >= 0; --s) z1[s] = fma(t1.y, xr[s][2 * p + 1], z1[s]);
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) z1[s] = fma(t1.y, xr[s][2 * p + 1], z1[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          q[s] = fma(z0[s], z0[s], q[s]);
          q[s] = fma(z1[s], z1[s], q[s]);
        }
      }

      const double c0 = sc[S0], c1 = sc[S1], c2 = sc[S2], c3 = sc[S3], c4 = sc[S4], wk = sc[S_WEIGHT];
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

      const int kc = kk % KC;
#pragma unroll
      for (int s = 0; s < S; ++s) {
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
        // online weighted log-sum-exp (same value as _regularize.pyx:72-81 up to rounding), branch-free:
        // one exp of -|lp - max| per pair whichever of the two is larger, so lanes never diverge here
        {
          const bool up = lp > run_max[s];
          const double e = exp(-fabs(lp - run_max[s]));
          run_sum[s] = fma(up ? run_sum[s] : wk, e, up ? wk : run_sum[s]);
          run_max[s] = up ? lp : run_max[s];
        }
        if (staged) {
          st_lp[kc * RWP + lane + 32 * s] = lp;
          if (a.aux_out) st_aux[kc * RWP + lane + 32 * s] = aux;
        }
      }

      // ---- flush a full staging chunk row-major: lanes run over the chunk's components ----
      if (staged && (kc == KC - 1 || kk == a.kl - 1)) {
        __syncwarp();
        const int nk = kc + 1, k_base = kk - kc;
        const int l_k = lane % KC, l_r = lane / KC;               // KC components x 32/KC rows per instruction
        if (l_k < nk) {
          const int col = __ldg(a.cols + k_base + l_k);
          for (int r = l_r; r < C::ROWS_PER_WARP; r += 32 / KC) {
            const int64_t row = row0 + r;
            if (row >= a.n) break;
            if (scratch) scratch[row * a.k_out + col] = st_lp[l_k * RWP + r];
            if (a.aux_out) a.aux_out[row * a.k_out + col] = st_aux[l_k * RWP + r];
          }
        }
        __syncwarp();
      }
    }

    // ---- per-sample results ----
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int64_t row = row0 + lane + 32 * s;
      if (row >= a.n) continue;
      const double lq = log(run_sum[s]) + run_max[s];             // _regularize.pyx:81
      const double w_n = a.sw ? __ldg(a.sw + row) : 1.0;
      part_w += w_n;
      if (a.logq) a.logq[row] = lq;
      if (a.mode != MODE_VB) part_a += w_n * lq;                  // pmc.pyx:388-391
      if (fa.rowstat) {
        fa.rowstat[2 * row] = run_max[s];
        fa.rowstat[2 * row + 1] = (a.mode != MODE_VB) ? 1.0 / (exp(lq) + kTiny)   // pmc.pyx:39-41
                                                      : 1.0 / run_sum[s];         // variational.pyx:728-755
      }
    }
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
