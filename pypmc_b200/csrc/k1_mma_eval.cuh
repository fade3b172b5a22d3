// k1_mma_eval.cuh -- K1, matrix-instruction form: the same fused per-sample x per-component log-pdf + mixture
// log-sum-exp as k1_fast_eval.cuh (same outputs, same reference citations: gauss.pyx:146-151, student_t.pyx:154-164,
// mixture.pyx:112-156, _regularize.pyx:57-83, pmc.pyx:23-43, variational.pyx:774-798), issued as FP64 matrix
// instructions (mma.sync.m8n8k4.f64, SASS DMMA).
//
// Why a third form: the DFMA forms read three 64-bit register operands per FMA and the register file delivers about
// one per clock, which caps them at 66-83 % of the FP64 pipe (profiles/r01_operand_delivery.md; k1_fast_eval measured
// 70 %).  A DMMA moves 256 FMAs with four operand registers per thread (K2 reached 87 % that way).  The triangular
// solve itself maps badly on 8x4 blocks (+25 % FMAs), so the quadratic form is expanded instead:
//     q_nk = (x' - d_k)^T M_k (x' - d_k) = theta_k . phi(x'),      x' = x - c,  d_k = mu_k - c,  M_k = T_k^T T_k
//     phi(x')  = [ 1 | x'_i | x'_i x'_j (j <= i) ]                 F = 1 + D + D(D+1)/2 features (496 at D=30)
//     theta_k  = [ |b_k|^2 | -2 (T_k^T b_k)_i | M_ii, 2 M_ij ]     b_k = T_k d_k       (k1_mma_prepare)
// i.e. a dense (N x F).(F x K) product with F (D(D+1)/2 + D + 1) FMAs per pair -- the algorithmic count, nothing
// padded but the last feature quad -- whose left operand is formed on the fly (one DMUL per fragment).
// Rounding: the terms are of size |b_k|^2 where the result may be small, so the absolute error of q is about
// sqrt(F) eps |b_k|^2; k1_prepare allows this form only while max_k |b_k|^2 <= kMmaMaxBias2 (error of q below ~1e-11),
// otherwise k1_fast_eval (|b| <= 1e4) or the exact-difference form run -- all decided on the device, no host sync.
//
// Mapping: persistent CTAs (one per SM), NW = 16 warps (four per scheduler: a warp issues at most one DMMA per ~32 clk,
// the pipe takes one per 16, and a warp in its epilogue issues none).  theta for ALL components of the launch stays in
// shared memory for the whole kernel ([feature quad][component][4], 127 KB at K=32, D=30), so there is no ring and no
// barrier in the sample loop; mixtures whose theta does not fit are evaluated in component groups (MmaArgs below).
// A warp owns 8 NB samples per tile: its private shared-memory slice is filled by cp.async with the rows of the NEXT
// tile while the epilogue of the current one runs (x - c in place, plus a constant column 1 and a zero column), then
// per feature quad it loads CB theta fragments, forms NB phi fragments (2 LDS + DMUL each) and
// issues NB x CB DMMAs into 8x8 (sample x component) accumulators.  A sample's K log-pdfs end up inside one quad of
// lanes, so the log-sum-exp is two shuffles deep and the N x K outputs leave as 16-byte stores, two full sectors per
// quad.  The second pass (rho = exp(lp) w_k / (exp(log q) + tiny), or the VB soft-max with its sum r log r) happens
// here too, on the log-pdfs still in registers -- k1_finish, the streaming pass behind the DFMA forms, returns at once.
#pragma once

#include "k1_fast_eval.cuh"

namespace pmc {

constexpr double kMmaMaxBias2 = 2.0e4;   // above this |b_k|^2 the DFMA forms run instead
constexpr int K1M_SCAL = 8;              // scalars per component kept in shared memory

struct MmaArgs {
  EvalArgs e;             // e.records = derived records ([T | -b | scalars]); only the scalars are read here
  const double* theta;    // [steps][KP][4]
  const double* shift;    // [d]
  const int* flag;        // flag[0] != 0: exact-difference form runs; else flag[1] != 0: this form runs; else k1_fast_eval
  int steps;              // feature quads = ceil(F / 4)
  int KP;                 // components (of this group) padded to 8 CB
  int YS;                 // row stride (doubles) of the staged samples: >= d + 2 and == 4 (mod 16)
  // Component groups: when theta of all components does not fit shared memory, the components are evaluated in
  // `ngroups` launches of at most KP each (e.records / e.cols / e.kl / theta describe THIS group).  The running
  // (max, weighted sum) of the log-sum-exp travels between the launches in rowstat; the last launch finishes log q
  // and leaves (max, 1 / denominator) there for k1_finish, which then does the second pass (SECOND == false).
  int group, ngroups;
  double* rowstat;        // [n, 2]; null when ngroups == 1 and no k1_finish pass follows
};

__host__ __device__ inline int k1m_features(int d) { return 1 + d + d * (d + 1) / 2; }
__host__ __device__ inline int k1m_row_stride(int d) { return ((d + 2 - 4 + 15) / 16) * 16 + 4; }
inline size_t k1m_smem_bytes(int d, int KP, int NB, int NW) {
  const int steps = (k1m_features(d) + 3) / 4, YS = k1m_row_stride(d);
  return sizeof(double) * (size_t(steps) * KP * 4 + size_t(KP) * K1M_SCAL + YS + size_t(NW) * 8 * NB * YS) +
         sizeof(int) * size_t(steps) * 4 + 16;
}

// lower-triangle index t -> (row, col), row-major
__device__ __forceinline__ void tri_index(int t, int& r, int& c) {
  r = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  c = t - r * (r + 1) / 2;
}

// ---------------------------------------------------------------------------------------------
// k1_mma_eval<CB, NB, NW, SECOND>: CB blocks of 8 components (KP = 8 CB), NB blocks of 8 samples per warp and tile,
// NW warps per CTA; SECOND: the launch also wants rho / r (the eval-only instantiation keeps no exponentials).
// ---------------------------------------------------------------------------------------------
template <int CB, int NB, int NW, bool SECOND>
__global__ void __launch_bounds__(NW * 32, 1) k1_mma_eval(const MmaArgs ma) {
  const EvalArgs& a = ma.e;
  if (ma.flag[0] != 0 || ma.flag[1] == 0) return;
  constexpr int RW = 8 * NB, TS = RW * NW, KP = 8 * CB;
  constexpr int UNR = (CB * NB <= 4) ? 4 : (CB * NB >= 12) ? 1 : 2;   // feature quads per loop body (about 16 DMMAs)
  const int D = a.d, YS = ma.YS, steps = ma.steps;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* theta_s = reinterpret_cast<double*>(smem_raw);               // [steps][KP][4]
  double* scal_s = theta_s + size_t(steps) * KP * 4;                   // [KP][8]
  double* cs = scal_s + KP * K1M_SCAL;                                 // [YS] shift
  double* y_all = cs + YS;                                             // [NW][RW][YS]
  int* tab = reinterpret_cast<int*>(y_all + size_t(NW) * RW * YS); // [steps * 4] (off_i | off_j << 8)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;

  // ---- prologue: theta, scalars, shift, feature table, constant columns ----
  {
    const int n2 = steps * KP * 2;
    const double2* src = reinterpret_cast<const double2*>(ma.theta);
    double2* dst = reinterpret_cast<double2*>(theta_s);
    for (int i = tid; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
    const int dp = (D + 1) & ~1, rl = record_len(dp), so = tri_len(dp) + dp;
    for (int i = tid; i < KP * K1M_SCAL; i += blockDim.x) {
      const int k = i / K1M_SCAL, s = i - k * K1M_SCAL;
      scal_s[i] = (k < a.kl) ? a.records[size_t(k) * rl + so + s] : 0.0;
    }
    for (int j = tid; j < YS; j += blockDim.x) cs[j] = (j < D) ? ma.shift[j] : 0.0;
    const int F = k1m_features(D);
    for (int f = tid; f < steps * 4; f += blockDim.x) {
      int oi = D + 1, oj = D + 1;                                       // zero column
      if (f == 0) { oi = D; oj = D; }
      else if (f <= D) { oi = f - 1; oj = D; }
      else if (f < F) tri_index(f - 1 - D, oi, oj);
      tab[f] = oi | (oj << 8);
    }
    for (int i = tid; i < NW * RW * YS; i += blockDim.x) y_all[i] = ((i % YS) == D) ? 1.0 : 0.0;
  }
  // contiguous output columns (the usual case) allow 16-byte stores
  const bool staged = a.lp_out != nullptr || a.resp_out != nullptr || a.aux_out != nullptr;
  int contig_l = 1;
  if (staged) {
    const int c0 = __ldg(a.cols);
    for (int k = tid; k < a.kl; k += blockDim.x) contig_l &= (__ldg(a.cols + k) == c0 + k);
    contig_l &= ((c0 & 1) == 0) && ((a.k_out & 1) == 0) && ((a.kl & 1) == 0);
    contig_l &= (a.lp_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.lp_out) & 15) == 0);
    contig_l &= (a.resp_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.resp_out) & 15) == 0);
    contig_l &= (a.aux_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.aux_out) & 15) == 0);
  }
  const bool contig = __syncthreads_and(contig_l) != 0;               // also publishes the prologue's stores
  const int col0 = staged ? __ldg(a.cols) : 0;

  double* yw = y_all + size_t(warp) * RW * YS;
  const double* yg = yw + g * YS;                                      // rows g + 8 nb
  const double* thl = theta_s + g * 4 + tq;                            // theta[s][8 cb + g][tq]
  const int* tabl = tab + tq;

  const int64_t num_tiles = (a.n + TS - 1) / TS;
  double part_a = 0.0, part_w = 0.0;

  // this warp's rows of a tile -> its shared-memory slice, asynchronously (LDGSTS, lane = column); rows beyond n
  // are zero-filled.  The shift is applied in place when the tile is picked up.
  auto stage_rows = [&](int64_t r0) {
    for (int j = lane; j < D; j += 32) {
      const uint32_t dst = smem_u32(yw + j);
#pragma unroll 8
      for (int r = 0; r < RW; ++r) {
        const int64_t row = r0 + r;
        const bool in = row < a.n;
        const double* src = in ? (a.x + row * a.ldx + j) : a.x;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + uint32_t(r * YS) * 8u), "l"(src),
                     "r"(in ? 8 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (int64_t(blockIdx.x) < num_tiles) stage_rows(int64_t(blockIdx.x) * TS + int64_t(warp) * RW);
  // The two warps of a scheduler would run in lock step: both in the DMMA loop (sharing the pipe), then both in the
  // latency-bound epilogue (pipe idle, 18 % of the time in the first profile).  Starting the second warp one
  // DMMA-loop later keeps them in opposite phases for the whole kernel: one warp's epilogue hides behind the other's loop.
  if (warp >= 4 && num_tiles >= 4 * int64_t(gridDim.x)) {
    const long long t0 = clock64(), wait = (long long)steps * (NB * CB * 16) * (warp / 4);
    while (clock64() - t0 < wait) {
    }
  }

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TS + int64_t(warp) * RW;
    // ---- pick up the staged rows: y = x - c in place (each lane owns the columns it copied) ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int j = lane; j < D; j += 32) {
      const double c = cs[j];
#pragma unroll 8
      for (int r = 0; r < RW; ++r) yw[r * YS + j] -= c;
    }
    __syncwarp();

    double acc[NB][CB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }

    // ---- q = phi . theta over the feature quads; the operands of quad s + 1 are fetched before the DMMAs of quad s ----
    double th_n[CB], yi_n[NB], yj_n[NB];
    auto fetch = [&](int s) {
      const int t = tabl[4 * s];
      const double* yi = yg + (t & 0xff);
      const double* yj = yg + (t >> 8);
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) th_n[cb] = thl[(s * KP + cb * 8) * 4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) { yi_n[nb] = yi[nb * 8 * YS]; yj_n[nb] = yj[nb * 8 * YS]; }
    };
    fetch(0);
#pragma unroll UNR
    for (int s = 0; s < steps; ++s) {
      double th[CB], ph[NB];
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) th[cb] = th_n[cb];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) ph[nb] = yi_n[nb] * yj_n[nb];
      fetch(min(s + 1, steps - 1));
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[nb][cb][0]), "+d"(acc[nb][cb][1])
                       : "d"(ph[nb]), "d"(th[cb]));
    }
    // ---- the slice is free: start copying this warp's rows of the CTA's next tile behind the epilogue ----
    __syncwarp();
    if (tile + gridDim.x < num_tiles) stage_rows(row0 + int64_t(gridDim.x) * TS);

    // ---- epilogue: lane holds q[sample 8 nb + g][component 8 cb + 2 tq + e] ----
    // Three phases over the whole tile (log-pdfs and their stores; maxima; exponentials and sums), so that the
    // NB x CB x 2 exponentials of a lane are independent instruction streams the scheduler can interleave.
    double mx[NB], m_prev[NB], s_prev[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      mx[nb] = a.max_init;
      m_prev[nb] = a.max_init;
      s_prev[nb] = 0.0;
      if (ma.group > 0) {                                               // running (max, sum) of the earlier groups
        const int64_t row = row0 + 8 * nb + g;
        if (row < a.n) {
          m_prev[nb] = ma.rowstat[2 * row];
          s_prev[nb] = ma.rowstat[2 * row + 1];
          mx[nb] = m_prev[nb];
        }
      }
    }
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * cb + 2 * tq + e;
        const double* sc = scal_s + k * K1M_SCAL;
        const bool pad = k >= a.kl;
        if (a.mode == MODE_GAUSS) {
          const double c0 = sc[S0];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const double l = c0 - 0.5 * fmax(acc[nb][cb][e], 0.0);                       // gauss.pyx:151
            acc[nb][cb][e] = pad ? -INFINITY : l;
          }
        } else if (a.mode == MODE_STUDENT_T) {
          const double c0 = sc[S0], c1 = sc[S1], c2 = sc[S2], c3 = sc[S3], c4 = sc[S4];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const double q = fmax(acc[nb][cb][e], 0.0);
            double t = q * c2;                                                           // student_t.pyx:159-164
            t += 1.0;
            t = log(t);
            t *= c1;
            acc[nb][cb][e] = pad ? -INFINITY : t + c0;
            if (a.aux_out && !pad && row0 + 8 * nb + g < a.n)                            // gamma_nk, pmc.pyx:610
              a.aux_out[size_t(row0 + 8 * nb + g) * a.k_out + (contig ? col0 + k : __ldg(a.cols + k))] = c4 / (c3 + q);
          }
        } else {
          const double c0 = sc[S0], c1 = sc[S1], c2 = sc[S2], c3 = sc[S3], c4 = sc[S4];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const double x = c3 + c4 * fmax(acc[nb][cb][e], 0.0);                        // variational.pyx:798
            acc[nb][cb][e] = pad ? -INFINITY : c0 + 0.5 * (c1 - c2 - x);                 // variational.pyx:691
            if (a.aux_out && !pad && row0 + 8 * nb + g < a.n)
              a.aux_out[size_t(row0 + 8 * nb + g) * a.k_out + (contig ? col0 + k : __ldg(a.cols + k))] = x;
          }
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) mx[nb] = fmax(mx[nb], acc[nb][cb][e]);
      }
    // log-pdfs leave as 16-byte stores when the output columns are contiguous (VB: log_rho is written normalised below)
    auto store_pairs = [&](double* out, const double (&v)[NB][CB][2]) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int64_t row = row0 + 8 * nb + g;
        if (row >= a.n) continue;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const int k = 8 * cb + 2 * tq;
          if (contig) {
            if (k < a.kl) *reinterpret_cast<double2*>(out + size_t(row) * a.k_out + col0 + k) = make_double2(v[nb][cb][0], v[nb][cb][1]);
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (k + e < a.kl) out[size_t(row) * a.k_out + __ldg(a.cols + k + e)] = v[nb][cb][e];
          }
        }
      }
    };
    if constexpr (SECOND) {
      if (a.lp_out && a.mode != MODE_VB) store_pairs(a.lp_out, acc);
    } else {
      // no fused second pass: like the DFMA forms, raw log-pdfs go to lp_out, or wait in resp_out for k1_finish
      double* const scratch = a.lp_out ? a.lp_out : a.resp_out;
      if (scratch) store_pairs(scratch, acc);
    }
    // weighted log-sum-exp over the components (same value as _regularize.pyx:72-81 up to rounding)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      mx[nb] = fmax(mx[nb], __shfl_xor_sync(0xffffffffu, mx[nb], 1));
      mx[nb] = fmax(mx[nb], __shfl_xor_sync(0xffffffffu, mx[nb], 2));
    }
    double ex[SECOND ? NB : 1][CB][2], sum[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) sum[nb] = 0.0;
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double wk = scal_s[(8 * cb + 2 * tq + e) * K1M_SCAL + S_WEIGHT];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const double t = wk * exp(acc[nb][cb][e] - mx[nb]);
          if constexpr (SECOND) ex[nb][cb][e] = t;
          sum[nb] += t;
        }
      }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      sum[nb] += __shfl_xor_sync(0xffffffffu, sum[nb], 1);
      sum[nb] += __shfl_xor_sync(0xffffffffu, sum[nb], 2);
      if (ma.group > 0) sum[nb] = fma(s_prev[nb], exp(m_prev[nb] - mx[nb]), sum[nb]);
    }
    // ---- per-sample results: lane tq == nb % 4 of the quad finishes sample nb and hands the quad what the
    //      second pass needs (the DFMA forms leave that pass to k1_finish; here the log-pdfs are still in registers) ----
    constexpr bool second = SECOND;     // host: (resp_out != null) || (mode == VB && lp_out != null)
    double f0[NB], f1[NB];          // mixtures: 1 / (exp(log q) + tiny), exp(max) * that;  VB: 1 / norm, ln(1 / norm)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int64_t row = row0 + 8 * nb + g;
      f0[nb] = 0.0;
      f1[nb] = 0.0;
      if (tq == (nb & 3)) {
        if (!SECOND && ma.group + 1 < ma.ngroups) {                     // more components to come
          if (row < a.n) {
            ma.rowstat[2 * row] = mx[nb];
            ma.rowstat[2 * row + 1] = sum[nb];
          }
        } else {
          const double lq = log(sum[nb]) + mx[nb];                      // _regularize.pyx:81
          if (row < a.n) {
            const double w_n = a.sw ? __ldg(a.sw + row) : 1.0;
            part_w += w_n;
            if (a.logq) a.logq[row] = lq;
            if (a.mode != MODE_VB) part_a += w_n * lq;                  // pmc.pyx:388-391
            if (!SECOND && ma.rowstat) {                                // for k1_finish
              ma.rowstat[2 * row] = mx[nb];
              ma.rowstat[2 * row + 1] = (a.mode != MODE_VB) ? 1.0 / (exp(lq) + kTiny)   // pmc.pyx:39-41
                                                            : 1.0 / sum[nb];            // variational.pyx:728-755
            }
          }
          if (second) {
            if (a.mode != MODE_VB) {
              f0[nb] = 1.0 / (exp(lq) + kTiny);                         // pmc.pyx:39-41
              f1[nb] = exp(mx[nb]) * f0[nb];
            } else {
              f0[nb] = 1.0 / sum[nb];                                   // variational.pyx:728-755
              f1[nb] = log(f0[nb]);
            }
          }
        }
      }
      if (second) {
        f0[nb] = __shfl_sync(0xffffffffu, f0[nb], (lane & ~3) | (nb & 3));
        f1[nb] = __shfl_sync(0xffffffffu, f1[nb], (lane & ~3) | (nb & 3));
      }
    }
    if constexpr (SECOND) {
      if (a.mode != MODE_VB) {
        // rho_nk = exp(lp_nk) w_k / (exp(log q_n) + tiny) = [w_k exp(lp_nk - max)] [exp(max) / (exp(log q_n) + tiny)];
        // where exp(lp_nk) is subnormal the reference's own rounding is reproduced by its literal formula
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              double r = ex[nb][cb][e] * f1[nb];
              if (acc[nb][cb][e] < -700.0 || mx[nb] > 700.0)
                r = exp(acc[nb][cb][e]) * scal_s[(8 * cb + 2 * tq + e) * K1M_SCAL + S_WEIGHT] * f0[nb];
              ex[nb][cb][e] = r;
            }
        store_pairs(a.resp_out, ex);
      } else {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int64_t row = row0 + 8 * nb + g;
          const double w_r = (a.sw && row < a.n) ? __ldg(a.sw + row) : 1.0;
          double s_rl = 0.0;
#pragma unroll
          for (int cb = 0; cb < CB; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double rv = ex[nb][cb][e] * f0[nb];
              if (rv == 0.0) rv = kTiny;                                              // variational.pyx:753-754
              const double lrn = (acc[nb][cb][e] - mx[nb]) + f1[nb];                  // variational.pyx:741,755
              ex[nb][cb][e] = rv;
              acc[nb][cb][e] = lrn;
              if (8 * cb + 2 * tq + e < a.kl) s_rl = fma(w_r * rv, lrn, s_rl);        // variational.pyx:1003-1013
            }
          if (row < a.n) part_a += s_rl;
        }
        if (a.resp_out) store_pairs(a.resp_out, ex);
        if (a.lp_out) store_pairs(a.lp_out, acc);
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
