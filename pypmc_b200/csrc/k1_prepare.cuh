// k1_prepare.cuh -- the one-CTA prologue of every K1 launch (included by pmcb200.cu only).
#pragma once
#include "k1_mma_eval.cuh"

namespace pmc {

// ---------------------------------------------------------------------------------------------
// k1_prepare: one CTA per evaluated component (a single CTA took 62 us at K=32, D=30: 0.6 % of a pass at N=1e7 and
// most of a small one).  Every CTA forms the shift c = sum_k w_k mu_k / sum_k w_k (plain mean if the weights are
// unusable) in the same fixed order, then its derived record [T_k | -T_k (mu_k - c) | scalars].  The last CTA to
// finish combines the per-component verdicts: flag[0] = (max |b| > kFastMaxBias or non-finite): the
// exact-difference form runs; flag[1] = (want_mma and max_k |b_k|^2 <= kMmaMaxBias2): the matrix-instruction form
// runs (k1_mma_eval.cuh); neither: k1_fast_eval.  flag[2] (arrival counter) and flag[3] (verdict bits) are zero
// between launches (cleared at allocation and by the last CTA).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k1_prepare(const double* __restrict__ records, int kl, int dp,
                                                  double* __restrict__ derived, double* __restrict__ shift,
                                                  int* __restrict__ flag, double* __restrict__ partials, int n_partials,
                                                  int want_mma) {
  const int nt = tri_len(dp), rl = record_len(dp), k = blockIdx.x;
  __shared__ double c_s[PMC_MAX_DP], b_s[PMC_MAX_DP];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_partials; i += gridDim.x * blockDim.x) partials[i] = 0.0;
  for (int j = threadIdx.x; j < dp; j += blockDim.x) {
    double sw = 0.0, sm = 0.0, su = 0.0;
    for (int kk = 0; kk < kl; ++kk) {                    // fixed order
      const double w = records[size_t(kk) * rl + nt + dp + S_WEIGHT];
      const double m = records[size_t(kk) * rl + nt + j];
      sw += w; sm += w * m; su += m;
    }
    double c = sm / sw;
    if (!(sw > 0.0) || !isfinite(c)) c = su / kl;
    if (!isfinite(c)) c = 0.0;
    c_s[j] = c;
    if (k == 0) shift[j] = c;
  }
  __syncthreads();
  const double* rec = records + size_t(k) * rl;
  for (int o = threadIdx.x; o < rl; o += blockDim.x) {
    double v = rec[o];
    if (o >= nt && o < nt + dp) {                        // centre slot -> -b_i,  i = o - nt
      const int i = o - nt, r = i >> 1;
      double b = 0.0;
      for (int j = 0; j <= (i | 1) && j < dp; ++j) {     // row i of T: 2x2 blocks (r, p = j/2)
        const double t = rec[2 * r * (r + 1) + 4 * (j >> 1) + 2 * (i & 1) + (j & 1)];
        b = fma(t, rec[nt + j] - c_s[j], b);
      }
      if (!(fabs(b) <= kFastMaxBias)) bad = 1;           // also catches NaN
      b_s[i] = b;
      v = -b;
    }
    derived[size_t(k) * rl + o] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < dp; ++i) s = fma(b_s[i], b_s[i], s);
    const int bits = (bad ? 1 : 0) | (!(s <= kMmaMaxBias2) ? 2 : 0);
    if (bits) atomicOr(&flag[3], bits);
    __threadfence();
    if (atomicAdd(&flag[2], 1) == int(gridDim.x) - 1) {  // last CTA: every verdict is in
      __threadfence();
      const int all = atomicExch(&flag[3], 0);
      flag[0] = all & 1;
      flag[1] = (want_mma && all == 0) ? 1 : 0;
      flag[2] = 0;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// k1_mma_prepare: one CTA per (padded) component: theta_k from the derived record [T | -b | scalars], and the last
// word on whether the matrix-instruction form may run.  The expanded quadratic form cancels terms of size
//     A_k = sum_f |theta_kf| |phi_f(d_k)| = |b_k|^2 + 2 sum_i |(T^T b)_i| |d_i| + sum_{j<=i} c_ij |M_ij| |d_i| |d_j|
// (d_k = mu_k - c) where q may be O(D); A_k equals |b_k|^2 for a well-conditioned component but grows to kappa |b_k|^2
// when the offset lies along a wide axis of an ill-conditioned covariance -- k1_prepare's |b_k|^2 test cannot see that.
// The absolute error of q is ~ sqrt(F) eps A_k, and q enters the exponent with the factor s_k = 1/2 (Gauss),
// (nu + D) / (2 nu) at most (Student-t), nu_k / 2 (VB, where T^T T = W_k only): the form is refused (flag[1] <- 0, the
// DFMA form k1_fast_eval takes the launch) unless 2 s_k A_k <= kMmaMaxBias2 for every component.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k1_mma_prepare(const double* __restrict__ derived, const double* __restrict__ records,
                                                      const double* __restrict__ shift, int kl, int KP, int d, int dp,
                                                      int steps, int mode, double* __restrict__ theta_all, int* __restrict__ flag) {
  if (flag[0] != 0 || flag[1] == 0) return;
  // block b = (component group, slot in the group); theta of group g is [steps / 2][KP][4][2] (k1m_theta_index) at g * steps * KP * 4
  const int grp = blockIdx.x / KP, slot = blockIdx.x - grp * KP, k = grp * KP + slot, tid = threadIdx.x;
  double* theta = theta_all + size_t(grp) * steps * KP * 4;
  const int nt = tri_len(dp), rl = record_len(dp), F = k1m_features(d);
  if (k >= kl) {                                                        // padding components: theta = 0
    for (int f = tid; f < steps * 4; f += blockDim.x) theta[k1m_theta_index(f, KP, slot)] = 0.0;
    return;
  }
  __shared__ double Ts[PMC_MAX_DP * PMC_MAX_DP];                        // dense T, row stride dp
  __shared__ double bs[PMC_MAX_DP], gs[PMC_MAX_DP], ds[PMC_MAX_DP];
  __shared__ double b2, red[8];
  const double* rec = derived + size_t(k) * rl;
  for (int e = tid; e < dp * dp; e += blockDim.x) {
    const int i = e / dp, j = e - i * dp;
    Ts[e] = (j <= i) ? rec[2 * (i >> 1) * ((i >> 1) + 1) + 4 * (j >> 1) + 2 * (i & 1) + (j & 1)] : 0.0;
  }
  for (int i = tid; i < dp; i += blockDim.x) {
    bs[i] = -rec[nt + i];
    ds[i] = (i < d) ? fabs(records[size_t(k) * rl + nt + i] - shift[i]) : 0.0;   // |d_i|
  }
  __syncthreads();
  for (int i = tid; i < d; i += blockDim.x) {                           // g = T^T b
    double g = 0.0;
    for (int r = i; r < d; ++r) g = fma(Ts[r * dp + i], bs[r], g);
    gs[i] = g;
  }
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s = fma(bs[i], bs[i], s);
    b2 = s;
  }
  __syncthreads();
  double size = 0.0;                                                    // this thread's share of A_k
  for (int f = tid; f < steps * 4; f += blockDim.x) {
    double v = 0.0, ph = 0.0;
    if (f == 0) { v = b2; ph = 1.0; }
    else if (f <= d) { v = -2.0 * gs[f - 1]; ph = ds[f - 1]; }
    else if (f < F) {
      int r, c;
      tri_index(f - 1 - d, r, c);
      double m = 0.0;
      for (int u = r; u < d; ++u) m = fma(Ts[u * dp + r], Ts[u * dp + c], m);     // (T^T T)_rc, c <= r
      v = (r == c) ? m : 2.0 * m;
      ph = ds[r] * ds[c];
    }
    theta[k1m_theta_index(f, KP, slot)] = v;
    size = fma(fabs(v), ph, size);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) size += __shfl_xor_sync(0xffffffffu, size, o);
  if ((tid & 31) == 0) red[tid >> 5] = size;
  __syncthreads();
  if (tid == 0) {
    double A = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) A += red[w];
    const double* sc = rec + nt + dp;
    const double s2 = (mode == MODE_GAUSS) ? 1.0 : (mode == MODE_STUDENT_T) ? sc[S4] / sc[S3] : sc[S4];   // 2 s_k
    if (!(s2 * A <= kMmaMaxBias2)) atomicExch(&flag[1], 0);             // also refuses NaN
  }
}

// ---------------------------------------------------------------------------------------------
// k1_finish: second pass of K1 as a streaming kernel (runs iff the fast form did: flag[0] == 0 and flag[1] == 0).
//   mixture modes : rho_nk = exp(lp_nk) w_k / (exp(log q_n) + tiny)                 pmc.pyx:39-41
//   VB mode       : r_nk = exp(lp - max) / norm (zeros -> tiny), log_rho <- lp - max + ln(1/norm),
//                   partial sums of w_n r_nk log r_nk                               variational.pyx:728-755, 1003-1013
// `scratch` holds lp_nk (written row-major by k1_fast_eval), rowstat the per-row (max, 1/denominator).
// A warp covers 32/hw rows per step with hw lanes running over the evaluated components (coalesced),
// four steps in flight.  Per-block partial sums are written in block order (deterministic).
// ---------------------------------------------------------------------------------------------
struct FinishArgs {
  int64_t n;
  int kl, k_out, mode, rl, w_off;   // record length and offset of the weight scalar inside a record
  const double* records;
  const int* cols;
  const double* rowstat;
  const double* sw;
  double* scratch;
  double* lp_out;
  double* resp_out;
  const int* flag;
  double* fin_partials;             // [gridDim.x] or null
  int mma_fused;                    // the matrix-instruction form, if it ran, did the second pass itself
};

__global__ void __launch_bounds__(256, 4) k1_finish(const FinishArgs a) {
  if (a.flag[0] != 0 || (a.flag[1] != 0 && a.mma_fused)) return;   // exact-difference form / fused DMMA form: pass already done
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hw = (a.kl > 16) ? 32 : (a.kl > 8) ? 16 : (a.kl > 4) ? 8 : 4;
  const int per = 32 / hw, l_k = lane % hw, l_r = lane / hw;
  constexpr int U = 4;
  const int64_t rows_per_step = int64_t(per) * U;
  const int64_t gwarp = int64_t(blockIdx.x) * (blockDim.x >> 5) + warp, nwarps = int64_t(gridDim.x) * (blockDim.x >> 5);
  double acc = 0.0;
  for (int kb = 0; kb < a.kl; kb += hw) {
    const int kk = kb + l_k;
    const bool live = kk < a.kl;
    const int col = live ? __ldg(a.cols + kk) : 0;
    const double wk = live ? __ldg(a.records + size_t(kk) * a.rl + a.w_off) : 0.0;
    for (int64_t r0 = gwarp * rows_per_step; r0 < a.n; r0 += nwarps * rows_per_step) {
      double lpv[U], mx[U], dv[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = r0 + u * per + l_r;
        ok[u] = live && row < a.n;
        lpv[u] = ok[u] ? a.scratch[size_t(row) * a.k_out + col] : 0.0;
        mx[u] = ok[u] ? __ldg(a.rowstat + 2 * row) : 0.0;
        dv[u] = ok[u] ? __ldg(a.rowstat + 2 * row + 1) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const int64_t row = r0 + u * per + l_r;
        const size_t o = size_t(row) * a.k_out + col;
        if (a.mode != MODE_VB) {
          a.resp_out[o] = exp(lpv[u]) * wk * dv[u];
        } else {
          const double lr = lpv[u] - mx[u];
          double rv = exp(lr) * dv[u];
          if (rv == 0.0) rv = kTiny;
          const double lrn = lr + log(dv[u]);
          if (a.resp_out) a.resp_out[o] = rv;
          if (a.lp_out) a.lp_out[o] = lrn;
          const double w_r = a.sw ? __ldg(a.sw + row) : 1.0;
          acc = fma(w_r * rv, lrn, acc);
        }
      }
    }
  }
  if (a.fin_partials) {
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < int(blockDim.x >> 5); ++w) s += red[w];
      a.fin_partials[blockIdx.x] = s;
    }
  }
}

// sums[0] += sum_b fin_partials[b] in block order (VB: sum_n w_n sum_k r log r)
__global__ void k1_reduce_finish(const double* __restrict__ fin_partials, int count, const int* __restrict__ flag,
                                 int mma_fused, double* __restrict__ sums) {
  if (flag[0] != 0 || (flag[1] != 0 && mma_fused) || threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < count; ++i) s += fin_partials[i];
  sums[0] += s;
}

}  // namespace pmc
