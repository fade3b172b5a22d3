// k1_prepare.cuh -- the one-CTA prologue of every K1 launch (included by pmcb200.cu only).
#pragma once
#include "k1_fast_eval.cuh"

namespace pmc {

// ---------------------------------------------------------------------------------------------
// k1_prepare: one CTA.  c = sum_k w_k mu_k / sum_k w_k (plain mean if the weights are unusable),
// derived record k = [T_k | -T_k (mu_k - c) | scalars], flag = (max |b| > kFastMaxBias or non-finite).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) k1_prepare(const double* __restrict__ records, int kl, int dp,
                                                     double* __restrict__ derived, double* __restrict__ shift,
                                                     int* __restrict__ flag, double* __restrict__ partials, int n_partials) {
  const int nt = tri_len(dp), rl = record_len(dp);
  __shared__ double c_s[PMC_MAX_DP];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  for (int i = threadIdx.x; i < n_partials; i += blockDim.x) partials[i] = 0.0;
  for (int j = threadIdx.x; j < dp; j += blockDim.x) {
    double sw = 0.0, sm = 0.0, su = 0.0;
    for (int k = 0; k < kl; ++k) {                       // fixed order
      const double w = records[size_t(k) * rl + nt + dp + S_WEIGHT];
      const double m = records[size_t(k) * rl + nt + j];
      sw += w; sm += w * m; su += m;
    }
    double c = sm / sw;
    if (!(sw > 0.0) || !isfinite(c)) c = su / kl;
    if (!isfinite(c)) c = 0.0;
    c_s[j] = c;
    shift[j] = c;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kl * rl; e += blockDim.x) {
    const int k = e / rl, o = e - k * rl;
    const double* rec = records + size_t(k) * rl;
    double v = rec[o];
    if (o >= nt && o < nt + dp) {                        // centre slot -> -b_i,  i = o - nt
      const int i = o - nt, r = i >> 1;
      double b = 0.0;
      for (int j = 0; j <= (i | 1) && j < dp; ++j) {     // row i of T: 2x2 blocks (r, p = j/2)
        const double t = rec[2 * r * (r + 1) + 4 * (j >> 1) + 2 * (i & 1) + (j & 1)];
        b = fma(t, rec[nt + j] - c_s[j], b);
      }
      if (!(fabs(b) <= kFastMaxBias)) bad = 1;           // also catches NaN
      v = -b;
    }
    derived[e] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) *flag = bad;
}

}  // namespace pmc
