// k4_weights.cuh -- K4: importance weights of a run and the weight-vector reductions, one pass, float64, sm_100a.
//
// Replaces (reference loops, /root/reference/pypmc):
//   ImportanceSampler._calculate_weights   sampler/importance_sampling.py:197-215   w_n = exp(log target(x_n) - log q(x_n))
//   perp                                   tools/convergence.py:6-39                needs sum w, sum w log w
//   ess                                    tools/convergence.py:42-72               needs sum w, sum w^2
//   PMC.log_likelihood's weighted mean     mix_adapt/pmc.pyx:388-391                needs sum w log q
// The reference makes one numpy pass per quantity over the N-vector (and the per-sample Python loop for the weights);
// here one streaming kernel writes w and leaves the five sums, so perp / ess of a run cost no pass of their own.
// HBM-bound: 24 bytes per sample (read log target, log q; write w).  Per-CTA partials, reduced in CTA order by the
// last CTA to finish (fixed summation order: run-to-run reproducible for a given grid).
#pragma once

#include "pmc_common.cuh"

namespace pmc {

constexpr int K4_THREADS = 256;
constexpr int K4_SUMS = 5;   // sum w | sum w log q | sum w^2 | sum w log w | number of nonzero weights

__global__ void __launch_bounds__(K4_THREADS) k4_weights(const double* __restrict__ log_target, const double* __restrict__ logq,
                                                         int64_t n, double* __restrict__ w_out, double* __restrict__ partials,
                                                         unsigned int* __restrict__ counter, double* __restrict__ sums) {
  double acc[K4_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double lq = logq[i];
    const double lw = log_target ? log_target[i] - lq : lq;     // without a target: logq already holds log w
    const double w = exp(lw);                                    // importance_sampling.py:209-215
    if (w_out) w_out[i] = w;
    acc[0] += w;
    acc[2] += w * w;
    if (w != 0.0) {                                              // zero weights contribute nothing (convergence.py:30-34)
      acc[1] += w * lq;
      acc[3] += w * lw;
      acc[4] += 1.0;
    }
  }
  __shared__ double red[K4_THREADS / 32][K4_SUMS];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < K4_SUMS; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
    if (lane == 0) red[warp][s] = acc[s];
  }
  __syncthreads();
  if (threadIdx.x < K4_SUMS) {
    double t = 0.0;
    for (int w = 0; w < K4_THREADS / 32; ++w) t += red[w][threadIdx.x];
    partials[size_t(blockIdx.x) * K4_SUMS + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < K4_SUMS) {
      double t = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) t += partials[size_t(b) * K4_SUMS + threadIdx.x];
      sums[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *counter = 0u;                        // clean for the next launch
  }
}

}  // namespace pmc
