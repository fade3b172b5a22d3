// microbench.cuh -- FP64 throughput probes: the roof that bounds K1/K2 is the DFMA pipe
// (SURVEY.md F4), and MEASURED_PEAKS.json has no FP64 entry, so bench.py measures it live.
#pragma once

#include "pmc_common.cuh"

namespace pmc {

constexpr int MB_THREADS = 384;
constexpr int MB_CHAINS = 8;

// which = 0: register-only DFMA chains.  1: one broadcast LDS.128 per 4 DFMA.  2: one per 2 DFMA.
template <int WHICH>
__global__ void __launch_bounds__(MB_THREADS, 1) mb_dfma(int iters, double seed, double* out) {
  __shared__ __align__(16) double tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double acc[MB_CHAINS];
#pragma unroll
  for (int c = 0; c < MB_CHAINS; ++c) acc[c] = seed + c + threadIdx.x * 1e-6;
  const double m0 = 1.0 - 1e-12, m1 = 1.0 + 1e-12;
  for (int it = 0; it < iters; ++it) {
    if (WHICH == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < MB_CHAINS; ++c) acc[c] = fma(acc[c], (u & 1) ? m0 : m1, 1e-30);
    } else {
      const double* t = tab + ((it * 64) & 1023);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (WHICH == 1) {  // 2 loads feed 8 DFMAs
          const double2 a = *reinterpret_cast<const double2*>(t + 4 * u);
          const double2 b = *reinterpret_cast<const double2*>(t + 4 * u + 2);
          acc[0] = fma(acc[0], a.x, 1e-30); acc[1] = fma(acc[1], a.x, 1e-30);
          acc[2] = fma(acc[2], a.y, 1e-30); acc[3] = fma(acc[3], a.y, 1e-30);
          acc[4] = fma(acc[4], b.x, 1e-30); acc[5] = fma(acc[5], b.x, 1e-30);
          acc[6] = fma(acc[6], b.y, 1e-30); acc[7] = fma(acc[7], b.y, 1e-30);
        } else {           // 4 loads feed 8 DFMAs
          const double2 a = *reinterpret_cast<const double2*>(t + 8 * u);
          const double2 b = *reinterpret_cast<const double2*>(t + 8 * u + 2);
          const double2 c = *reinterpret_cast<const double2*>(t + 8 * u + 4);
          const double2 d = *reinterpret_cast<const double2*>(t + 8 * u + 6);
          acc[0] = fma(acc[0], a.x, 1e-30); acc[1] = fma(acc[1], a.y, 1e-30);
          acc[2] = fma(acc[2], b.x, 1e-30); acc[3] = fma(acc[3], b.y, 1e-30);
          acc[4] = fma(acc[4], c.x, 1e-30); acc[5] = fma(acc[5], c.y, 1e-30);
          acc[6] = fma(acc[6], d.x, 1e-30); acc[7] = fma(acc[7], d.y, 1e-30);
        }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < MB_CHAINS; ++c) s += acc[c];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// FP64 tensor-core probe (mma.sync m8n8k4): not used by the product kernels (north_star: no tensor
// cores on this path); measured only to document where the DMMA roof sits relative to DFMA.
__global__ void __launch_bounds__(MB_THREADS, 1) mb_dmma(int iters, double seed, double* out) {
  double c[MB_CHAINS][2];
#pragma unroll
  for (int i = 0; i < MB_CHAINS; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
  const double a = 1.0 + 1e-12 * threadIdx.x, b = 1.0 - 1e-12 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < MB_CHAINS; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1])
                     : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < MB_CHAINS; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace pmc
