// k3_propose.cuh -- K3: draw samples from a Gaussian / Student-t mixture on the device, float64, sm_100a.
//
// Replaces (reference loops, /root/reference/pypmc):
//   MixtureDensity.propose     density/mixture.pyx:159-212   (component blocks in component order, `trace` origin array)
//   Gauss.propose              density/gauss.pyx:159-163     (one Python-level draw per sample: mu + L z)
//   StudentT.propose           density/student_t.pyx:49-55, 172-176   (mu + L z sqrt(dof / chi2_dof))
//
// The host draws the per-component counts with the caller's numpy generator exactly like the reference
// (rng.multinomial(N, weights), mixture.pyx:193), so the block structure and the `trace` array are the
// reference's; the normal / chi-square variates come from a counter-based Philox4x32-10 stream keyed by
// (seed, global sample index), so a sample does not depend on the grid, the tile or the rank that draws it.
// Bit parity with numpy's Mersenne Twister is impossible; parity is statistical.
//
// Shape: D normals (Box-Muller in double precision, ~250 FP64 instructions per pair) + D(D+1)/2 FMAs + 8 D bytes
// written per sample: FP64-pipe bound by the transcendental functions, ~2 ms per 1e7 x 30.
// Mapping: one thread per sample; the sample's normals sit in a per-thread shared-memory row (dynamic
// indexing without local memory), the Cholesky factor of its component is read from global memory through L1
// (all lanes of a warp share the component except at block boundaries, so the loads are broadcasts).
#pragma once

#include <curand_kernel.h>

#include "pmc_common.cuh"

namespace pmc {

constexpr int K3_THREADS = 128;

struct ProposeArgs {
  int64_t n;             // samples drawn by this launch
  int64_t ldx;
  int d, k;
  const double* means;   // [k, d]
  const double* chol;    // [k, d, d] lower-triangular L (Sigma = L L^T), row-major
  const double* dofs;    // [k] or null (Gaussian)
  const int64_t* starts; // [k + 1] first row of each component's block (starts[k] = n)
  unsigned long long seed;
  unsigned long long index0;   // global index of row 0 (rank offset), selects the Philox subsequence
  double* x;             // [n, ldx]
  int* latent;           // [n] or null: component that generated each row
};

// chi-square(nu) = 2 Gamma(nu/2): Marsaglia & Tsang (2000), with the alpha < 1 boost.
__device__ inline double k3_chisquare(curandStatePhilox4_32_10_t* st, double nu) {
  double alpha = 0.5 * nu, boost = 1.0;
  if (alpha < 1.0) {
    boost = pow(curand_uniform_double(st), 1.0 / alpha);
    alpha += 1.0;
  }
  const double dd = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
  for (int it = 0; it < 1000; ++it) {
    const double xn = curand_normal_double(st);
    double v = 1.0 + c * xn;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = curand_uniform_double(st);
    if (log(u) < 0.5 * xn * xn + dd - dd * v + dd * log(v)) return 2.0 * dd * v * boost;
  }
  return 2.0 * dd * boost;   // not reached in practice (acceptance > 95 % per trial)
}

// Shared memory: per-thread rows of normals [K3_THREADS][D | 1], the block starts [k + 1], and -- for tiles whose rows
// all belong to one component, i.e. all but K - 1 tiles of a launch -- that component's Cholesky factor and mean
// ([D][D] + [D]), staged once and kept while consecutive tiles stay in the component: the triangular product then reads
// L through broadcast LDS instead of one dependent global load per FMA, with two accumulators per output to halve the
// dependency chain (4.9 -> see profiles/ ms per 1e7 x 30 samples).
__host__ __device__ inline size_t k3_smem_bytes(int d, int k) {
  return sizeof(double) * (size_t(K3_THREADS) * (d | 1) + size_t(d) * d + d) + sizeof(int64_t) * size_t(k + 1) + 16;
}

__global__ void __launch_bounds__(K3_THREADS) k3_propose(const ProposeArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int D = a.d, ZS = D | 1;                                     // odd row stride: conflict-free columns
  double* zall = reinterpret_cast<double*>(smem_raw);
  double* zrow = zall + size_t(threadIdx.x) * ZS;
  double* Ls = zall + size_t(K3_THREADS) * ZS;                       // [D][D] factor of the staged component
  double* mus = Ls + size_t(D) * D;                                  // [D]
  int64_t* starts_s = reinterpret_cast<int64_t*>(mus + D);
  for (int i = threadIdx.x; i <= a.k; i += blockDim.x) starts_s[i] = a.starts[i];
  __syncthreads();
  int staged = -1;                                                   // component whose factor sits in Ls (block-uniform)

  for (int64_t base = int64_t(blockIdx.x) * K3_THREADS; base < a.n; base += int64_t(gridDim.x) * K3_THREADS) {
    const int64_t row = base + threadIdx.x;
    const int64_t last = (base + K3_THREADS <= a.n ? base + K3_THREADS : a.n) - 1;
    // component of a row: last c with starts[c] <= row (empty blocks have starts[c] == starts[c+1])
    auto component = [&](int64_t r) {
      int lo = 0, hi = a.k;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (starts_s[mid] <= r) lo = mid; else hi = mid;
      }
      return lo;
    };
    const int c_first = component(base), c_last = component(last);
    const bool uniform = c_first == c_last;                           // same for every thread of the block
    if (uniform && staged != c_first) {
      const double* L = a.chol + size_t(c_first) * D * D;
      for (int e = threadIdx.x; e < D * D; e += K3_THREADS) Ls[e] = __ldg(L + e);
      for (int e = threadIdx.x; e < D; e += K3_THREADS) mus[e] = __ldg(a.means + size_t(c_first) * D + e);
      staged = c_first;
    }
    int c = c_first;
    double scale = 1.0;
    if (row < a.n) {
      if (!uniform) c = component(row);
      curandStatePhilox4_32_10_t st;
      curand_init(a.seed, a.index0 + static_cast<unsigned long long>(row), 0ULL, &st);
      for (int j = 0; j < D; j += 2) {
        const double2 z = curand_normal2_double(&st);
        zrow[j] = z.x;
        if (j + 1 < D) zrow[j + 1] = z.y;
      }
      if (a.dofs) {
        const double nu = a.dofs[c];
        scale = sqrt(nu / k3_chisquare(&st, nu));                    // student_t.pyx:55
      }
      if (a.latent) a.latent[row] = c;
    }
    __syncthreads();                                                  // Ls / mus visible (and every zrow written)
    if (row < a.n) {
      if (uniform) {
        for (int i = D - 1; i >= 0; --i) {                           // descending: x_i may overwrite z_i
          const double* Li = Ls + i * D;
          double acc0 = 0.0, acc1 = 0.0;
          int j = 0;
          for (; j + 1 <= i; j += 2) {                                // gauss.pyx:50-52: dot(cholesky_sigma, z)
            acc0 = fma(Li[j], zrow[j], acc0);
            acc1 = fma(Li[j + 1], zrow[j + 1], acc1);
          }
          if (j <= i) acc0 = fma(Li[j], zrow[j], acc0);
          zrow[i] = mus[i] + (acc0 + acc1) * scale;
        }
      } else {
        const double* L = a.chol + size_t(c) * D * D;
        const double* mu = a.means + size_t(c) * D;
        for (int i = D - 1; i >= 0; --i) {
          double acc = 0.0;
          const double* Li = L + size_t(i) * D;
          for (int j = 0; j <= i; ++j) acc = fma(__ldg(Li + j), zrow[j], acc);
          zrow[i] = __ldg(mu + i) + acc * scale;
        }
      }
    }
    __syncthreads();
    // coalesced copy of the block's rows (contiguous in global memory when ldx == d)
    const int rows = int((a.n - base < K3_THREADS) ? (a.n - base) : K3_THREADS);
    if (a.ldx == D) {
      double* dst = a.x + base * D;
      if (ZS == D) {
        for (int e = threadIdx.x; e < rows * D; e += K3_THREADS) dst[e] = zall[e];
      } else {                                                        // rows of D values at stride D + 1
        int r = threadIdx.x / D, jj = threadIdx.x - r * D;            // element e = r D + jj, advanced by K3_THREADS per step
        const int dr = K3_THREADS / D, dj = K3_THREADS - dr * D;
        for (int e = threadIdx.x; e < rows * D; e += K3_THREADS) {
          dst[e] = zall[size_t(r) * ZS + jj];
          r += dr; jj += dj;
          if (jj >= D) { jj -= D; ++r; }
        }
      }
    } else {
      for (int e = threadIdx.x; e < rows * D; e += K3_THREADS) {
        const int r = e / D, jj = e - r * D;
        a.x[(base + r) * a.ldx + jj] = zall[size_t(r) * ZS + jj];
      }
    }
    __syncthreads();
  }
}

}  // namespace pmc
