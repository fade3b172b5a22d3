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
// Shape: D normals + D(D+1)/2 FMAs + 8 D bytes written per sample.  The library's Box-Muller (curand_normal2_double:
// log, sqrt, sincospi in double precision, ~250 FP64 instructions per pair) made the kernel FP64-pipe bound by the
// transcendental functions, 2 of 3.3 ms per 1e7 x 30; k3_normal2 below does the same transformation in ~50 (the table
// logarithm of the other kernels, a Taylor sine / cosine on [-pi/4, pi/4], uniforms built from the random bits by an
// exponent trick instead of integer-to-double conversions).
// Mapping: one thread per sample; the sample's normals sit in a per-thread shared-memory row (dynamic
// indexing without local memory), the Cholesky factor of its component is read from global memory through L1
// (all lanes of a warp share the component except at block boundaries, so the loads are broadcasts).
#pragma once

#include <curand_kernel.h>

#include "pmc_common.cuh"
#include "k1_exp_table.cuh"

namespace pmc {

constexpr int K3_THREADS = 128;

// Two independent standard normals from the next four 32-bit words of the sample's Philox stream (Box-Muller):
//   u1 = 2 - [1, 2) in (0, 1] (52 random bits), r = sqrt(-2 ln u1);  angle = 2 pi u2, u2 = [1, 2) - 1 in [0, 1)
// ln u1: table logarithm (x = 2^e m, c_j = 1 + (j + 1/2)/128, r = m / c_j - 1, degree-6 polynomial; k2_log_pos without
// its fallback -- u1 is a positive normal number by construction).  Sine and cosine: the angle in quadrant units a = 4 u2,
// q = round(a), x = (a - q) pi/2 in [-pi/4, pi/4], Taylor polynomials to x^15 / x^16 (truncation < 5e-17), quadrant by
// sign flips and a swap.  Absolute error of a normal ~ 2e-16 (1 + |z|); the largest |z| is sqrt(2 * 52 ln 2) = 8.49.
__device__ __forceinline__ double2 k3_normal2(curandStatePhilox4_32_10_t* st, const double* __restrict__ ltab /* [1/c_j | ln c_j] */) {
  const uint4 w = curand4(st);
  const double d1 = __hiloint2double(int(0x3ff00000u | (w.x >> 12)), int(w.y));     // [1, 2)
  const double d2 = __hiloint2double(int(0x3ff00000u | (w.z >> 12)), int(w.w));
  const double u1 = 2.0 - d1;                                                      // (0, 1]
  // -- ln u1 --
  const int hi = __double2hiint(u1);
  const int j = (hi >> 13) & 127;
  const double e = double((hi >> 20) - 1023);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u1));
  const double rr = fma(m, ltab[j], -1.0);
  double p = fma(rr, -1.66666666666666657e-01, 2.00000000000000011e-01);
  p = fma(rr, p, -0.25);
  p = fma(rr, p, 3.33333333333333315e-01);
  p = fma(rr, p, -0.5);
  p = fma(rr * rr, p, rr);
  const double lnu = fma(e, 0x1.62e42fefa38p-1, (p + ltab[128 + j]) + e * 0x1.ef35793c7673p-45);
  const double rad = sqrt(-2.0 * lnu);
  // -- sin / cos of 2 pi u2 --
  const double a = fma(d2, 4.0, -4.0);                                             // [0, 4), exact
  const double magic = 6755399441055744.0;                                         // 1.5 * 2^52
  const double t = a + magic;
  const int q = __double2loint(t);                                                 // round(a) in 0..4
  const double x = (a - (t - magic)) * 0x1.921fb54442d18p+0;                       // pi/2
  const double x2 = x * x;
  double sp = fma(x2, -0x1.ae7f3e733b81fp-41, 0x1.6124613a86d09p-33);
  sp = fma(x2, sp, -0x1.ae64567f544e4p-26);
  sp = fma(x2, sp, 0x1.71de3a556c734p-19);
  sp = fma(x2, sp, -0x1.a01a01a01a01ap-13);
  sp = fma(x2, sp, 0x1.1111111111111p-7);
  sp = fma(x2, sp, -0x1.5555555555555p-3);
  const double sn = fma(x * x2, sp, x);
  double cp = fma(x2, 0x1.ae7f3e733b81fp-45, -0x1.93974a8c07c9dp-37);
  cp = fma(x2, cp, 0x1.1eed8eff8d898p-29);
  cp = fma(x2, cp, -0x1.27e4fb7789f5cp-22);
  cp = fma(x2, cp, 0x1.a01a01a01a01ap-16);
  cp = fma(x2, cp, -0x1.6c16c16c16c17p-10);
  cp = fma(x2, cp, 0x1.5555555555555p-5);
  cp = fma(x2, cp, -0.5);
  const double cs = fma(x2, cp, 1.0);
  // quadrant q (mod 4): (cos, sin) = (c, s), (-s, c), (-c, -s), (s, -c)
  double c0 = (q & 1) ? sn : cs, s0 = (q & 1) ? cs : sn;
  if ((q + 1) & 2) c0 = -c0;
  if (q & 2) s0 = -s0;
  return make_double2(rad * c0, rad * s0);
}

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
  return sizeof(double) * (size_t(K3_THREADS) * (d | 1) + size_t(d) * ((d + 1) & ~1) + ((d + 1) & ~1) + 256) +
         sizeof(int64_t) * size_t(k + 1) + 16;
}

// DMAX > 0 (D <= DMAX <= 40): the sample's normals and the triangular product stay in REGISTERS -- the shared-memory row per
// thread made the product cost two LDS per FMA (1.5 of 3.1 ms at 1e7 x 30: LSU-bound); here a broadcast LDS.128 of the
// factor feeds two FMAs and nothing else is loaded.  The Philox words are consumed in the same order, so both forms
// draw the same samples.  DMAX = 0: any D, normals in the per-thread shared-memory row.
template <int DMAX>
__global__ void __launch_bounds__(K3_THREADS) k3_propose(const ProposeArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int D = a.d, ZS = D | 1, DL = (D + 1) & ~1;                  // odd row stride: conflict-free columns; even stride of L
  double* zall = reinterpret_cast<double*>(smem_raw);
  double* zrow = zall + size_t(threadIdx.x) * ZS;
  double* Ls = zall + size_t(K3_THREADS) * ZS;                       // [D][DL] factor of the staged component (16-byte aligned rows)
  double* mus = Ls + size_t(D) * DL;                                 // [DL]
  double* ltab = mus + DL;                                           // [128 | 128]: 1 / c_j, ln c_j (k3_normal2)
  int64_t* starts_s = reinterpret_cast<int64_t*>(ltab + 256);
  for (int i = threadIdx.x; i <= a.k; i += blockDim.x) starts_s[i] = a.starts[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { ltab[i] = kLogInvC[i]; ltab[128 + i] = kLogC[i]; }
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
      for (int e = threadIdx.x; e < D * D; e += K3_THREADS) {
        const int i = e / D, j = e - i * D;
        Ls[i * DL + j] = __ldg(L + e);
      }
      for (int e = threadIdx.x; e < D; e += K3_THREADS) mus[e] = __ldg(a.means + size_t(c_first) * D + e);
      staged = c_first;
    }
    int c = c_first;
    double scale = 1.0;
    double zr[DMAX > 0 ? DMAX : 1];
    if (row < a.n) {
      if (!uniform) c = component(row);
      curandStatePhilox4_32_10_t st;
      curand_init(a.seed, a.index0 + static_cast<unsigned long long>(row), 0ULL, &st);
      if constexpr (DMAX > 0) {
#pragma unroll
        for (int j = 0; j < DMAX; j += 2) {
          if (j < D) {
            const double2 z = k3_normal2(&st, ltab);
            zr[j] = z.x;
            if (j + 1 < DMAX) zr[j + 1] = z.y;
          }
        }
        if (!uniform) {                                               // the rare tile across a component boundary: generic product
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            if (j < D) zrow[j] = zr[j];
        }
      } else {
        for (int j = 0; j < D; j += 2) {
          const double2 z = k3_normal2(&st, ltab);
          zrow[j] = z.x;
          if (j + 1 < D) zrow[j + 1] = z.y;
        }
      }
      if (a.dofs) {
        const double nu = a.dofs[c];
        scale = sqrt(nu / k3_chisquare(&st, nu));                    // student_t.pyx:55
      }
      if (a.latent) a.latent[row] = c;
    }
    __syncthreads();                                                  // Ls / mus visible (and every zrow written)
    if (row < a.n) {
      if (uniform && DMAX > 0) {
        if constexpr (DMAX > 0) {
          // x_i = mu_i + scale sum_{j <= i} L_ij z_j, i descending so that x_i may replace z_i; all indices static
#pragma unroll
          for (int i = DMAX - 1; i >= 0; --i) {
            if (i < D) {
              const double* Li = Ls + i * DL;
              double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
              for (int j = 0; j + 1 <= i; j += 2) {
                const double2 l2 = *reinterpret_cast<const double2*>(Li + j);
                acc0 = fma(l2.x, zr[j], acc0);
                acc1 = fma(l2.y, zr[j + 1], acc1);
              }
              if ((i & 1) == 0) acc0 = fma(Li[i], zr[i], acc0);
              zr[i] = mus[i] + (acc0 + acc1) * scale;
            }
          }
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            if (j < D) zrow[j] = zr[j];
        }
      } else if (uniform) {
        for (int i = D - 1; i >= 0; --i) {                           // descending: x_i may overwrite z_i
          const double* Li = Ls + i * DL;
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
