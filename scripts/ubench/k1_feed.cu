// k1_feed.cu -- K1's DMMA loop (k1_mma_eval.cuh) fed one feature QUAD at a time (today: table LDS.32, 2 y loads, CB theta
// LDS.64 per NB x CB DMMAs) against one feature PAIR OF QUADS at a time (table LDS.64, 4 y loads, CB theta LDS.128 per
// 2 NB x CB DMMAs): an LDS instruction of either width costs the sub-partition's DMMA issue about the same
// (scripts/ubench/k2_feed.cu), so half the theta loads should buy back part of the idle pipe.
//   PAIR = 0: quad at a time, operands of quad s+1 fetched before the DMMAs of quad s (th / th_n double buffer)
//   PAIR = 1: pair at a time, y / table of the next pair prefetched, theta double-buffered (th / th_n as double2)
//   PAIR = 2: pair at a time, theta fragments reloaded IN PLACE right after their last DMMA of the pair (no second buffer)
//   PAIR = 3: PAIR 2 with the row operand y_i reloaded only every second pair (what keeping it in registers along a row of
//             the triangle would save: no measurable gain at C2 / C3, see the log)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k1_feed k1_feed.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

template <int CB, int NB, int NW, int PAIR>
__global__ void __launch_bounds__(NW * 32, 1) k(int steps, int tiles, int D, double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  constexpr int KP = 8 * CB, RS = 8 * NB + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  double* theta_s = reinterpret_cast<double*>(raw);                       // PAIR 0: [steps][KP][4]; else [steps/2][KP][4][2]
  const int slice = (D + 2) * RS;
  double* y_all = theta_s + size_t(steps) * KP * 4;
  int* tab = reinterpret_cast<int*>(y_all + size_t(NW) * slice);         // PAIR 0: [steps][4]; else [steps/2][4][2]
  for (int i = threadIdx.x; i < steps * KP * 4; i += blockDim.x) theta_s[i] = 1e-3 * (i % 97);
  for (int i = threadIdx.x; i < NW * slice; i += blockDim.x) y_all[i] = 1.0 + 1e-6 * (i % 31);
  for (int f = threadIdx.x; f < steps * 4; f += blockDim.x) {
    int t = f, r = 0;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    r %= D;
    const int c = (t - r * (r + 1) / 2) % D;
    const int word = (r * RS * 8) | ((c * RS * 8) << 16);
    if (PAIR == 0) tab[f] = word;
    else { const int s = f >> 2, q4 = f & 3; tab[((s >> 1) * 4 + q4) * 2 + (s & 1)] = word; }
  }
  __syncthreads();
  const double* yw = y_all + size_t(warp) * slice;
  const char* ylane = reinterpret_cast<const char*>(yw + g * NB);
  double acc[NB][CB][2];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }

  auto ld_y = [&](unsigned off, double (&v)[NB]) {
    if constexpr (NB == 2) { const double2 t = *reinterpret_cast<const double2*>(ylane + off); v[0] = t.x; v[1] = t.y; }
    else v[0] = *reinterpret_cast<const double*>(ylane + off);
  };

  for (int tile = 0; tile < tiles; ++tile) {
    if constexpr (PAIR == 0) {
      const double* thl = theta_s + g * 4 + tq;
      const int* tabl = tab + tq;
      double th_n[CB], yi_n[NB], yj_n[NB];
      auto fetch = [&](int s) {
        const unsigned t = unsigned(tabl[4 * s]);
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) th_n[cb] = thl[(s * KP + cb * 8) * 4];
        ld_y(t & 0xffffu, yi_n); ld_y(t >> 16, yj_n);
      };
      fetch(0);
      constexpr int UNR = (CB * NB <= 4) ? 4 : (CB * NB >= 12) ? 1 : 2;
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
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[nb][cb][0], acc[nb][cb][1], ph[nb], th[cb]);
      }
    } else {
      const int pairs = steps / 2;
      const double2* thl = reinterpret_cast<const double2*>(theta_s) + g * 4 + tq;     // theta[p][8 cb + g][tq] = (quad 2p, quad 2p+1)
      const int2* tabl = reinterpret_cast<const int2*>(tab) + tq;
      double yi_n[2][NB], yj_n[2][NB];
      double2 th[CB], th_n[PAIR == 1 ? CB : 1];
      auto fetch_y = [&](int p) {
        const int2 t = tabl[4 * p];
        if (PAIR != 3 || (p & 1) == 0) {                      // PAIR 3: the row operand y_i stays in registers for two pairs
          ld_y(unsigned(t.x) & 0xffffu, yi_n[0]);
          ld_y(unsigned(t.y) & 0xffffu, yi_n[1]);
        }
        ld_y(unsigned(t.x) >> 16, yj_n[0]);
        ld_y(unsigned(t.y) >> 16, yj_n[1]);
      };
      fetch_y(0);
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) {
        if constexpr (PAIR == 1) th_n[cb] = thl[(0 * KP + cb * 8) * 4]; else th[cb] = thl[(0 * KP + cb * 8) * 4];
      }
      constexpr int UNR = (CB * NB <= 4) ? 2 : 1;
#pragma unroll UNR
      for (int p = 0; p < pairs; ++p) {
        double ph[2][NB];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) ph[q][nb] = yi_n[q][nb] * yj_n[q][nb];
        const int pn = min(p + 1, pairs - 1);
        fetch_y(pn);
        if constexpr (PAIR == 1) {
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) th[cb] = th_n[cb];
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) th_n[cb] = thl[(pn * KP + cb * 8) * 4];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
              for (int cb = 0; cb < CB; ++cb) DMMA(acc[nb][cb][0], acc[nb][cb][1], ph[q][nb], q ? th[cb].y : th[cb].x);
        } else {
          // component blocks in groups of G: quad 0 then quad 1 of the group (same accumulator G NB DMMAs apart), then the
          // group's theta fragments are reloaded for the next pair
          constexpr int G = (CB % 4 == 0) ? 4 : (CB % 2 == 0) ? 2 : 1;
#pragma unroll
          for (int c0 = 0; c0 < CB; c0 += G) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int cb = c0; cb < c0 + G; ++cb) DMMA(acc[nb][cb][0], acc[nb][cb][1], ph[q][nb], q ? th[cb].y : th[cb].x);
#pragma unroll
            for (int cb = c0; cb < c0 + G; ++cb) th[cb] = thl[(pn * KP + cb * 8) * 4];
          }
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) s += acc[nb][cb][0] + acc[nb][cb][1];
  if (s == 12345.678) out[0] = s;
}

template <int CB, int NB, int NW, int PAIR>
int run(int sms, int D, double* out, double peak) {
  const int F = 1 + D + D * (D + 1) / 2, steps = ((F + 3) / 4 + 1) & ~1, tiles = 60, RS = 8 * NB + 4;
  const size_t smem = sizeof(double) * (size_t(steps) * 8 * CB * 4 + size_t(NW) * (D + 2) * RS) + sizeof(int) * steps * 4 + 16;
  if (smem > 227 * 1024) { printf("CB=%d NB=%d D=%d nw=%d: does not fit (%zu B)\n", CB, NB, D, NW, smem); return 0; }
  CK(cudaFuncSetAttribute(k<CB, NB, NW, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); k<CB, NB, NW, PAIR><<<sms, NW * 32, smem>>>(steps, tiles, D, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  const double gf = 2.0 * 256.0 * double(NB * CB) * steps * tiles * NW * sms / best * 1e-6;
  printf("CB=%d NB=%d D=%d nw=%2d PAIR=%d: %8.3f ms %9.1f GFLOP/s %5.1f %%\n", CB, NB, D, NW, PAIR, best, gf, 100.0 * gf / peak);
  return 0;
}

#define ALL(CB, NB, D) run<CB, NB, 16, 2>(sms, D, out, peak); run<CB, NB, 16, 3>(sms, D, out, peak);

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); double* out; CK(cudaMalloc(&out, 64));
  const int sms = p.multiProcessorCount;
  const double peak = 2.0 * 64 * sms * p.clockRate * 1e-6;
  printf("%s, %d SMs, %.0f MHz, nominal FP64 peak %.0f GFLOP/s\n", p.name, sms, p.clockRate * 1e-3, peak);
  ALL(4, 2, 30) ALL(2, 2, 40) ALL(8, 2, 20) ALL(8, 1, 20) ALL(3, 2, 30) ALL(6, 2, 20)
  return 0;
}
