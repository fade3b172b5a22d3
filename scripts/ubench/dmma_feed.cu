// dmma_feed.cu -- what keeps K1's FP64 matrix-instruction loop (k1_mma_eval.cuh) below the DMMA peak?
// The loop of one warp: per feature quad, CB theta fragments (B operand) and NB phi fragments (A operand, each the
// product of two shared-memory loads) feed NB x CB mma.sync.m8n8k4.f64 into 8x8 accumulators.  This probe runs that
// loop WITHOUT epilogue or sample staging, adding one ingredient at a time:
//   bit 0 (1): operands come from shared memory every step (else: loaded once, register resident)
//   bit 1 (2): the A fragment is formed by a DMUL of two loaded values (else: one value used directly)
//   bit 2 (4): the two sample blocks of a lane sit next to each other (transposed layout, one LDS.128 serves both)
//   bit 3 (8): the (i, j) column offsets come from a table in shared memory (one extra LDS.32 per step)
// for NW = 8 and 16 warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_feed dmma_feed.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int CB, int NB, int MODE>
__global__ void __launch_bounds__(512, 1) k(int steps, int tiles, int D, double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  constexpr int KP = 8 * CB;
  constexpr bool LOADS = MODE & 1, MUL = MODE & 2, TRANS = MODE & 4, TAB = MODE & 8;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  double* theta_s = reinterpret_cast<double*>(raw);                       // [steps][KP][4]
  const int YS = TRANS ? (8 * NB + 4) : (((D + 2 - 4 + 15) / 16) * 16 + 4);  // transposed: [col][8 NB + 4]; else [row][YS]
  const int ylen = TRANS ? (D + 2) * YS : 8 * NB * YS;
  double* y_all = theta_s + size_t(steps) * KP * 4;
  int* tab = reinterpret_cast<int*>(y_all + size_t(nw) * ylen);
  for (int i = threadIdx.x; i < steps * KP * 4; i += blockDim.x) theta_s[i] = 1e-3 * (i % 97);
  for (int i = threadIdx.x; i < nw * ylen; i += blockDim.x) y_all[i] = 1.0 + 1e-6 * (i % 31);
  for (int f = threadIdx.x; f < steps * 4; f += blockDim.x) {
    int t = f, r = 0;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    r %= D;
    const int c = (t - r * (r + 1) / 2) % D;
    tab[f] = TRANS ? ((r * YS * 8) | ((c * YS * 8) << 16)) : (r | (c << 8));
  }
  __syncthreads();
  double* yw = y_all + size_t(warp) * ylen;
  double acc[NB][CB][2];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }
  const double* thl = theta_s + g * 4 + tq;
  double th_n[CB], yi_n[NB], yj_n[NB];
  auto fetch = [&](int s) {
    int oi, oj;
    if (TAB) { const int t = tab[4 * s + tq]; oi = TRANS ? (t & 0xffff) : (t & 0xff); oj = TRANS ? (t >> 16) : (t >> 8); }
    else { oi = TRANS ? ((s % D) * YS * 8) : (s % D); oj = TRANS ? (((s * 4 + tq) % D) * YS * 8) : ((s * 4 + tq) % D); }
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) th_n[cb] = thl[(s * KP + cb * 8) * 4];
    if (TRANS) {
      const char* base = reinterpret_cast<const char*>(yw) + g * NB * 8;
      if (NB == 2) {
        const double2 a = *reinterpret_cast<const double2*>(base + oi), b = *reinterpret_cast<const double2*>(base + oj);
        yi_n[0] = a.x; yi_n[NB - 1] = a.y; yj_n[0] = b.x; yj_n[NB - 1] = b.y;
      } else {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          yi_n[nb] = *reinterpret_cast<const double*>(base + oi + nb * 8);
          yj_n[nb] = *reinterpret_cast<const double*>(base + oj + nb * 8);
        }
      }
    } else {
      const double* yi = yw + g * YS + oi;
      const double* yj = yw + g * YS + oj;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) { yi_n[nb] = yi[nb * 8 * YS]; yj_n[nb] = yj[nb * 8 * YS]; }
    }
  };
  fetch(0);
  for (int tile = 0; tile < tiles; ++tile) {
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
      double th[CB], ph[NB];
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) th[cb] = th_n[cb];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) ph[nb] = MUL ? yi_n[nb] * yj_n[nb] : yi_n[nb];
      if (LOADS) fetch(min(s + 1, steps - 1));
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int cb = 0; cb < CB; ++cb)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[nb][cb][0]), "+d"(acc[nb][cb][1]) : "d"(ph[nb]), "d"(th[cb]));
    }
  }
  double s = 0;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) s += acc[nb][cb][0] + acc[nb][cb][1];
  if (s == 12345.678) out[0] = s;
}

template <int CB, int NB, int MODE>
int run(int sms, int nw, int D, double* out, double peak) {
  const int F = 1 + D + D * (D + 1) / 2, steps = (F + 3) / 4, tiles = 60;
  const int YS = (MODE & 4) ? (8 * NB + 4) : (((D + 2 - 4 + 15) / 16) * 16 + 4);
  const int ylen = (MODE & 4) ? (D + 2) * YS : 8 * NB * YS;
  const size_t smem = sizeof(double) * (size_t(steps) * 8 * CB * 4 + size_t(nw) * ylen) + sizeof(int) * steps * 4 + 16;
  if (smem > 227 * 1024) { printf("CB=%d NB=%d D=%d mode=%2d nw=%2d: does not fit (%zu B)\n", CB, NB, D, MODE, nw, smem); return 0; }
  CK(cudaFuncSetAttribute(k<CB, NB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); k<CB, NB, MODE><<<sms, nw * 32, smem>>>(steps, tiles, D, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  const double gf = 2.0 * 256.0 * double(NB * CB) * steps * tiles * nw * sms / best * 1e-6;
  printf("CB=%d NB=%d D=%d mode=%2d (%s%s%s%s) nw=%2d: %8.3f ms %9.1f GFLOP/s %5.1f %%\n", CB, NB, D, MODE,
         (MODE & 1) ? "lds " : "reg ", (MODE & 2) ? "dmul " : "", (MODE & 4) ? "transposed " : "", (MODE & 8) ? "table" : "",
         nw, best, gf, 100.0 * gf / peak);
  return 0;
}

#define ALL(CB, NB, D)                                                                      \
  for (int nw = 8; nw <= 16; nw += 8) {                                                      \
    run<CB, NB, 0>(sms, nw, D, out, peak); run<CB, NB, 2>(sms, nw, D, out, peak);           \
    run<CB, NB, 1>(sms, nw, D, out, peak); run<CB, NB, 3>(sms, nw, D, out, peak);           \
    run<CB, NB, 11>(sms, nw, D, out, peak); run<CB, NB, 7>(sms, nw, D, out, peak);          \
    run<CB, NB, 15>(sms, nw, D, out, peak);                                                  \
  }

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); double* out; CK(cudaMalloc(&out, 64));
  const int sms = p.multiProcessorCount;
  const double peak = 2.0 * 64 * sms * p.clockRate * 1e-6;   // GFLOP/s: 64 DFMA / clk / SM
  printf("%s, %d SMs, %.0f MHz, nominal FP64 peak %.0f GFLOP/s\n", p.name, sms, p.clockRate * 1e-3, peak);
  ALL(4, 2, 30)
  ALL(2, 2, 40)
  ALL(8, 2, 20)
  ALL(8, 1, 20)
  ALL(2, 4, 40)
  return 0;
}
