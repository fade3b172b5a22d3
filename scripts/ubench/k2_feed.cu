// k2_feed.cu -- what keeps K2's consumer loop (k2_suffstats.cuh, DMMA form) below the DMMA peak?
// One warp per step of 4 samples: CB A-fragments (v, LDS.64), FB B-fragments (y_i * y_j: 2 LDS.64 + 1 DMUL),
// CB x FB mma.sync.m8n8k4.f64.  This probe runs that loop on a static stage (no producer, no mbarriers) and varies
//   VEC = 1: row-major stage [sample][column], one LDS.64 per operand and 4-sample step (what K2 does today)
//   VEC = 2: transposed stage [8-sample block][column][8 slots], samples s and s + 4 adjacent: one LDS.128 per operand
//            serves two 4-sample steps (half the LDS instructions per DMMA)
//   NW  = consumer warps per SM (8 = two per scheduler, 12, 16)
// and, first, pure DMMA chains with 4 / 8 warps per SM (the per-warp issue interval).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k2_feed k2_feed.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

template <int NACC>
__global__ void __launch_bounds__(512, 1) pure(int iters, double* out) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) DMMA(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 12345.678) out[0] = s;
}

// TN samples per stage, D+1 columns of [y, 1], KP components
template <int CB, int FB, int VEC, int NW>
__global__ void __launch_bounds__(NW * 32, 1) feed(int tiles, int TN, int D, int KP, double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int DP = (D + 2 + 3) & ~3;
  // VEC 1: V [TN][VS], Y [TN][YS]; VEC 2: V [TN/8][KP][8], Y [TN/8][DP][8]
  const int VS = ((KP + 15) / 16) * 16 + 4, YS = ((DP + 15) / 16) * 16 + 4;
  double* Vs = reinterpret_cast<double*>(raw);
  double* Ys = Vs + (VEC == 1 ? TN * VS : TN * KP);
  const int vlen = (VEC == 1 ? TN * VS : TN * KP), ylen = (VEC == 1 ? TN * YS : TN * DP);
  for (int i = threadIdx.x; i < vlen; i += blockDim.x) Vs[i] = 1e-3 * (i % 97);
  for (int i = threadIdx.x; i < ylen; i += blockDim.x) Ys[i] = 1.0 + 1e-6 * (i % 31);
  __syncthreads();
  const int F = 1 + D + D * (D + 1) / 2;
  int off_i[FB], off_j[FB];
#pragma unroll
  for (int fb = 0; fb < FB; ++fb) {
    const int f = ((warp * FB + fb) * 8 + g) % F;
    int oi = D, oj = D;
    if (f == 0) { oi = D; oj = D; }
    else if (f <= D) { oi = f - 1; oj = D; }
    else { int t = f - 1 - D, r = 0; while ((r + 1) * (r + 2) / 2 <= t) ++r; oi = r; oj = t - r * (r + 1) / 2; }
    off_i[fb] = oi; off_j[fb] = oj;
  }
  double acc[CB][FB][2];
#pragma unroll
  for (int cb = 0; cb < CB; ++cb)
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) { acc[cb][fb][0] = 0.0; acc[cb][fb][1] = 0.0; }

  for (int tile = 0; tile < tiles; ++tile) {
    if constexpr (VEC == 1) {
#pragma unroll 2
      for (int n0 = 0; n0 < TN; n0 += 4) {
        const double* vrow = Vs + (n0 + tq) * VS + g;
        const double* yrow = Ys + (n0 + tq) * YS;
        double av[CB], bv[FB];
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) av[cb] = vrow[cb * 8];
#pragma unroll
        for (int fb = 0; fb < FB; ++fb) bv[fb] = yrow[off_i[fb]] * yrow[off_j[fb]];
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[cb][fb][0], acc[cb][fb][1], av[cb], bv[fb]);
      }
    } else if constexpr (VEC == 3) {
      // VEC 2 + software pipeline: the v fragments and the first feature block's y operands of the NEXT 8-sample block are
      // fetched during the DMMAs of this one, so a pass starts with its first DMMAs instead of an LDS round trip
      const char* vb = reinterpret_cast<const char*>(Vs) + (g * 8 + 2 * tq) * 8;
      const char* yb = reinterpret_cast<const char*>(Ys) + (2 * tq) * 8;
      constexpr int PA = (CB < 4) ? CB : 4;                  // v fragments fetched ahead
      double2 av_n[PA], yi_n, yj_n;
#pragma unroll
      for (int cb = 0; cb < PA; ++cb) av_n[cb] = *reinterpret_cast<const double2*>(vb + cb * 8 * 64);
      yi_n = *reinterpret_cast<const double2*>(yb + off_i[0] * 64);
      yj_n = *reinterpret_cast<const double2*>(yb + off_j[0] * 64);
      for (int n0 = 0; n0 < TN; n0 += 8) {
        const char* vblk = vb + (n0 / 8) * KP * 64;
        const char* yblk = yb + (n0 / 8) * DP * 64;
        const int nn = (n0 + 8 < TN) ? n0 + 8 : 0;
        const char* vnext = vb + (nn / 8) * KP * 64;
        const char* ynext = yb + (nn / 8) * DP * 64;
        double2 av[CB];
        double b0[FB], b1[FB];
#pragma unroll
        for (int cb = 0; cb < PA; ++cb) av[cb] = av_n[cb];
        b0[0] = yi_n.x * yj_n.x; b1[0] = yi_n.y * yj_n.y;
#pragma unroll
        for (int cb = PA; cb < CB; ++cb) av[cb] = *reinterpret_cast<const double2*>(vblk + cb * 8 * 64);
#pragma unroll
        for (int fb = 1; fb < FB; ++fb) {
          const double2 yi = *reinterpret_cast<const double2*>(yblk + off_i[fb] * 64);
          const double2 yj = *reinterpret_cast<const double2*>(yblk + off_j[fb] * 64);
          b0[fb] = yi.x * yj.x; b1[fb] = yi.y * yj.y;
        }
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[cb][fb][0], acc[cb][fb][1], av[cb].x, b0[fb]);
#pragma unroll
        for (int cb = 0; cb < PA; ++cb) av_n[cb] = *reinterpret_cast<const double2*>(vnext + cb * 8 * 64);
        yi_n = *reinterpret_cast<const double2*>(ynext + off_i[0] * 64);
        yj_n = *reinterpret_cast<const double2*>(ynext + off_j[0] * 64);
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[cb][fb][0], acc[cb][fb][1], av[cb].y, b1[fb]);
      }
    } else {
      // slot order inside an 8-sample block: (0,4,1,5,2,6,3,7): lane tq reads slots 2 tq, 2 tq + 1 = samples tq, tq + 4
      const char* vb = reinterpret_cast<const char*>(Vs) + (g * 8 + 2 * tq) * 8;
      const char* yb = reinterpret_cast<const char*>(Ys) + (2 * tq) * 8;
      for (int n0 = 0; n0 < TN; n0 += 8) {
        const char* vblk = vb + (n0 / 8) * KP * 64;
        const char* yblk = yb + (n0 / 8) * DP * 64;
        double2 av[CB], yi[FB], yj[FB];
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) av[cb] = *reinterpret_cast<const double2*>(vblk + cb * 8 * 64);
#pragma unroll
        for (int fb = 0; fb < FB; ++fb) {
          yi[fb] = *reinterpret_cast<const double2*>(yblk + off_i[fb] * 64);
          yj[fb] = *reinterpret_cast<const double2*>(yblk + off_j[fb] * 64);
        }
        double b0[FB], b1[FB];
#pragma unroll
        for (int fb = 0; fb < FB; ++fb) { b0[fb] = yi[fb].x * yj[fb].x; b1[fb] = yi[fb].y * yj[fb].y; }
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[cb][fb][0], acc[cb][fb][1], av[cb].x, b0[fb]);
#pragma unroll
        for (int fb = 0; fb < FB; ++fb)
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) DMMA(acc[cb][fb][0], acc[cb][fb][1], av[cb].y, b1[fb]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int cb = 0; cb < CB; ++cb)
#pragma unroll
    for (int fb = 0; fb < FB; ++fb) s += acc[cb][fb][0] + acc[cb][fb][1];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
int run_pure(int sms, int nw, double* out, double peak) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000; float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); pure<NACC><<<sms, nw * 32>>>(iters, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  const double gf = 2.0 * 256.0 * NACC * double(iters) * nw * sms / best * 1e-6;
  printf("pure DMMA, %2d independent accumulators, nw=%2d: %8.3f ms %9.1f GFLOP/s %5.1f %%\n", NACC, nw, best, gf, 100.0 * gf / peak);
  return 0;
}

template <int CB, int FB, int VEC, int NW>
int run(int sms, int D, double* out, double peak) {
  const int nw = NW, KP = 8 * CB, TN = 96, tiles = 40, DP = (D + 2 + 3) & ~3;
  const int VS = ((KP + 15) / 16) * 16 + 4, YS = ((DP + 15) / 16) * 16 + 4;
  const size_t smem = sizeof(double) * (VEC == 1 ? size_t(TN) * (VS + YS) : size_t(TN) * (KP + DP));
  CK(cudaFuncSetAttribute(feed<CB, FB, VEC, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); feed<CB, FB, VEC, NW><<<sms, nw * 32, smem>>>(tiles, TN, D, KP, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  const double gf = 2.0 * 256.0 * double(CB * FB) * (TN / 4) * tiles * nw * sms / best * 1e-6;
  printf("CB=%d FB=%2d D=%d VEC=%d nw=%2d: %8.3f ms %9.1f GFLOP/s %5.1f %%\n", CB, FB, D, VEC, nw, best, gf, 100.0 * gf / peak);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); double* out; CK(cudaMalloc(&out, 64));
  const int sms = p.multiProcessorCount;
  const double peak = 2.0 * 64 * sms * p.clockRate * 1e-6;
  printf("%s, %d SMs, %.0f MHz, nominal FP64 peak %.0f GFLOP/s\n", p.name, sms, p.clockRate * 1e-3, peak);
  run_pure<8>(sms, 4, out, peak);
  // 8 consumer warps: K2's shapes at C2, C4, C3 and the in-between block counts
  run<4, 8, 1, 8>(sms, 30, out, peak); run<4, 8, 2, 8>(sms, 30, out, peak); run<4, 8, 3, 8>(sms, 30, out, peak);
  run<2, 14, 1, 8>(sms, 40, out, peak); run<2, 14, 2, 8>(sms, 40, out, peak); run<2, 14, 3, 8>(sms, 40, out, peak);
  run<8, 4, 1, 8>(sms, 20, out, peak); run<8, 4, 2, 8>(sms, 20, out, peak); run<8, 4, 3, 8>(sms, 20, out, peak);
  run<5, 4, 1, 8>(sms, 20, out, peak); run<5, 4, 2, 8>(sms, 20, out, peak); run<5, 4, 3, 8>(sms, 20, out, peak);
  run<6, 4, 1, 8>(sms, 30, out, peak); run<6, 4, 2, 8>(sms, 30, out, peak); run<6, 4, 3, 8>(sms, 30, out, peak);
  run<3, 8, 1, 8>(sms, 30, out, peak); run<3, 8, 2, 8>(sms, 30, out, peak); run<3, 8, 3, 8>(sms, 30, out, peak);
  run<4, 6, 3, 12>(sms, 30, out, peak); run<2, 9, 3, 12>(sms, 40, out, peak); run<8, 3, 3, 12>(sms, 20, out, peak);
  return 0;
}
