// tdeliver.cu -- micro-benchmark: how fast can the triangular factor T reach the FP64 pipe?
// Variants of K1's inner loop (q = ||T x - b||^2, x register-resident, D = 2H) with T delivered from
//   SRC 0: shared memory, broadcast LDS.128 (the K1 of round 1)
//   SRC 1: constant bank, LDCU.64 -> uniform register operand of DFMA
//   SRC 2: constant bank, 128-bit loads
//   SRC 3: split: row pairs r < RS from shared memory, the rest from the constant bank
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tdeliver tdeliver.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int CONST_DOUBLES = 8000;
__constant__ double cT[CONST_DOUBLES];

__host__ __device__ constexpr int tri_len(int DP) { return (DP / 2) * (DP / 2 + 1) * 2; }

template <int DP, int S, int SRC, int RS, int TH>
__global__ void __launch_bounds__(TH, 1) kern(const double* __restrict__ gT, const double* __restrict__ x, double* out,
                                               int kl, int iters) {
  constexpr int H = DP / 2, NT = tri_len(DP);
  extern __shared__ __align__(16) double sT[];
  if (SRC == 0 || SRC == 3) {
    for (int i = threadIdx.x; i < kl * NT; i += blockDim.x) sT[i] = gT[i];
    __syncthreads();
  }
  double y[S][DP];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < DP; ++j) y[s][j] = x[(size_t(blockIdx.x * blockDim.x + threadIdx.x) * S + s) * DP + j];
  double acc[S];
#pragma unroll
  for (int s = 0; s < S; ++s) acc[s] = 0;
  for (int it = 0; it < iters; ++it) {
    for (int k = 0; k < kl; ++k) {
      const double* ts = sT + k * NT;
      const double* tc = cT + k * NT;
      double q[S];
#pragma unroll
      for (int s = 0; s < S; ++s) q[s] = 0;
#pragma unroll
      for (int r = 0; r < H; ++r) {
        double z0[S], z1[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { z0[s] = 0; z1[s] = 0; }
#pragma unroll
        for (int p = 0; p <= r; ++p) {
          double a, b, c, d;
          const int off = 2 * r * (r + 1) + 4 * p;
          if (SRC == 0 || (SRC == 3 && r < RS)) {
            const double2 t0 = *reinterpret_cast<const double2*>(ts + off);
            const double2 t1 = *reinterpret_cast<const double2*>(ts + off + 2);
            a = t0.x; b = t0.y; c = t1.x; d = t1.y;
          } else if (SRC == 2) {
            const double2 t0 = *reinterpret_cast<const double2*>(tc + off);
            const double2 t1 = *reinterpret_cast<const double2*>(tc + off + 2);
            a = t0.x; b = t0.y; c = t1.x; d = t1.y;
          } else {
            a = tc[off]; b = tc[off + 1]; c = tc[off + 2]; d = tc[off + 3];
          }
#pragma unroll
          for (int s = 0; s < S; ++s) {
            z0[s] = fma(a, y[s][2 * p], z0[s]);
            if (p < r) z0[s] = fma(b, y[s][2 * p + 1], z0[s]);
            z1[s] = fma(c, y[s][2 * p], z1[s]);
            z1[s] = fma(d, y[s][2 * p + 1], z1[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) { q[s] = fma(z0[s], z0[s], q[s]); q[s] = fma(z1[s], z1[s], q[s]); }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] += q[s];
    }
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s) r += acc[s];
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = r;
}

template <int DP, int S, int SRC, int RS, int TH>
void run(const char* name, int kl, const double* gT, const double* x, double* out, int sms) {
  constexpr int NT = tri_len(DP);
  if (kl * NT > CONST_DOUBLES) return;
  const size_t smem = (SRC == 0 || SRC == 3) ? size_t(kl) * NT * 8 : 0;
  CK(cudaFuncSetAttribute(kern<DP, S, SRC, RS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 6400 / kl / S * 2;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    kern<DP, S, SRC, RS, TH><<<sms, TH, smem>>>(gT, x, out, kl, iters);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) best = ms < best ? ms : best;
  }
  const double H = DP / 2;
  const double fma_per = (DP * (DP + 1) / 2.0) + DP;   // triangular (structural zeros skipped) + squares
  (void)H;
  const double flops = 2.0 * fma_per * S * double(TH) * sms * kl * iters;
  printf("%-34s DP=%d S=%d threads=%d kl=%2d  %8.3f ms  %8.1f GFLOP/s\n", name, DP, S, TH, kl, best, flops / best * 1e-6);
  fflush(stdout);
}

int main() {
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, sms);
  std::vector<double> hT(CONST_DOUBLES);
  for (int i = 0; i < CONST_DOUBLES; ++i) hT[i] = 1e-3 * ((i * 2654435761u) % 1000) / 1000.0;
  CK(cudaMemcpyToSymbol(cT, hT.data(), CONST_DOUBLES * 8));
  double *gT, *x, *out;
  CK(cudaMalloc(&gT, CONST_DOUBLES * 8)); CK(cudaMemcpy(gT, hT.data(), CONST_DOUBLES * 8, cudaMemcpyHostToDevice));
  const size_t nx = size_t(sms) * 512 * 8 * 40;
  CK(cudaMalloc(&x, nx * 8)); CK(cudaMemset(x, 0, nx * 8));
  CK(cudaMalloc(&out, size_t(sms) * 512 * 8));
  for (int kl : {8}) {
    run<30, 2, 0, 0, 384>("smem S=2 12w", kl, gT, x, out, sms);
    run<30, 2, 0, 0, 352>("smem S=2 11w", kl, gT, x, out, sms);
    run<30, 2, 0, 0, 320>("smem S=2 10w", kl, gT, x, out, sms);
    run<30, 2, 0, 0, 288>("smem S=2 9w", kl, gT, x, out, sms);
    run<30, 2, 0, 0, 256>("smem S=2 8w", kl, gT, x, out, sms);
    run<30, 3, 0, 0, 256>("smem S=3 8w", kl, gT, x, out, sms);
    run<30, 3, 0, 0, 224>("smem S=3 7w", kl, gT, x, out, sms);
    run<30, 3, 0, 0, 192>("smem S=3 6w", kl, gT, x, out, sms);
    run<20, 4, 0, 0, 256>("smem S=4 8w", kl, gT, x, out, sms);
    run<20, 4, 0, 0, 320>("smem S=4 10w", kl, gT, x, out, sms);
    run<20, 4, 0, 0, 384>("smem S=4 12w", kl, gT, x, out, sms);
    run<20, 3, 0, 0, 384>("smem S=3 12w", kl, gT, x, out, sms);
    run<20, 3, 0, 0, 448>("smem S=3 14w", kl, gT, x, out, sms);
    run<20, 5, 0, 0, 256>("smem S=5 8w", kl, gT, x, out, sms);
    run<20, 6, 0, 0, 256>("smem S=6 8w", kl, gT, x, out, sms);
    run<40, 2, 0, 0, 256>("smem S=2 8w", kl, gT, x, out, sms);
    run<40, 2, 0, 0, 320>("smem S=2 10w", kl, gT, x, out, sms);
    run<40, 2, 0, 0, 384>("smem S=2 12w", kl, gT, x, out, sms);
    run<40, 3, 0, 0, 256>("smem S=3 8w", kl, gT, x, out, sms);
    run<10, 4, 0, 0, 384>("smem S=4 12w", kl, gT, x, out, sms);
    run<10, 8, 0, 0, 384>("smem S=8 12w", kl, gT, x, out, sms);
    run<10, 6, 0, 0, 512>("smem S=6 16w", kl, gT, x, out, sms);
  }
  return 0;
}
