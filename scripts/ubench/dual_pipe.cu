// dual_pipe.cu -- are the FP64 FMA pipe (DFMA) and the FP64 tensor pipe (DMMA) separate units that run concurrently?
// 8 warps per SM: MODE 0 = all DFMA, 1 = all DMMA, 2 = warps 0-3 DFMA + warps 4-7 DMMA (one of each per sub-partition),
// each warp doing the SAME amount of its own work in every mode.  If the pipes were independent, mode 2 would take
// as long as the slower half alone; if they share the FMA lanes, as long as both halves back to back.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, double seed, double* out) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = (MODE == 1) || (MODE == 2 && warp >= 4);
  double s = 0;
  if (!do_mma) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = seed + c + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int u = 0; u < 8; ++u)        // 64 DFMA = 64 FMA per thread per iteration
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fma(acc[c], (u & 1) ? 1.0 - 1e-12 : 1.0 + 1e-12, 1e-30);
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
  } else {
    double c2[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c2[i][0] = seed + i; c2[i][1] = seed - i; }
    const double a = 1.0 + 1e-12 * threadIdx.x, b = 1.0 - 1e-12 * threadIdx.x;
    for (int it = 0; it < iters; ++it)   // 8 DMMA = 8 * 256 / 32 = 64 FMA per thread per iteration
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2[i][0]), "+d"(c2[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c2[i][0] + c2[i][1];
  }
  if (s == 12345.678) out[0] = s;
}
template <int MODE>
int run(int sms, double* out, const char* name) {
  const int iters = 100000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); k<MODE><<<sms, 256>>>(iters, 1.0, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  printf("%-40s %8.3f ms  %8.1f GFLOP/s\n", name, best, 2.0 * 64 * double(iters) * 256 * sms / best * 1e-6);
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); double* out; CK(cudaMalloc(&out, 64));
  run<0>(p.multiProcessorCount, out, "8 warps DFMA");
  run<1>(p.multiProcessorCount, out, "8 warps DMMA");
  run<2>(p.multiProcessorCount, out, "4 warps DFMA + 4 warps DMMA");
  return 0;
}
