// dfma_rf.cu -- is the register file (3 x 64-bit reads per DFMA) a roof below the DFMA issue rate?
// 8 warps per SM, 12 independent accumulators per thread, operands:
//   V0: fma(acc, imm, imm)                 one register read per DFMA (the "peak" probe)
//   V1: fma(a[i], b[j], acc[c])            three register reads, operands rotate (no reuse)
//   V2: fma(a[i], b[s], acc[c]) with the same a[i] for 3 consecutive DFMAs (K1's pattern: T shared by S samples)
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int V>
__global__ void __launch_bounds__(256, 1) k(int iters, const double* __restrict__ in, double* out) {
  double a[12], b[12], acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { a[i] = in[i + threadIdx.x]; b[i] = in[64 + i + threadIdx.x]; acc[i] = 0.0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 12; ++u) {
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        if (V == 0) acc[c] = fma(acc[c], 1.0000000001, 1e-30);
        if (V == 1) acc[c] = fma(a[(c + u) % 12], b[(c * 5 + u) % 12], acc[c]);
        if (V == 2) acc[c] = fma(a[(c / 3 + u) % 12], b[(c % 3 + 3 * (u % 4)) % 12], acc[c]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 12; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
int run(int sms, const double* in, double* out, const char* name) {
  const int iters = 20000;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); k<V><<<sms, 256>>>(iters, in, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  printf("%-44s %8.3f ms  %8.1f GFLOP/s\n", name, best, 2.0 * 144 * double(iters) * 256 * sms / best * 1e-6);
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double *in, *out; CK(cudaMalloc(&in, 4096 * 8)); CK(cudaMemset(in, 0, 4096 * 8)); CK(cudaMalloc(&out, p.multiProcessorCount * 256 * 8));
  run<0>(p.multiProcessorCount, in, out, "V0 fma(acc, imm, imm)");
  run<1>(p.multiProcessorCount, in, out, "V1 fma(a[i], b[j], acc) rotating operands");
  run<2>(p.multiProcessorCount, in, out, "V2 fma(a[i], b[s], acc) a shared by 3 DFMAs");
  return 0;
}
