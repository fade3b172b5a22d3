// dfma_lat.cu -- how many independent DFMA chains does a B200 scheduler need?  (8 warps = 2 per sub-partition)
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int C, int TH>
__global__ void __launch_bounds__(TH, 1) k(int iters, double seed, double* out) {
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = seed + c + threadIdx.x * 1e-6;
  const double m0 = 1.0 - 1e-12, m1 = 1.0 + 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], (u & 1) ? m0 : m1, 1e-30);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
  if (s == 12345.678) out[0] = s;
}
template <int C, int TH>
int run(int sms, double* out) {
  const int iters = 40000 / C;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); k<C, TH><<<sms, TH>>>(iters, 1.0, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  const double fl = 2.0 * 16 * C * double(iters) * TH * sms;
  // cycles per dependent DFMA of one chain = time * clk / (iters*16)
  printf("warps=%2d chains/warp=%d  %8.3f ms  %8.1f GFLOP/s   %.2f clk between dependent DFMAs of a chain\n", TH / 32, C, best,
         fl / best * 1e-6, best * 1e-3 * 1.965e9 / (double(iters) * 16));
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); double* out; CK(cudaMalloc(&out, 64));
  int s = p.multiProcessorCount;
  run<1, 128>(s, out); run<2, 128>(s, out); run<3, 128>(s, out); run<4, 128>(s, out); run<6, 128>(s, out); run<8, 128>(s, out);
  run<1, 256>(s, out); run<2, 256>(s, out); run<3, 256>(s, out); run<4, 256>(s, out); run<6, 256>(s, out); run<8, 256>(s, out);
  run<2, 384>(s, out); run<3, 384>(s, out); run<4, 384>(s, out);
  return 0;
}
