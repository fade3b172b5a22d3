// dfma_snake.cu -- outer-product accumulation acc[i][j] += a[i] * b[j] on the FP64 pipe: does the ORDER of the
// DFMAs matter?  The register file delivers about one 64-bit warp operand per clock per sub-partition
// (dfma_rf.cu), a DFMA wants three; the operand-reuse cache can serve an operand the previous DFMA read in the
// same slot.  Row-major order re-uses a[i] only; snake order alternates the shared operand (a, then b, then a ...)
// so every DFMA needs just ONE new input operand plus its accumulator.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void dfma(double& acc, double a, double b) {
  asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b));
}

// I x J accumulators; ORDER 0 = row-major (i outer), 1 = snake, 2 = snake via asm volatile
template <int I, int J, int ORDER>
__global__ void __launch_bounds__(256, 1) k(int iters, const double* __restrict__ in, double* out) {
  double a[I], b[J], acc[I][J];
#pragma unroll
  for (int i = 0; i < I; ++i) a[i] = in[i + threadIdx.x];
#pragma unroll
  for (int j = 0; j < J; ++j) b[j] = in[64 + j + threadIdx.x];
#pragma unroll
  for (int i = 0; i < I; ++i)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < I; ++i) {
#pragma unroll
        for (int jj = 0; jj < J; ++jj) {
          const int j = (ORDER != 0 && (i & 1)) ? (J - 1 - jj) : jj;
          if (ORDER == 2) dfma(acc[i][j], a[i], b[j]);
          else acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
      }
      // perturb the operands a little so that nothing is loop-invariant across u
#pragma unroll
      for (int i = 0; i < I; ++i) a[i] += 1e-9;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < I; ++i)
#pragma unroll
    for (int j = 0; j < J; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int I, int J, int ORDER>
int run(int sms, const double* in, double* out, const char* name) {
  const int iters = 200000 / (I * J);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); k<I, J, ORDER><<<sms, 256>>>(iters, in, out); CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) best = ms < best ? ms : best;
  }
  printf("%-40s %dx%d  %8.3f ms  %8.1f GFLOP/s (DFMA only; +%d DADD per %d DFMA not counted)\n", name, I, J, best,
         2.0 * 4 * I * J * double(iters) * 256 * sms / best * 1e-6, I, I * J);
  return 0;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); const int s = p.multiProcessorCount;
  double *in, *out; CK(cudaMalloc(&in, 4096 * 8)); CK(cudaMemset(in, 0, 4096 * 8)); CK(cudaMalloc(&out, s * 256 * 8));
  run<2, 3, 0>(s, in, out, "row-major"); run<2, 3, 1>(s, in, out, "snake"); run<2, 3, 2>(s, in, out, "snake, asm volatile");
  run<4, 3, 0>(s, in, out, "row-major"); run<4, 3, 1>(s, in, out, "snake"); run<4, 3, 2>(s, in, out, "snake, asm volatile");
  run<16, 4, 0>(s, in, out, "row-major"); run<16, 4, 1>(s, in, out, "snake"); run<16, 4, 2>(s, in, out, "snake, asm volatile");
  run<4, 16, 0>(s, in, out, "row-major"); run<4, 16, 1>(s, in, out, "snake"); run<4, 16, 2>(s, in, out, "snake, asm volatile");
  run<8, 8, 0>(s, in, out, "row-major"); run<8, 8, 1>(s, in, out, "snake"); run<8, 8, 2>(s, in, out, "snake, asm volatile");
  return 0;
}
