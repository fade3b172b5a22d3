// pmc_common.cuh -- shared definitions for the sm_100a kernels of the mixture-density /
// proposal-update hot path (K1 = fused log-pdf + log-sum-exp + responsibilities,
// K2 = weighted sufficient statistics).  See DESIGN.md for the data layout.
#pragma once

#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>

namespace pmc {

// ---------------------------------------------------------------------------------------------
// Packed component record (all doubles), one per evaluated component, DP = D rounded up to even:
//
//   [0, NT)            T, the lower-triangular factor with T^T T = Sigma^-1 (or = W for VB),
//                      stored as 2x2 blocks: for row pair r = 0..DP/2-1, column pair p = 0..r
//                        { T[2r][2p], T[2r][2p+1], T[2r+1][2p], T[2r+1][2p+1] }
//                      (T[2r][2r+1] is the structural zero above the diagonal).
//                      Block (r,p) sits at 2 r (r+1) + 4 p.   NT = DP/2 (DP/2+1) 2.
//   [NT, NT+DP)        centre (mu_k, or m_k for VB); entry DP-1 is 0 when D is odd.
//   [NT+DP, NT+DP+8)   scalars, meaning depends on the mode (see Scalar enum).
//
// A record is a multiple of 16 bytes, so one cp.async.bulk moves it into shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kNumScalars = 8;
#define PMC_MAX_DP 64      // largest padded dimension the K1 instantiations cover
#define PMC_MAX_WARPS 16   // stride of the per-warp partial sums (>= warps per CTA of every K1 variant)

__host__ __device__ constexpr int tri_len(int DP) { return (DP / 2) * (DP / 2 + 1) * 2; }
__host__ __device__ constexpr int record_len(int DP) { return tri_len(DP) + DP + kNumScalars; }

enum Mode : int { MODE_GAUSS = 0, MODE_STUDENT_T = 1, MODE_VB = 2 };

// scalar slots
//   MODE_GAUSS     : S0 = log_normalization                              (gauss.pyx:54-56)
//   MODE_STUDENT_T : S0 = log_normalization, S1 = prefactor -(nu+D)/2, S2 = 1/nu, S3 = nu,
//                    S4 = nu + D                                         (student_t.pyx:32-34,116-117)
//   MODE_VB        : S0 = E[ln pi_k], S1 = E[ln det Lambda_k], S2 = D ln(2 pi), S3 = D / beta_k,
//                    S4 = nu_k                                           (variational.pyx:691,798)
//   all modes      : S5 = mixture weight w_k (1 for VB)
enum Scalar : int { S0 = 0, S1 = 1, S2 = 2, S3 = 3, S4 = 4, S_WEIGHT = 5 };

constexpr double kTiny = 2.2250738585072014e-308;  // numpy.finfo('d').tiny  (pmc.pyx:32)

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed for a
// contiguous record).  SASS: UBLKCP.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// global -> shared bulk copy; completion (bytes) is signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// error plumbing for the C ABI
// ---------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);

// cudaFuncSetAttribute is per device: remember per device whether a kernel's shared-memory limit was raised
struct PerDeviceFlag {
  bool done[64] = {};
  bool& here() {
    int dev = 0;
    cudaGetDevice(&dev);
    return done[(dev >= 0 && dev < 64) ? dev : 0];
  }
};

#define PMC_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::pmc::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

#define PMC_REQUIRE(cond, msg)                                                                     \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::pmc::set_last_error(std::string("pmcb200: ") + (msg));                                     \
      return 2;                                                                                    \
    }                                                                                              \
  } while (0)

}  // namespace pmc
