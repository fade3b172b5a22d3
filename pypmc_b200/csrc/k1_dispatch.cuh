// k1_dispatch.cuh -- launch table of the K1 instantiations (one per even padded dimension DP).
#pragma once
#include "k1_mixture_eval.cuh"

namespace pmc {

// returns cudaError_t as int; grid <= #SMs (persistent CTAs, one per SM)
template <int DP>
int k1_launch_dp(const EvalArgs& a, int grid, cudaStream_t stream);

int k1_launch(int dp, const EvalArgs& a, int grid, cudaStream_t stream);
int k1_tile_rows(int dp);   // samples per CTA tile for this DP
int k1_warps(int dp);       // warps per CTA for this DP

#define PMC_K1_INSTANTIATE(DP)                                                                      \
  template <>                                                                                        \
  int k1_launch_dp<DP>(const EvalArgs& a, int grid, cudaStream_t stream) {                          \
    static bool attr_set = false;                                                                    \
    if (!attr_set) {                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(k1_mixture_eval<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           int(EvalCfg<DP>::SMEM_BYTES));                            \
      if (e != cudaSuccess) return int(e);                                                           \
      attr_set = true;                                                                               \
    }                                                                                                \
    k1_mixture_eval<DP><<<grid, EvalCfg<DP>::NW * 32, EvalCfg<DP>::SMEM_BYTES, stream>>>(a);         \
    return int(cudaGetLastError());                                                                  \
  }

}  // namespace pmc
