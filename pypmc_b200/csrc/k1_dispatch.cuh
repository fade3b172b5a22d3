// k1_dispatch.cuh -- launch table of the K1 instantiations (one per even padded dimension DP).
// One logical K1 launch = the fast form (k1_fast_eval.cuh) followed by the exact-difference form
// (k1_mixture_eval.cuh) on the same stream; the device-side flag written by k1_prepare decides which of the
// two does the work (the other returns immediately).
#pragma once
#include "k1_mma_eval.cuh"

namespace pmc {

struct K1Launch {
  EvalArgs base;            // records = the caller's packed records (T | centre | scalars)
  const double* derived;    // records with the centre slot replaced by -b (k1_prepare)
  const double* shift;      // [DP]
  const int* flag;
  double* rowstat;          // [n, 2] for k1_finish, or null
  // matrix-instruction form (k1_mma_eval.cuh); mma.groups == 0: not offered for this launch
  struct MmaPlan {
    static constexpr int kMaxGroups = 16;
    int groups = 0;                       // component groups, one k1_mma_eval launch each
    int k0[kMaxGroups] = {}, count[kMaxGroups] = {}, cb[kMaxGroups] = {};   // first component, components, blocks of 8
    size_t theta_off[kMaxGroups] = {};    // offset (doubles) of the group's theta [steps / 2][8 cb][4][2]
    size_t theta_len = 0;
    int steps = 0;
  } mma;
  const double* theta = nullptr;
};

// Plans the component groups of the matrix-instruction form for kl components of dimension d; false if the form
// does not apply (too few components, tiny D, or a grouping that would pad the component count by more than 20 %).
bool k1_mma_plan(int kl, int d, K1Launch::MmaPlan* plan);
int k1_mma_launch(const K1Launch& l, int sm_count, cudaStream_t stream);
int k1_mma_nb(int cb, bool second);   // sample blocks per warp of the instantiation k1_mma_launch picks

// returns cudaError_t as int; grid <= #SMs (persistent CTAs, one per SM)
template <int DP>
int k1_launch_dp(const K1Launch& l, int grid, cudaStream_t stream);

int k1_launch(int dp, const K1Launch& l, int grid, cudaStream_t stream);
int k1_tile_rows(int dp);   // samples per CTA tile for this DP (same for both forms)

#define PMC_K1_INSTANTIATE(DP)                                                                       \
  template <>                                                                                         \
  int k1_launch_dp<DP>(const K1Launch& l, int grid, cudaStream_t stream) {                           \
    static_assert(FastCfg<DP>::TS == EvalCfg<DP>::TS, "both K1 forms must tile the samples alike");  \
    static_assert(FastCfg<DP>::NW <= PMC_MAX_WARPS && EvalCfg<DP>::NW <= PMC_MAX_WARPS, "partials"); \
    static_assert(FastCfg<DP>::SMEM_STAGED + 16 <= 227 * 1024, "staging does not fit");             \
    static PerDeviceFlag attr_flag;                                                                   \
    bool& attr_set = attr_flag.here();                                                                \
    if (!attr_set) {                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(k1_mixture_eval<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           int(EvalCfg<DP>::SMEM_BYTES));                             \
      if (e != cudaSuccess) return int(e);                                                            \
      e = cudaFuncSetAttribute(k1_fast_eval<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                               int(FastCfg<DP>::SMEM_STAGED + 16));                                   \
      if (e != cudaSuccess) return int(e);                                                            \
      attr_set = true;                                                                                \
    }                                                                                                 \
    FastArgs fa{l.base, l.shift, l.flag, l.rowstat};                                                          \
    fa.e.records = l.derived;                                                                         \
    const bool staged = l.base.lp_out || l.base.resp_out || l.base.aux_out;                           \
    const size_t smem = (staged ? FastCfg<DP>::SMEM_STAGED : FastCfg<DP>::SMEM_BASE) + 16;            \
    k1_fast_eval<DP><<<grid, FastCfg<DP>::NW * 32, smem, stream>>>(fa);                               \
    cudaError_t e = cudaGetLastError();                                                               \
    if (e != cudaSuccess) return int(e);                                                              \
    EvalArgs ea = l.base;                                                                             \
    ea.flag = l.flag;                                                                                 \
    k1_mixture_eval<DP><<<grid, EvalCfg<DP>::NW * 32, EvalCfg<DP>::SMEM_BYTES, stream>>>(ea);         \
    return int(cudaGetLastError());                                                                   \
  }

}  // namespace pmc
