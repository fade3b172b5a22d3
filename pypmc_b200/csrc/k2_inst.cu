// K2 instantiations for component-block counts between the powers of two (CB = 3, 5, 6, 7): with them a mixture of
// K components runs ceil(K / 8) blocks instead of the next power of two (K = 40: 5 blocks instead of 8).
#define PMC_K2_TEMPLATE_ONLY
#include "k2_suffstats.cuh"

namespace pmc {

template <int CB, int FB>
static int launch(const StatsArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(k2_suffstats<CB, FB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return 1;
    attr_set = true;
  }
  k2_suffstats<CB, FB><<<grid, K2_THREADS, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int k2_launch_extra(int cb, int fb, const StatsArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
#define PMC_K2X(CBV, FBV) \
  if (cb == CBV && fb == FBV) return launch<CBV, FBV>(a, grid, smem, stream);
  PMC_K2X(3, 2) PMC_K2X(3, 4) PMC_K2X(3, 6) PMC_K2X(3, 8) PMC_K2X(3, 10)
  PMC_K2X(5, 2) PMC_K2X(5, 3) PMC_K2X(5, 4) PMC_K2X(5, 5) PMC_K2X(5, 6)
  PMC_K2X(6, 2) PMC_K2X(6, 3) PMC_K2X(6, 4) PMC_K2X(6, 5)
  PMC_K2X(7, 2) PMC_K2X(7, 3) PMC_K2X(7, 4)
#undef PMC_K2X
  return 2;
}

}  // namespace pmc
