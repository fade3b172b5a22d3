// K1 instantiations, group 1 (split over translation units so they compile in parallel)
#include "k1_dispatch.cuh"
namespace pmc {
PMC_K1_INSTANTIATE(18)
PMC_K1_INSTANTIATE(20)
PMC_K1_INSTANTIATE(22)
PMC_K1_INSTANTIATE(24)
PMC_K1_INSTANTIATE(26)
PMC_K1_INSTANTIATE(28)
PMC_K1_INSTANTIATE(30)
PMC_K1_INSTANTIATE(32)
}  // namespace pmc
