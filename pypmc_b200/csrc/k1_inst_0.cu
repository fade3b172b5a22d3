// K1 instantiations, group 0 (split over translation units so they compile in parallel)
#include "k1_dispatch.cuh"
namespace pmc {
PMC_K1_INSTANTIATE(2)
PMC_K1_INSTANTIATE(4)
PMC_K1_INSTANTIATE(6)
PMC_K1_INSTANTIATE(8)
PMC_K1_INSTANTIATE(10)
PMC_K1_INSTANTIATE(12)
PMC_K1_INSTANTIATE(14)
PMC_K1_INSTANTIATE(16)
}  // namespace pmc
