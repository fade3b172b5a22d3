// K1 instantiations, group 3 (split over translation units so they compile in parallel)
#include "k1_dispatch.cuh"
namespace pmc {
PMC_K1_INSTANTIATE(50)
PMC_K1_INSTANTIATE(52)
PMC_K1_INSTANTIATE(54)
PMC_K1_INSTANTIATE(56)
PMC_K1_INSTANTIATE(58)
PMC_K1_INSTANTIATE(60)
PMC_K1_INSTANTIATE(62)
PMC_K1_INSTANTIATE(64)
}  // namespace pmc
