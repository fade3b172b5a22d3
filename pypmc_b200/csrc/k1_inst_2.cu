// K1 instantiations, group 2 (split over translation units so they compile in parallel)
#include "k1_dispatch.cuh"
namespace pmc {
PMC_K1_INSTANTIATE(34)
PMC_K1_INSTANTIATE(36)
PMC_K1_INSTANTIATE(38)
PMC_K1_INSTANTIATE(40)
PMC_K1_INSTANTIATE(42)
PMC_K1_INSTANTIATE(44)
PMC_K1_INSTANTIATE(46)
PMC_K1_INSTANTIATE(48)
}  // namespace pmc
