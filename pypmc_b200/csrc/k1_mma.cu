// K1 matrix-instruction form: instantiations and launch (k1_mma_eval.cuh)
#include "k1_dispatch.cuh"

#include <algorithm>

namespace pmc {

constexpr size_t kSmemLimit = 227 * 1024;

bool k1_mma_config(int kl, int d, int* cb, int* nb) {
  if (kl < 9 || kl > 128 || d < 8) return false;      // few components / tiny D: the DFMA form's epilogue-bound regime
  const int c = (kl <= 16) ? 2 : (kl <= 32) ? 4 : (kl <= 64) ? 8 : 16;
  const int n = (c == 16) ? 2 : 4;
  if (k1m_smem_bytes(d, 8 * c, n) > kSmemLimit) return false;
  *cb = c;
  *nb = n;
  return true;
}

template <int CB, int NB>
static int launch(const MmaArgs& ma, int sm_count, size_t smem, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k1_mma_eval<CB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit));
    if (e != cudaSuccess) return int(e);
    attr_set = true;
  }
  const int ts = 8 * NB * K1M_NW;
  const int64_t tiles = (ma.e.n + ts - 1) / ts;
  const int grid = int(std::min<int64_t>(tiles, sm_count));
  k1_mma_eval<CB, NB><<<grid, K1M_NW * 32, smem, stream>>>(ma);
  return int(cudaGetLastError());
}

int k1_mma_launch(const K1Launch& l, int sm_count, cudaStream_t stream) {
  MmaArgs ma{l.base, l.theta, l.shift, l.flag, l.rowstat, l.mma_steps, l.mma_kp, l.mma_ys};
  ma.e.records = l.derived;
  const size_t smem = k1m_smem_bytes(l.base.d, l.mma_kp, l.mma_nb);
  if (l.mma_cb == 2 && l.mma_nb == 4) return launch<2, 4>(ma, sm_count, smem, stream);
  if (l.mma_cb == 4 && l.mma_nb == 4) return launch<4, 4>(ma, sm_count, smem, stream);
  if (l.mma_cb == 8 && l.mma_nb == 4) return launch<8, 4>(ma, sm_count, smem, stream);
  if (l.mma_cb == 16 && l.mma_nb == 2) return launch<16, 2>(ma, sm_count, smem, stream);
  return int(cudaErrorInvalidValue);
}

}  // namespace pmc
