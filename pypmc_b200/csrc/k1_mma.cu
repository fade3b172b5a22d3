// K1 matrix-instruction form: instantiations and launch (k1_mma_eval.cuh)
#include "k1_dispatch.cuh"

#include <algorithm>
#include <cstdlib>

namespace pmc {

constexpr size_t kSmemLimit = 227 * 1024;

// (CB, NB, NW) variants compiled in.  16 warps (four per scheduler) measured best: a warp issues at most one DMMA
// per ~32 clk, the pipe takes one per 16, and a warp in its epilogue issues none (C2: 8 warps 11.3 ms, 12 warps
// 10.7 ms, 16 warps 10.5 ms; C3: (8,4,8) 12.3 ms, (8,2,12) 11.5 ms, (8,2,16) 11.8 ms with spills, (8,1,16) 11.1 ms).
static bool have_variant(int cb, int nb, int nw) {
  return (cb == 2 && nb == 2 && nw == 16) || (cb == 4 && nb == 2 && nw == 16) || (cb == 8 && nb == 1 && nw == 16) ||
         (cb == 4 && nb == 4 && nw == 8) || (cb == 8 && nb == 2 && nw == 12);
}

bool k1_mma_config(int kl, int d, int* cb, int* nb, int* nw, int* groups) {
  if (kl < 9 || d < 8) return false;                  // few components / tiny D: the DFMA form's epilogue-bound regime
  int en = 0, ew = 0;
  if (const char* env = getenv("PMCB200_K1_MMA_CFG"))        // "NB,NW": tuning runs
    if (sscanf(env, "%d,%d", &en, &ew) != 2) en = ew = 0;
  // Largest component block count whose theta (all feature quads x 8 CB components) fits shared memory beside the
  // sample slices; more components than that are evaluated in groups (one launch each, log-sum-exp carried in
  // rowstat).  Padding of the last group is issued work: accept a grouping only while the padded count stays within
  // 15 % of kl -- beyond that the DFMA form (70 %) is the better choice.
  const int first = (kl <= 16) ? 2 : (kl <= 32) ? 4 : 8;
  for (int c = first; c >= 2; c /= 2) {
    int n = (c == 8) ? 1 : 2, w = 16;
    if (have_variant(c, en, ew)) { n = en; w = ew; }
    if (k1m_smem_bytes(d, 8 * c, n, w) > kSmemLimit) continue;
    const int g = (kl + 8 * c - 1) / (8 * c);
    if (g > 1 && (g * 8 * c * 100 > kl * 115 || (c < 4 && d < 33))) continue;   // (CB = 2 is LDS-bound: only where DFMA is weak)
    *cb = c;
    *nb = n;
    *nw = w;
    *groups = g;
    return true;
  }
  return false;
}

template <int CB, int NB, int NW, bool SECOND>
static int launch(const MmaArgs& ma, int sm_count, size_t smem, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k1_mma_eval<CB, NB, NW, SECOND>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit));
    if (e != cudaSuccess) return int(e);
    attr_set = true;
  }
  const int ts = 8 * NB * NW;
  const int64_t tiles = (ma.e.n + ts - 1) / ts;
  const int grid = int(std::min<int64_t>(tiles, sm_count));
  k1_mma_eval<CB, NB, NW, SECOND><<<grid, NW * 32, smem, stream>>>(ma);
  return int(cudaGetLastError());
}

int k1_mma_launch(const K1Launch& l, int sm_count, cudaStream_t stream) {
  const size_t smem = k1m_smem_bytes(l.base.d, l.mma_kp, l.mma_nb, l.mma_nw);
  const bool second_pass = (l.base.resp_out != nullptr) || (l.base.mode == MODE_VB && l.base.lp_out != nullptr);
  const bool second = second_pass && l.mma_groups == 1;              // fused; with groups k1_finish does it
  const int rl = record_len((l.base.d + 1) & ~1);
  for (int g = 0; g < l.mma_groups; ++g) {
    MmaArgs ma{l.base, l.theta + size_t(g) * l.mma_steps * l.mma_kp * 4, l.shift, l.flag, l.mma_steps, l.mma_kp, l.mma_ys,
               g, l.mma_groups, l.rowstat};
    ma.e.records = l.derived + size_t(g) * l.mma_kp * rl;
    ma.e.cols = l.base.cols + g * l.mma_kp;
    ma.e.kl = std::min(l.mma_kp, l.base.kl - g * l.mma_kp);
    if (g + 1 < l.mma_groups) ma.e.partials = nullptr;               // the sums belong to the last launch
    int rc = int(cudaErrorInvalidValue);
#define PMC_K1M_CASE(CBV, NBV, NWV)                                                             \
    if (l.mma_cb == CBV && l.mma_nb == NBV && l.mma_nw == NWV)                                  \
      rc = second ? launch<CBV, NBV, NWV, true>(ma, sm_count, smem, stream) : launch<CBV, NBV, NWV, false>(ma, sm_count, smem, stream);
    PMC_K1M_CASE(2, 2, 16) PMC_K1M_CASE(4, 2, 16) PMC_K1M_CASE(8, 1, 16) PMC_K1M_CASE(4, 4, 8) PMC_K1M_CASE(8, 2, 12)
#undef PMC_K1M_CASE
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace pmc
