// K1 matrix-instruction form: instantiations, component-group plan and launch (k1_mma_eval.cuh)
#include "k1_dispatch.cuh"

#include <algorithm>
#include <cstdlib>

namespace pmc {

constexpr size_t kSmemLimit = 227 * 1024;

// Tile variant per component-block count CB: 16 warps (four per scheduler) measured best -- a warp issues at most
// one DMMA per ~32 clk, the pipe takes one per 16, and a warp in its epilogue issues none (C2: 8 warps 11.3 ms, 12
// warps 10.7 ms, 16 warps 10.5 ms; C3: (8,4,8) 12.3 ms, (8,2,12) 11.5 ms, (8,2,16) 11.8 ms with spills, (8,1,16)
// 11.1 ms; profiles/r01f_k1_mma_variants.md).  NB = 2 sample blocks per warp while the accumulators leave room: always
// for the eval-only instantiation (theta fragments then feed two DMMAs each: K=56, D=20 4.63 vs 4.99 ms), up to CB = 5
// with the fused second pass, whose second register array would spill beyond that (C3 VB 17.8 vs 13.7 ms; round 2:
// (8,2,12) with 168 registers still spills 420 B and measures 13.7 vs 12.5 ms).  The
// per-sample arithmetic does not depend on NB, so both instantiations give the same log q bit for bit.
static int nb_for(int cb, bool second) {
  static const char* env = getenv("PMCB200_K1_NB1");              // tuning runs: one sample block for the fused second pass from CB = N on
  const int from = env ? atoi(env) : 6;
  return (second && cb >= 5 && cb >= from) ? 1 : 2;
}
constexpr int kWarps = 16;
int k1_mma_nb(int cb, bool second) { return nb_for(cb, second); }

bool k1_mma_plan(int kl, int d, K1Launch::MmaPlan* plan) {
  if (kl < 9 || d < 8) return false;                  // few components / tiny D: the DFMA form's epilogue-bound regime
  int cbmax = 0;                                      // largest block count whose theta fits beside the sample slices
  for (int c = 8; c >= 2; --c)
    if (k1m_smem_bytes(d, 8 * c, 2, kWarps) <= kSmemLimit) { cbmax = c; break; }
  if (cbmax == 0) return false;
  const int full = kl / (8 * cbmax), rest = kl - full * 8 * cbmax;
  const int cb_rest = rest ? std::max(2, (rest + 7) / 8) : 0;
  const int groups = full + (rest ? 1 : 0), slots = full * 8 * cbmax + 8 * cb_rest;
  // padding is issued work: beyond 20 % the DFMA form (70 % of the pipe) wins; CB = 2 groups are bound by the
  // shared-memory pipe (74 %) and only pay where the DFMA form is weak (D > 32)
  if (groups > K1Launch::MmaPlan::kMaxGroups || slots * 100 > kl * 120) return false;
  if (groups > 1 && cbmax < 4 && d < 33) return false;
  plan->groups = groups;
  plan->steps = k1m_steps(d);
  size_t off = 0;
  for (int g = 0; g < groups; ++g) {
    plan->k0[g] = g * 8 * cbmax;
    plan->cb[g] = (g < full) ? cbmax : cb_rest;
    plan->count[g] = (g < full) ? 8 * cbmax : rest;
    plan->theta_off[g] = off;
    off += size_t(plan->steps) * 8 * plan->cb[g] * 4;
  }
  plan->theta_len = off;
  return true;
}

template <int CB, int NB, int NW, bool SECOND, int DIAG = 0>
static int launch(const MmaArgs& ma, int sm_count, size_t smem, cudaStream_t stream) {
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k1_mma_eval<CB, NB, NW, SECOND, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemLimit));
    if (e != cudaSuccess) return int(e);
    attr_set = true;
  }
  const int ts = 8 * NB * NW;
  const int64_t tiles = (ma.e.n + ts - 1) / ts;
  const int grid = int(std::min<int64_t>(tiles, sm_count));
  k1_mma_eval<CB, NB, NW, SECOND, DIAG><<<grid, NW * 32, smem, stream>>>(ma);
  return int(cudaGetLastError());
}

int k1_mma_launch(const K1Launch& l, int sm_count, cudaStream_t stream) {
  const K1Launch::MmaPlan& p = l.mma;
  const bool second_pass = (l.base.resp_out != nullptr) || (l.base.mode == MODE_VB && l.base.lp_out != nullptr);
  const bool second = second_pass && p.groups == 1;                  // fused; with groups k1_finish does it
  const int rl = record_len((l.base.d + 1) & ~1);
  for (int g = 0; g < p.groups; ++g) {
    const int cb = p.cb[g], nb = nb_for(cb, second);
    MmaArgs ma{l.base, l.theta + p.theta_off[g], l.shift, l.flag, p.steps, 8 * cb, g, p.groups, l.rowstat};
    ma.e.records = l.derived + size_t(p.k0[g]) * rl;
    ma.e.cols = l.base.cols + p.k0[g];
    ma.e.kl = p.count[g];
    if (g + 1 < p.groups) ma.e.partials = nullptr;                   // the sums belong to the last launch
    const size_t smem = k1m_smem_bytes(l.base.d, 8 * cb, nb, kWarps);
    int rc = int(cudaErrorInvalidValue);
    // measurement builds: PMCB200_K1_DIAG=1..7 swaps in a kernel with parts removed (WRONG RESULTS; timing only)
    const char* diag_env = getenv("PMCB200_K1_DIAG");
    const int diag = diag_env ? atoi(diag_env) : 0;
    if (diag && cb == 4 && !second) {
      switch (diag) {
        case 1: rc = launch<4, 2, kWarps, false, 1>(ma, sm_count, smem, stream); break;
        case 2: rc = launch<4, 2, kWarps, false, 2>(ma, sm_count, smem, stream); break;
        case 3: rc = launch<4, 2, kWarps, false, 3>(ma, sm_count, smem, stream); break;
        case 4: rc = launch<4, 2, kWarps, false, 4>(ma, sm_count, smem, stream); break;
        case 6: rc = launch<4, 2, kWarps, false, 6>(ma, sm_count, smem, stream); break;
        default: break;
      }
      if (rc != 0) return rc;
      continue;
    }
#define PMC_K1M_CASE(CBV, NB_ALT)                                                              \
    if (cb == CBV)                                                                              \
      rc = !second ? launch<CBV, 2, kWarps, false>(ma, sm_count, smem, stream)                  \
           : (nb == 2) ? launch<CBV, 2, kWarps, true>(ma, sm_count, smem, stream)               \
                       : launch<CBV, NB_ALT, kWarps, true>(ma, sm_count, smem, stream);
    PMC_K1M_CASE(2, 2) PMC_K1M_CASE(3, 2) PMC_K1M_CASE(4, 2) PMC_K1M_CASE(5, 1) PMC_K1M_CASE(6, 1)
    PMC_K1M_CASE(7, 1) PMC_K1M_CASE(8, 1)
#undef PMC_K1M_CASE
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace pmc
