// pmcb200.cu -- C ABI (include/pmcb200.h) over the sm_100a kernels K1 (k1_mixture_eval.cuh) and
// K2 (k2_suffstats.cuh).  No torch types, no CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/pmcb200.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "k1_dispatch.cuh"
#include "k1_prepare.cuh"
#include "k2_suffstats.cuh"
#include "k3_propose.cuh"
#include "k4_weights.cuh"
#include "microbench.cuh"

namespace pmc {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

int k1_launch(int dp, const K1Launch& a, int grid, cudaStream_t stream) {
  switch (dp) {
#define PMC_CASE(DP) \
  case DP:           \
    return k1_launch_dp<DP>(a, grid, stream);
    PMC_CASE(2) PMC_CASE(4) PMC_CASE(6) PMC_CASE(8) PMC_CASE(10) PMC_CASE(12) PMC_CASE(14) PMC_CASE(16)
    PMC_CASE(18) PMC_CASE(20) PMC_CASE(22) PMC_CASE(24) PMC_CASE(26) PMC_CASE(28) PMC_CASE(30) PMC_CASE(32)
    PMC_CASE(34) PMC_CASE(36) PMC_CASE(38) PMC_CASE(40) PMC_CASE(42) PMC_CASE(44) PMC_CASE(46) PMC_CASE(48)
    PMC_CASE(50) PMC_CASE(52) PMC_CASE(54) PMC_CASE(56) PMC_CASE(58) PMC_CASE(60) PMC_CASE(62) PMC_CASE(64)
#undef PMC_CASE
    default:
      return int(cudaErrorInvalidValue);
  }
}

template <int DP>
struct CfgQuery {
  static int tile() { return EvalCfg<DP>::TS; }
};

int k1_tile_rows(int dp) {
  switch (dp) {
#define PMC_CASE(DP) \
  case DP:           \
    return CfgQuery<DP>::tile();
    PMC_CASE(2) PMC_CASE(4) PMC_CASE(6) PMC_CASE(8) PMC_CASE(10) PMC_CASE(12) PMC_CASE(14) PMC_CASE(16)
    PMC_CASE(18) PMC_CASE(20) PMC_CASE(22) PMC_CASE(24) PMC_CASE(26) PMC_CASE(28) PMC_CASE(30) PMC_CASE(32)
    PMC_CASE(34) PMC_CASE(36) PMC_CASE(38) PMC_CASE(40) PMC_CASE(42) PMC_CASE(44) PMC_CASE(46) PMC_CASE(48)
    PMC_CASE(50) PMC_CASE(52) PMC_CASE(54) PMC_CASE(56) PMC_CASE(58) PMC_CASE(60) PMC_CASE(62) PMC_CASE(64)
#undef PMC_CASE
    default:
      return 0;
  }
}

// sums[0..1] = fixed-order sum of the per-warp partial pairs
__global__ void k1_reduce_sums(double* __restrict__ partials, int count, double* __restrict__ sums) {
  double a = 0.0, w = 0.0;
  for (int i = threadIdx.x; i < count; i += 32) {
    a += partials[2 * i];
    w += partials[2 * i + 1];
    partials[2 * i] = 0.0;      // leave the slots clean for the next launch that shares this workspace
    partials[2 * i + 1] = 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    w += __shfl_xor_sync(0xffffffffu, w, o);
  }
  if (threadIdx.x == 0) {
    sums[0] = a;
    sums[1] = w;
  }
}

}  // namespace pmc

using namespace pmc;

// K2: component-block counts 2..8 (true) or powers of two only (false) by default; see pmcb200_suffstats
#ifndef K2_FINE_BLOCKS_DEFAULT
#define K2_FINE_BLOCKS_DEFAULT 1
#endif

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct PinBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct pmcb200_ctx {
  int device = 0;
  int sm_count = 0;
  DevBuf ws;              // partial sums of K2 / microbenchmark scratch
  DevBuf k1ws;            // K1: derived records, shift, flag, per-warp partial sums (device-pointer entry point)
  DevBuf pws;             // K3: block starts
  DevBuf cws;             // K2: per-CTA partial column sums (gamma)
  DevBuf k1row;           // K1: per-row (max, 1/denominator) handed from k1_fast_eval to k1_finish
  DevBuf wws;             // K4: arrival counter + per-CTA partial sums
  cudaStream_t copy_stream[2] = {nullptr, nullptr};
  // host pipeline: per-slot device buffers
  DevBuf hx[2], hw[2], hlogq[2], hlp[2], hresp[2], haux[2], hws[2], hsums[2], hrow[2];
  DevBuf hrec, hcols;
  // pinned staging for pageable caller memory (host pipeline): per-slot bounce buffers + completion events
  PinBuf px[2], pw[2], plogq[2], plp[2], presp[2], paux[2];
  cudaEvent_t slot_done[2] = {nullptr, nullptr};
  int64_t launches = 0;
  // what the last device-pointer K1 launch offered (pmcb200_last_k1_kernel)
  int last_dp = 0, last_groups = 0, last_cb = 0, last_nb = 0, last_second = 0;
};

static int ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return 0;
  if (b.p) PMC_CUDA_CHECK(cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  PMC_CUDA_CHECK(cudaMalloc(&b.p, bytes));
  PMC_CUDA_CHECK(cudaMemset(b.p, 0, bytes));     // k1_prepare's arrival counter / verdict bits start at zero
  PMC_CUDA_CHECK(cudaDeviceSynchronize());       // (allocation time only) the caller's stream may be non-blocking
  b.bytes = bytes;
  return 0;
}

static int ensure_pinned(PinBuf& b, size_t bytes) {
  if (bytes <= b.bytes) return 0;
  if (b.p) PMC_CUDA_CHECK(cudaFreeHost(b.p));
  b.p = nullptr;
  b.bytes = 0;
  PMC_CUDA_CHECK(cudaMallocHost(&b.p, bytes));
  b.bytes = bytes;
  return 0;
}

// true for page-locked (cudaMallocHost / cudaHostRegister / managed) memory: DMA reaches it directly
static bool host_is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int copy_threads() {
  // PMCB200_COPY_THREADS overrides; default: half the hardware threads, at most 8.  Measured on the 16-vCPU GPU box
  // (scripts/bench_pageable.py, 4e6 x 30 pageable rows through multi_evaluate): 4 threads 24.8 GB/s, 8 threads 32.4,
  // 12: 30.6, 16: 28.3, 24: 25.6 -- the host's memory system, not the thread count, is what stops short of the
  // ~55 GB/s the PCIe link takes from pinned memory
  static const int n = [] {
    const char* env = getenv("PMCB200_COPY_THREADS");
    if (env && atoi(env) > 0) return std::min(64, atoi(env));
    return int(std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2)));
  }();
  return n;
}

// rows x d doubles from a (possibly strided) source into a (possibly strided) destination, split over host threads:
// one thread moves ~10 GB/s, a PCIe 5 x16 link wants ~55 GB/s
static void par_copy_rows(double* dst, int64_t ld_dst, const double* src, int64_t ld_src, int64_t rows, int d) {
  const size_t total = size_t(rows) * d * sizeof(double);
  const int nt = (total < (size_t(1) << 20)) ? 1 : copy_threads();
  auto work = [=](int64_t lo, int64_t hi) {
    if (ld_dst == d && ld_src == d) {
      std::memcpy(dst + lo * d, src + lo * d, size_t(hi - lo) * d * sizeof(double));
    } else {
      for (int64_t r = lo; r < hi; ++r) std::memcpy(dst + r * ld_dst, src + r * ld_src, size_t(d) * sizeof(double));
    }
  };
  if (nt == 1) {
    work(0, rows);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  const int64_t per = (rows + nt - 1) / nt;
  for (int i = 1; i < nt; ++i) {
    const int64_t lo = std::min<int64_t>(rows, i * per), hi = std::min<int64_t>(rows, lo + per);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  work(0, std::min<int64_t>(rows, per));
  for (auto& th : pool) th.join();
}

extern "C" {

int pmcb200_version(void) { return PMCB200_VERSION; }

const char* pmcb200_last_error(void) { return g_last_error.c_str(); }

int pmcb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pmcb200_create(int device, pmcb200_ctx** out) {
  PMC_REQUIRE(out != nullptr, "pmcb200_create: out is NULL");
  int n = 0;
  PMC_CUDA_CHECK(cudaGetDeviceCount(&n));
  PMC_REQUIRE(device >= 0 && device < n, "pmcb200_create: no such CUDA device (this library has no CPU fallback)");
  PMC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  PMC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  PMC_REQUIRE(prop.major >= 10, "pmcb200_create: kernels are built for sm_100a (Blackwell) only");
  pmcb200_ctx* c = new pmcb200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < 2; ++i) PMC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream[i], cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) PMC_CUDA_CHECK(cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming));
  *out = c;
  return 0;
}

int pmcb200_destroy(pmcb200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  DevBuf* all[] = {&c->ws, &c->k1ws, &c->k1row, &c->pws, &c->cws, &c->wws, &c->hrec, &c->hcols};
  for (DevBuf* b : all)
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < 2; ++i) {
    DevBuf* slot[] = {&c->hx[i], &c->hw[i], &c->hlogq[i], &c->hlp[i], &c->hresp[i], &c->haux[i], &c->hws[i], &c->hsums[i], &c->hrow[i]};
    for (DevBuf* b : slot)
      if (b->p) cudaFree(b->p);
    if (c->copy_stream[i]) cudaStreamDestroy(c->copy_stream[i]);
    PinBuf* pins[] = {&c->px[i], &c->pw[i], &c->plogq[i], &c->plp[i], &c->presp[i], &c->paux[i]};
    for (PinBuf* b : pins)
      if (b->p) cudaFreeHost(b->p);
    if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]);
  }
  delete c;
  return 0;
}

int pmcb200_record_len(int d) {
  if (d < 1 || d > PMCB200_MAX_DIM) return -1;
  return record_len((d + 1) & ~1);
}

int pmcb200_pack_record(int d, const double* t, const double* center, const double* scalars, double* rec) {
  PMC_REQUIRE(d >= 1 && d <= PMCB200_MAX_DIM, "pack_record: dimension must be in 1..64");
  PMC_REQUIRE(t && center && scalars && rec, "pack_record: NULL argument");
  const int dp = (d + 1) & ~1, h = dp / 2, nt = tri_len(dp);
  auto T = [&](int i, int j) -> double { return (i < d && j < d && j <= i) ? t[size_t(i) * d + j] : 0.0; };
  for (int r = 0; r < h; ++r)
    for (int p = 0; p <= r; ++p) {
      double* b = rec + 2 * r * (r + 1) + 4 * p;
      b[0] = T(2 * r, 2 * p);
      b[1] = T(2 * r, 2 * p + 1);
      b[2] = T(2 * r + 1, 2 * p);
      b[3] = T(2 * r + 1, 2 * p + 1);
    }
  for (int j = 0; j < dp; ++j) rec[nt + j] = (j < d) ? center[j] : 0.0;
  for (int s = 0; s < kNumScalars; ++s) rec[nt + dp + s] = scalars[s];
  return 0;
}

static int eval_validate(int64_t n, int64_t ldx, int d, int kl, int k_out, int mode, const void* lp, const void* resp,
                         const void* aux) {
  PMC_REQUIRE(d >= 1 && d <= PMCB200_MAX_DIM, "mixture_eval: dimension must be in 1..64");
  PMC_REQUIRE(n >= 0 && ldx >= d, "mixture_eval: bad n / ldx");
  PMC_REQUIRE(kl >= 1, "mixture_eval: at least one component must be evaluated");
  PMC_REQUIRE(mode >= 0 && mode <= 2, "mixture_eval: unknown mode");
  PMC_REQUIRE((!lp && !resp && !aux) || k_out >= 1, "mixture_eval: k_out must be >= 1 when N x K outputs are requested");
  return 0;
}

// One logical K1 launch on stream st: [k1_prepare] -> k1_fast_eval -> k1_mixture_eval (one of the two works, see
// k1_dispatch.cuh) [-> k1_reduce_sums].  `prep` holds the derived records / shift / flag / per-warp partial sums;
// with `prepared` the prepare kernel already ran for these records on this stream (host pipeline: once per call).
static int eval_prepare(pmcb200_ctx* c, DevBuf& prep, const EvalArgs& a, cudaStream_t st, K1Launch* out) {
  const int dp = (a.d + 1) & ~1;
  const int rl = record_len(dp);
  const size_t n_part = size_t(c->sm_count) * PMC_MAX_WARPS * 2;
  // layout (doubles): [flags (4 ints) | derived records | shift | per-warp partials | finish partials | theta]; the flags
  // sit at a FIXED place because k1_prepare's counter / verdict bits must read zero whatever was launched before
  const size_t off_flag = 0, off_rec = 2, off_shift = off_rec + size_t(a.kl) * rl, off_part = off_shift + PMC_MAX_DP;
  const size_t off_fin = off_part + n_part, n_fin = size_t(c->sm_count) * 8;
  // matrix-instruction form: offered unless PMCB200_K1_FORM=dfma (comparison runs), k1_prepare has the last word
  K1Launch::MmaPlan plan;
  const char* form_env = getenv("PMCB200_K1_FORM");
  const bool want_mma = !(form_env && std::string(form_env) == "dfma") && k1_mma_plan(a.kl, a.d, &plan);
  const size_t off_theta = off_fin + n_fin, n_theta = want_mma ? plan.theta_len : 0;
  if (int rc = ensure(prep, (off_theta + n_theta) * sizeof(double))) return rc;
  double* base = static_cast<double*>(prep.p);
  k1_prepare<<<a.kl, 128, 0, st>>>(a.records, a.kl, dp, base + off_rec, base + off_shift, reinterpret_cast<int*>(base + off_flag),
                                base + off_part, int(n_part), want_mma ? 1 : 0);
  PMC_CUDA_CHECK(cudaGetLastError());
  c->launches++;
  out->mma = K1Launch::MmaPlan();
  if (want_mma) {
    for (int g = 0; g < plan.groups; ++g) {            // theta per component group
      k1_mma_prepare<<<8 * plan.cb[g], 256, 0, st>>>(base + off_rec + size_t(plan.k0[g]) * rl, a.records + size_t(plan.k0[g]) * rl,
                                                     base + off_shift, plan.count[g], 8 * plan.cb[g], a.d, dp, plan.steps,
                                                     a.mode, base + off_theta + plan.theta_off[g],
                                                     reinterpret_cast<int*>(base + off_flag));
      PMC_CUDA_CHECK(cudaGetLastError());
      c->launches++;
    }
    out->theta = base + off_theta;
    out->mma = plan;
  }
  out->derived = base + off_rec;
  out->shift = base + off_shift;
  out->flag = reinterpret_cast<int*>(base + off_flag);
  out->base = a;
  out->base.partials = base + off_part;
  out->rowstat = base + off_fin;       // (re-used as the pointer to the finish kernel's per-block partial sums)
  return 0;
}

static int eval_launch(pmcb200_ctx* c, const K1Launch& prep, DevBuf& rowbuf, const EvalArgs& a0, double* sums_dev,
                       cudaStream_t st) {
  K1Launch l = prep;   // derived / shift / flag of the prepared records
  double* partials = prep.base.partials;
  double* fin_partials = prep.rowstat;
  l.base = a0;
  l.base.partials = sums_dev ? partials : nullptr;
  const bool second_pass = (a0.resp_out != nullptr) || (a0.mode == MODE_VB && a0.lp_out != nullptr);
  l.rowstat = nullptr;
  if (second_pass || l.mma.groups > 1) {
    if (int rc = ensure(rowbuf, size_t(a0.n) * 2 * sizeof(double))) return rc;
    l.rowstat = static_cast<double*>(rowbuf.p);
  }
  const int dp = (a0.d + 1) & ~1;
  const int ts = k1_tile_rows(dp);
  const int64_t tiles = (a0.n + ts - 1) / ts;
  const int grid = int(std::min<int64_t>(tiles, c->sm_count));
  if (&rowbuf == &c->k1row) {
    const bool fused = second_pass && l.mma.groups == 1;
    c->last_dp = dp; c->last_groups = l.mma.groups; c->last_cb = l.mma.groups ? l.mma.cb[0] : 0;
    c->last_nb = l.mma.groups ? k1_mma_nb(l.mma.cb[0], fused) : 0; c->last_second = fused ? 1 : 0;
  }
  if (l.mma.groups > 0) {
    const int em = k1_mma_launch(l, c->sm_count, st);
    if (em != 0) {
      set_last_error(std::string("k1 mma launch: ") + cudaGetErrorString(cudaError_t(em)));
      return 1;
    }
    c->launches += l.mma.groups;
  }
  const int e = k1_launch(dp, l, grid, st);
  if (e != 0) {
    set_last_error(std::string("k1 launch: ") + cudaGetErrorString(cudaError_t(e)));
    return 1;
  }
  c->launches += 2;
  const int fin_grid = c->sm_count * 8;
  const bool fin_sum = second_pass && sums_dev && a0.mode == MODE_VB;
  if (second_pass) {
    FinishArgs f;
    f.n = a0.n; f.kl = a0.kl; f.k_out = a0.k_out; f.mode = a0.mode;
    f.rl = record_len(dp); f.w_off = tri_len(dp) + dp + S_WEIGHT;
    f.records = l.derived; f.cols = a0.cols; f.rowstat = l.rowstat; f.sw = a0.sw;
    f.scratch = a0.lp_out ? a0.lp_out : a0.resp_out; f.lp_out = a0.lp_out; f.resp_out = a0.resp_out;
    f.flag = l.flag; f.fin_partials = fin_sum ? fin_partials : nullptr;
    f.mma_fused = (l.mma.groups == 1) ? 1 : 0;
    k1_finish<<<fin_grid, 256, 0, st>>>(f);
    PMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
  }
  if (sums_dev) {
    // partial slots of CTAs / warps that did not run hold the zeros written by k1_prepare (or by the last reduce)
    k1_reduce_sums<<<1, 32, 0, st>>>(partials, c->sm_count * PMC_MAX_WARPS, sums_dev);
    PMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
    if (fin_sum) {
      k1_reduce_finish<<<1, 32, 0, st>>>(fin_partials, fin_grid, l.flag, (l.mma.groups == 1) ? 1 : 0, sums_dev);
      PMC_CUDA_CHECK(cudaGetLastError());
      c->launches++;
    }
  }
  return 0;
}

int pmcb200_mixture_eval(pmcb200_ctx* c, const double* x, int64_t n, int64_t ldx, int d, const double* records,
                         const int* cols, int kl, int k_out, int mode, double max_init, double* logq, double* lp,
                         double* resp, double* aux, const double* weights, double* sums, void* stream) {
  PMC_REQUIRE(c != nullptr, "mixture_eval: NULL context");
  if (int rc = eval_validate(n, ldx, d, kl, k_out, mode, lp, resp, aux)) return rc;
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    if (sums) PMC_CUDA_CHECK(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    return 0;
  }
  PMC_REQUIRE(x && records && cols, "mixture_eval: NULL input");
  EvalArgs a{x, n, ldx, d, records, cols, kl, k_out, mode, max_init, logq, lp, resp, aux, weights, nullptr, nullptr};
  K1Launch prep;
  if (int rc = eval_prepare(c, c->k1ws, a, st, &prep)) return rc;
  return eval_launch(c, prep, c->k1row, a, sums, st);
}

int pmcb200_suffstats(pmcb200_ctx* c, const double* x, int64_t n, int64_t ldx, int d, const double* shift,
                      const double* rho, const double* gamma, int k, int ld_rho, const double* weights, double* out,
                      void* stream) {
  PMC_REQUIRE(c != nullptr, "suffstats: NULL context");
  PMC_REQUIRE(d >= 1 && d <= 256, "suffstats: dimension must be in 1..256");
  PMC_REQUIRE(k >= 1 && ld_rho >= k && n >= 0 && ldx >= d, "suffstats: bad sizes");
  PMC_REQUIRE(out != nullptr, "suffstats: NULL output");
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int F = 1 + d + d * (d + 1) / 2;
  const int64_t len = int64_t(k) * (F + 2);
  if (n == 0) {
    PMC_CUDA_CHECK(cudaMemsetAsync(out, 0, len * sizeof(double), st));
    return 0;
  }
  PMC_REQUIRE(x && shift && rho, "suffstats: NULL input");
  StatsArgs a;
  a.x = x; a.n = n; a.ldx = ldx; a.d = d; a.shift = shift; a.rho = rho; a.gamma = gamma; a.sw = weights;
  a.k = k; a.ld_rho = ld_rho; a.F = F;
  a.P0 = (d + 1) / 2;
  a.Bq = a.P0 * (a.P0 + 1) / 2;
  a.Lq = (d + 4) / 4;
  a.KP = ((k + K2_TK - 1) / K2_TK) * K2_TK;
  a.LT = (a.KP / K2_TK) * (a.Bq + a.Lq);
  // consumer form: FP64 matrix instructions (default) or the DFMA register tile (PMCB200_K2_FORM=dfma)
  static const char* form_env = getenv("PMCB200_K2_FORM");
  const bool mma = !(form_env && std::string(form_env) == "dfma");
  int gy, cbw = 0, fbw = 0;
  if (mma) {
    // component blocks of 8 per warp tile (CB x FB, at most 32 C tiles = 64 accumulators): the blocks that hold
    // components, split evenly over the CTAs that share them; PMCB200_K2_BLOCKS=pow2 restores the earlier choice
    // (next power of two), kept for comparison runs
    static const char* blocks_env = getenv("PMCB200_K2_BLOCKS");
    const bool fine = K2_FINE_BLOCKS_DEFAULT ? !(blocks_env && std::string(blocks_env) == "pow2")
                                             : (blocks_env && std::string(blocks_env) == "fine");
    int cb_total = (k + 7) / 8, cchunks = (cb_total + 7) / 8;
    cbw = std::max(2, (cb_total + cchunks - 1) / cchunks);
    if (!fine) {
      cb_total = a.KP / 8;
      cbw = (cb_total <= 2) ? 2 : (cb_total <= 4) ? 4 : 8;
      cchunks = (cb_total + cbw - 1) / cbw;
    }
    a.nFB = (F + 7) / 8;
    const int nw = K2_CONSUMERS / 32, fb_max = 32 / cbw;
    a.fchunks = (a.nFB + nw * fb_max - 1) / (nw * fb_max);
    const int need = (a.nFB + nw * a.fchunks - 1) / (nw * a.fchunks);   // feature blocks per warp, spread evenly
    // instantiated FB values per CB (the smallest one >= need is used; blocks beyond the warp's share are zero work
    // that is still issued, so a tight FB matters: D=40 gives 108 blocks = 13.5 per warp -> FB 14, not 16)
    static const int fb2[] = {2, 4, 6, 8, 10, 12, 14, 16}, fb4[] = {1, 2, 3, 4, 5, 6, 7, 8}, fb8[] = {1, 2, 3, 4};
    static const int fb3[] = {2, 4, 6, 8, 10}, fb5[] = {2, 3, 4, 5, 6}, fb6[] = {2, 3, 4, 5}, fb7[] = {2, 3, 4};
    const int* tab = (cbw == 2) ? fb2 : (cbw == 3) ? fb3 : (cbw == 4) ? fb4 : (cbw == 5) ? fb5 : (cbw == 6) ? fb6 : (cbw == 7) ? fb7 : fb8;
    const int ntab = (cbw == 2) ? 8 : (cbw == 3) ? 5 : (cbw == 4) ? 8 : (cbw == 5) ? 5 : (cbw == 6) ? 4 : (cbw == 7) ? 3 : 4;
    fbw = tab[ntab - 1];
    for (int i = 0; i < ntab; ++i)
      if (tab[i] >= need) { fbw = tab[i]; break; }
    a.fbw = fbw;
    gy = a.fchunks * cchunks;
    a.DP4 = ((d + 2 + 3) / 4) * 4;                       // [y, 1, 0...]: always at least one zero column (index d+1)
    a.chunked = (cchunks > 1) ? 1 : 0;                   // several CTAs share the component blocks: each stages its own
    a.cb8 = cbw * 8;
    a.kchunk = ((a.cb8 + 15) / 16) * 16;
    a.VS = a.chunked ? a.kchunk : a.KP;                  // transposed stage, blocks of 8 samples: V [tn/8][VS][8] | Y [tn/8][YS][8]
    a.YS = a.DP4;                                        // (64 bytes per column: neighbouring columns fall in opposite bank halves)
  } else {
    a.nFB = 0;
    a.fchunks = 1;
    a.fbw = 0;
    a.chunked = 0; a.cb8 = 0; a.kchunk = 0;
    gy = (a.LT + K2_CONSUMERS - 1) / K2_CONSUMERS;       // lane tiles -> CTAs of 8 consumer warps
    a.DP4 = a.Lq * 4;
    a.VS = a.KP;
    a.YS = a.DP4;
  }
  const size_t per_row = sizeof(double) * (size_t(a.VS) + a.YS);
  const size_t fixed = sizeof(double) * a.DP4 + 2 * K2_STAGES * sizeof(uint64_t) + 128;
  int tn = int((size_t(210) * 1024 - fixed) / (K2_STAGES * per_row));
  static const char* tn_env = getenv("PMCB200_K2_TN");             // tuning runs: rows per stage (default: at most 96)
  tn = std::min(tn_env ? std::max(8, atoi(tn_env)) : 96, tn) & ~7;
  PMC_REQUIRE(tn >= 8, "suffstats: K and D too large for the shared-memory pipeline");
  a.tn = tn;
  const size_t smem = K2_STAGES * per_row * tn + fixed;
  const int64_t tiles = (n + tn - 1) / tn;
  const int gx = int(std::max<int64_t>(1, std::min<int64_t>(tiles, std::max(1, c->sm_count / gy))));
  if (int rc = ensure(c->ws, size_t(gx) * len * sizeof(double))) return rc;
  a.partial = static_cast<double*>(c->ws.p);
  auto launch = [&](auto kernel, PerDeviceFlag& flag) -> int {
    bool& attr_set = flag.here();
    if (!attr_set) {
      PMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set = true;
    }
    kernel<<<dim3(gx, gy), K2_THREADS, smem, st>>>(a);
    PMC_CUDA_CHECK(cudaGetLastError());
    return 0;
  };
  static PerDeviceFlag flags[32];
  int rc = 2;
  if (!mma) rc = launch(k2_suffstats<0, 1>, flags[0]);
#define PMC_K2_CASE(CBV, FBV, IDX) else if (cbw == CBV && fbw == FBV) rc = launch(k2_suffstats<CBV, FBV>, flags[IDX]);
  PMC_K2_CASE(2, 2, 1) PMC_K2_CASE(2, 4, 2) PMC_K2_CASE(2, 6, 3) PMC_K2_CASE(2, 8, 4) PMC_K2_CASE(2, 10, 5)
  PMC_K2_CASE(2, 12, 6) PMC_K2_CASE(2, 14, 7) PMC_K2_CASE(2, 16, 8)
  PMC_K2_CASE(4, 1, 9) PMC_K2_CASE(4, 2, 10) PMC_K2_CASE(4, 3, 11) PMC_K2_CASE(4, 4, 12) PMC_K2_CASE(4, 5, 13)
  PMC_K2_CASE(4, 6, 14) PMC_K2_CASE(4, 7, 15) PMC_K2_CASE(4, 8, 16)
  PMC_K2_CASE(8, 1, 17) PMC_K2_CASE(8, 2, 18) PMC_K2_CASE(8, 3, 19) PMC_K2_CASE(8, 4, 20)
#undef PMC_K2_CASE
  if (rc == 2 && mma) {
    rc = k2_launch_extra(cbw, fbw, a, dim3(gx, gy), smem, st);
    if (rc == 1) set_last_error("suffstats: launch of a k2_inst.cu instantiation failed");
  }
  PMC_REQUIRE(rc != 2, "suffstats: no kernel instantiation for this tile shape");
  if (rc) return rc;
  k2_reduce_partials<<<unsigned((len + 255) / 256), 256, 0, st>>>(a.partial, gx, len, out);
  PMC_CUDA_CHECK(cudaGetLastError());
  c->launches += 2;
  // columns 0 (A, with gamma) and F+1 (L) of the rows: a streaming pass of its own (k2_colsums)
  const int cgrid = gamma ? int(std::min<int64_t>(int64_t(c->sm_count) * 8, std::max<int64_t>(1, n / 64))) : 0;
  if (gamma) {
    int kc = 1;
    while (kc < k && kc < 256) kc <<= 1;
    if (int rc2 = ensure(c->cws, size_t(cgrid) * 2 * k * sizeof(double))) return rc2;
    k2_colsums<<<cgrid, 256, 0, st>>>(rho, gamma, weights, n, k, ld_rho, kc, static_cast<double*>(c->cws.p));
    PMC_CUDA_CHECK(cudaGetLastError());
    c->launches++;
  }
  k2_colsums_final<<<unsigned(k), 256, 0, st>>>(static_cast<const double*>(c->cws.p), cgrid, k, F + 2, F,
                                                               gamma ? 1 : 0, out);
  PMC_CUDA_CHECK(cudaGetLastError());
  c->launches++;
  return 0;
}

int pmcb200_mixture_eval_host(pmcb200_ctx* c, const double* x, int64_t n, int64_t ldx, int d, const double* records,
                              const int* cols, int kl, int k_out, int mode, double max_init, double* logq, double* lp,
                              double* resp, double* aux, const double* weights, double* sums, int64_t chunk_rows) {
  PMC_REQUIRE(c != nullptr, "mixture_eval_host: NULL context");
  if (int rc = eval_validate(n, ldx, d, kl, k_out, mode, lp, resp, aux)) return rc;
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  if (sums) sums[0] = sums[1] = 0.0;
  if (n == 0) return 0;
  PMC_REQUIRE(x && records && cols, "mixture_eval_host: NULL input");
  const int dp = (d + 1) & ~1;
  const int rl = record_len(dp);
  if (chunk_rows <= 0) chunk_rows = int64_t(1) << 18;
  chunk_rows = std::min(chunk_rows, n);
  const bool need_scratch = (resp != nullptr) || (mode == MODE_VB && lp != nullptr);

  if (int rc = ensure(c->hrec, size_t(kl) * rl * sizeof(double))) return rc;
  if (int rc = ensure(c->hcols, size_t(kl) * sizeof(int))) return rc;
  PMC_CUDA_CHECK(cudaMemcpyAsync(c->hrec.p, records, size_t(kl) * rl * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream[0]));
  PMC_CUDA_CHECK(cudaMemcpyAsync(c->hcols.p, cols, size_t(kl) * sizeof(int), cudaMemcpyHostToDevice, c->copy_stream[0]));
  PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[0]));

  const size_t nk_bytes = size_t(chunk_rows) * std::max(k_out, 1) * sizeof(double);
  for (int s = 0; s < 2; ++s) {
    if (int rc = ensure(c->hx[s], size_t(chunk_rows) * d * sizeof(double))) return rc;
    if (weights)
      if (int rc = ensure(c->hw[s], size_t(chunk_rows) * sizeof(double))) return rc;
    if (int rc = ensure(c->hlogq[s], size_t(chunk_rows) * sizeof(double))) return rc;
    if (lp || (need_scratch && !resp))
      if (int rc = ensure(c->hlp[s], nk_bytes)) return rc;
    if (resp)
      if (int rc = ensure(c->hresp[s], nk_bytes)) return rc;
    if (aux)
      if (int rc = ensure(c->haux[s], nk_bytes)) return rc;
    if (int rc = ensure(c->hsums[s], 2 * sizeof(double))) return rc;
  }

  const int64_t nchunks = (n + chunk_rows - 1) / chunk_rows;
  std::vector<double> chunk_sums(size_t(nchunks) * 2, 0.0);
  K1Launch prep[2];
  for (int s = 0; s < 2 && s < nchunks; ++s) {   // derived records / shift / flag once per stream, not per chunk
    EvalArgs a{nullptr, 0, d, d, static_cast<const double*>(c->hrec.p), static_cast<const int*>(c->hcols.p), kl, k_out,
               mode, max_init, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (int rc = eval_prepare(c, c->hws[s], a, c->copy_stream[s], &prep[s])) return rc;
  }
  // Pageable caller memory (an ordinary numpy array) reaches the device at ~10 GB/s through the driver's own staging;
  // here it is copied by several host threads into pinned bounce buffers first (and results back out the same way),
  // which keeps the PCIe link busy.  Page-locked caller memory is used in place.
  const bool st_x = !host_is_pinned(x), st_w = weights && !host_is_pinned(weights);
  const bool st_q = logq && !host_is_pinned(logq), st_lp = lp && !host_is_pinned(lp);
  const bool st_r = resp && !host_is_pinned(resp), st_a = aux && !host_is_pinned(aux);
  const bool any_out_staged = st_q || st_lp || st_r || st_a;
  for (int s = 0; s < 2; ++s) {
    if (st_x) if (int rc = ensure_pinned(c->px[s], size_t(chunk_rows) * d * sizeof(double))) return rc;
    if (st_w) if (int rc = ensure_pinned(c->pw[s], size_t(chunk_rows) * sizeof(double))) return rc;
    if (st_q) if (int rc = ensure_pinned(c->plogq[s], size_t(chunk_rows) * sizeof(double))) return rc;
    if (st_lp) if (int rc = ensure_pinned(c->plp[s], nk_bytes)) return rc;
    if (st_r) if (int rc = ensure_pinned(c->presp[s], nk_bytes)) return rc;
    if (st_a) if (int rc = ensure_pinned(c->paux[s], nk_bytes)) return rc;
  }
  struct Pending { int64_t r0 = 0, rows = 0; bool valid = false; } pend[2];
  auto drain = [&](int s) -> int {       // wait for the slot's chunk and move its staged outputs to the caller's arrays
    if (!pend[s].valid) return 0;
    PMC_CUDA_CHECK(cudaEventSynchronize(c->slot_done[s]));
    const int64_t r0 = pend[s].r0, rows = pend[s].rows;
    if (st_q) par_copy_rows(logq + r0, 1, static_cast<const double*>(c->plogq[s].p), 1, rows, 1);
    if (st_lp) par_copy_rows(lp + r0 * k_out, k_out, static_cast<const double*>(c->plp[s].p), k_out, rows, k_out);
    if (st_r) par_copy_rows(resp + r0 * k_out, k_out, static_cast<const double*>(c->presp[s].p), k_out, rows, k_out);
    if (st_a) par_copy_rows(aux + r0 * k_out, k_out, static_cast<const double*>(c->paux[s].p), k_out, rows, k_out);
    pend[s].valid = false;
    return 0;
  };
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int s = int(ci & 1);
    cudaStream_t st = c->copy_stream[s];
    const int64_t r0 = ci * chunk_rows, rows = std::min(chunk_rows, n - r0);
    // the slot's previous chunk (ci-2) was queued on the same stream, so stream order protects the DEVICE buffers;
    // the pinned bounce buffers are host-written, so their previous chunk must have completed
    if (int rc = drain(s)) return rc;
    if (st_x) {
      par_copy_rows(static_cast<double*>(c->px[s].p), d, x + r0 * ldx, ldx, rows, d);
      PMC_CUDA_CHECK(cudaMemcpyAsync(c->hx[s].p, c->px[s].p, size_t(rows) * d * sizeof(double), cudaMemcpyHostToDevice, st));
    } else if (ldx == d) {
      PMC_CUDA_CHECK(cudaMemcpyAsync(c->hx[s].p, x + r0 * ldx, size_t(rows) * d * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
      PMC_CUDA_CHECK(cudaMemcpy2DAsync(c->hx[s].p, size_t(d) * sizeof(double), x + r0 * ldx, size_t(ldx) * sizeof(double),
                                       size_t(d) * sizeof(double), size_t(rows), cudaMemcpyHostToDevice, st));
    }
    if (weights) {
      const double* wsrc = weights + r0;
      if (st_w) {
        std::memcpy(c->pw[s].p, wsrc, size_t(rows) * sizeof(double));
        wsrc = static_cast<const double*>(c->pw[s].p);
      }
      PMC_CUDA_CHECK(cudaMemcpyAsync(c->hw[s].p, wsrc, size_t(rows) * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    EvalArgs a{static_cast<const double*>(c->hx[s].p), rows, d, d,
               static_cast<const double*>(c->hrec.p), static_cast<const int*>(c->hcols.p), kl, k_out, mode, max_init,
               static_cast<double*>(c->hlogq[s].p),
               lp ? static_cast<double*>(c->hlp[s].p) : nullptr,
               resp ? static_cast<double*>(c->hresp[s].p) : nullptr,
               aux ? static_cast<double*>(c->haux[s].p) : nullptr,
               weights ? static_cast<const double*>(c->hw[s].p) : nullptr, nullptr, nullptr};
    if (int rc = eval_launch(c, prep[s], c->hrow[s], a, sums ? static_cast<double*>(c->hsums[s].p) : nullptr, st)) return rc;
    const size_t out_bytes = size_t(rows) * k_out * sizeof(double);
    if (logq)
      PMC_CUDA_CHECK(cudaMemcpyAsync(st_q ? c->plogq[s].p : static_cast<void*>(logq + r0), c->hlogq[s].p,
                                     size_t(rows) * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (lp)
      PMC_CUDA_CHECK(cudaMemcpyAsync(st_lp ? c->plp[s].p : static_cast<void*>(lp + r0 * k_out), c->hlp[s].p, out_bytes,
                                     cudaMemcpyDeviceToHost, st));
    if (resp)
      PMC_CUDA_CHECK(cudaMemcpyAsync(st_r ? c->presp[s].p : static_cast<void*>(resp + r0 * k_out), c->hresp[s].p, out_bytes,
                                     cudaMemcpyDeviceToHost, st));
    if (aux)
      PMC_CUDA_CHECK(cudaMemcpyAsync(st_a ? c->paux[s].p : static_cast<void*>(aux + r0 * k_out), c->haux[s].p, out_bytes,
                                     cudaMemcpyDeviceToHost, st));
    if (sums)
      PMC_CUDA_CHECK(cudaMemcpyAsync(&chunk_sums[size_t(ci) * 2], c->hsums[s].p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (st_x || st_w || any_out_staged) {
      PMC_CUDA_CHECK(cudaEventRecord(c->slot_done[s], st));
      pend[s] = Pending{r0, rows, true};
    }
  }
  for (int64_t ci = std::max<int64_t>(0, nchunks - 2); ci < nchunks; ++ci)   // oldest first
    if (int rc = drain(int(ci & 1))) return rc;
  PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[0]));
  PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[1]));
  if (sums)
    for (int64_t ci = 0; ci < nchunks; ++ci) {
      sums[0] += chunk_sums[size_t(ci) * 2];
      sums[1] += chunk_sums[size_t(ci) * 2 + 1];
    }
  return 0;
}

int pmcb200_upload(pmcb200_ctx* c, double* dst, const double* src, int64_t rows, int d, int64_t ld_src) {
  PMC_REQUIRE(c != nullptr, "upload: NULL context");
  PMC_REQUIRE(rows >= 0 && d >= 1 && ld_src >= d, "upload: bad sizes");
  if (rows == 0) return 0;
  PMC_REQUIRE(dst && src, "upload: NULL pointer");
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  if (host_is_pinned(src)) {
    PMC_CUDA_CHECK(cudaMemcpy2DAsync(dst, size_t(d) * sizeof(double), src, size_t(ld_src) * sizeof(double),
                                     size_t(d) * sizeof(double), size_t(rows), cudaMemcpyHostToDevice, c->copy_stream[0]));
    PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[0]));
    return 0;
  }
  const int64_t chunk = std::max<int64_t>(1, (int64_t(64) << 20) / (int64_t(d) * 8));   // ~64 MB per bounce buffer
  for (int s = 0; s < 2; ++s)
    if (int rc = ensure_pinned(c->px[s], size_t(std::min(chunk, rows)) * d * sizeof(double))) return rc;
  bool busy[2] = {false, false};
  int64_t ci = 0;
  for (int64_t r0 = 0; r0 < rows; r0 += chunk, ++ci) {
    const int s = int(ci & 1);
    const int64_t n = std::min(chunk, rows - r0);
    if (busy[s]) PMC_CUDA_CHECK(cudaEventSynchronize(c->slot_done[s]));     // the bounce buffer is free again
    par_copy_rows(static_cast<double*>(c->px[s].p), d, src + r0 * ld_src, ld_src, n, d);
    PMC_CUDA_CHECK(cudaMemcpyAsync(dst + r0 * d, c->px[s].p, size_t(n) * d * sizeof(double), cudaMemcpyHostToDevice,
                                   c->copy_stream[s]));
    PMC_CUDA_CHECK(cudaEventRecord(c->slot_done[s], c->copy_stream[s]));
    busy[s] = true;
  }
  PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[0]));
  PMC_CUDA_CHECK(cudaStreamSynchronize(c->copy_stream[1]));
  return 0;
}

int pmcb200_mixture_propose(pmcb200_ctx* c, int64_t n, int d, int k, const double* means, const double* chol,
                            const double* dofs, const int64_t* starts_host, uint64_t seed, uint64_t index0, double* x,
                            int64_t ldx, int* latent, void* stream) {
  PMC_REQUIRE(c != nullptr, "mixture_propose: NULL context");
  PMC_REQUIRE(d >= 1 && d <= 256 && k >= 1 && n >= 0 && ldx >= d, "mixture_propose: bad sizes");
  PMC_REQUIRE(starts_host != nullptr, "mixture_propose: NULL starts");
  PMC_REQUIRE(starts_host[0] == 0 && starts_host[k] == n, "mixture_propose: starts must run from 0 to n");
  for (int i = 0; i < k; ++i) PMC_REQUIRE(starts_host[i] <= starts_host[i + 1], "mixture_propose: starts must not decrease");
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) return 0;
  PMC_REQUIRE(means && chol && x, "mixture_propose: NULL input");
  if (int rc = ensure(c->pws, size_t(k + 1) * sizeof(int64_t))) return rc;
  // pageable source: the copy is staged before the call returns, so the caller's array may die at once
  PMC_CUDA_CHECK(cudaMemcpyAsync(c->pws.p, starts_host, size_t(k + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  ProposeArgs a{n, ldx, d, k, means, chol, dofs, static_cast<const int64_t*>(c->pws.p), seed, index0, x, latent};
  const size_t smem = k3_smem_bytes(d, k);
  PMC_REQUIRE(smem <= 200 * 1024, "mixture_propose: too many components for the shared-memory table");
  const int64_t blocks = (n + K3_THREADS - 1) / K3_THREADS;
  const int grid = int(std::min<int64_t>(blocks, int64_t(c->sm_count) * 8));
  static PerDeviceFlag attr_flags[6];
  auto launch = [&](auto kernel, PerDeviceFlag& flag) -> int {
    bool& attr_set = flag.here();
    if (!attr_set) {
      PMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    kernel<<<grid, K3_THREADS, smem, st>>>(a);
    PMC_CUDA_CHECK(cudaGetLastError());
    return 0;
  };
  // register-resident form up to D = 40 (PMCB200_K3_FORM=smem keeps the shared-memory rows, for comparison runs)
  static const char* k3_env = getenv("PMCB200_K3_FORM");
  const bool regs = !(k3_env && std::string(k3_env) == "smem");
  int rc;
  if (regs && d <= 8) rc = launch(k3_propose<8>, attr_flags[1]);
  else if (regs && d <= 16) rc = launch(k3_propose<16>, attr_flags[2]);
  else if (regs && d <= 24) rc = launch(k3_propose<24>, attr_flags[3]);
  else if (regs && d <= 32) rc = launch(k3_propose<32>, attr_flags[4]);
  else if (regs && d <= 40) rc = launch(k3_propose<40>, attr_flags[5]);
  else rc = launch(k3_propose<0>, attr_flags[0]);
  if (rc) return rc;
  c->launches++;
  return 0;
}

int pmcb200_importance_weights(pmcb200_ctx* c, const double* log_target, const double* logq, int64_t n, double* w,
                               double* sums, void* stream) {
  PMC_REQUIRE(c != nullptr, "importance_weights: NULL context");
  PMC_REQUIRE(n >= 0 && sums != nullptr, "importance_weights: bad arguments");
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    PMC_CUDA_CHECK(cudaMemsetAsync(sums, 0, K4_SUMS * sizeof(double), st));
    return 0;
  }
  PMC_REQUIRE(logq != nullptr, "importance_weights: NULL input");
  const int grid = int(std::min<int64_t>((n + K4_THREADS - 1) / K4_THREADS, int64_t(c->sm_count) * 8));
  if (int rc = ensure(c->wws, 16 + size_t(c->sm_count) * 8 * K4_SUMS * sizeof(double))) return rc;   // zeroed at allocation
  unsigned int* counter = static_cast<unsigned int*>(c->wws.p);
  double* partials = reinterpret_cast<double*>(static_cast<char*>(c->wws.p) + 16);
  k4_weights<<<grid, K4_THREADS, 0, st>>>(log_target, logq, n, w, partials, counter, sums);
  PMC_CUDA_CHECK(cudaGetLastError());
  c->launches++;
  return 0;
}

int pmcb200_fp64_peak(pmcb200_ctx* c, int which, int iters, double* gflops_out, double* ms_out) {
  PMC_REQUIRE(c != nullptr, "fp64_peak: NULL context");
  PMC_REQUIRE(which >= 0 && which <= 3 && iters > 0, "fp64_peak: bad arguments");
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  if (int rc = ensure(c->ws, 64)) return rc;
  double* out = static_cast<double*>(c->ws.p);
  const int grid = c->sm_count * 2;
  cudaEvent_t e0, e1;
  PMC_CUDA_CHECK(cudaEventCreate(&e0));
  PMC_CUDA_CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
    PMC_CUDA_CHECK(cudaEventRecord(e0, 0));
    switch (which) {
      case 0: mb_dfma<0><<<grid, MB_THREADS>>>(iters, 1.0, out); break;
      case 1: mb_dfma<1><<<grid, MB_THREADS>>>(iters, 1.0, out); break;
      case 2: mb_dfma<2><<<grid, MB_THREADS>>>(iters, 1.0, out); break;
      default: mb_dmma<<<grid, MB_THREADS>>>(iters, 1.0, out); break;
    }
    PMC_CUDA_CHECK(cudaGetLastError());
    PMC_CUDA_CHECK(cudaEventRecord(e1, 0));
    PMC_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    PMC_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
    c->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // flops: FMA = 2 flop.  DFMA kernels: 64 FMA / thread / iter.  DMMA: 64 mma / warp / iter, 8*8*4 FMA each.
  const double threads = double(grid) * MB_THREADS;
  const double fma = (which == 3) ? (threads / 32.0) * 64.0 * 256.0 * iters : threads * 64.0 * iters;
  if (gflops_out) *gflops_out = 2.0 * fma / (best * 1e-3) * 1e-9;
  if (ms_out) *ms_out = best;
  return 0;
}

int pmcb200_last_k1_kernel(pmcb200_ctx* c, char* buf, int len) {
  PMC_REQUIRE(c != nullptr && buf != nullptr && len > 0, "last_k1_kernel: bad arguments");
  buf[0] = 0;
  if (!c->k1ws.p || c->last_dp == 0) return 0;                       // nothing launched yet
  PMC_CUDA_CHECK(cudaSetDevice(c->device));
  PMC_CUDA_CHECK(cudaDeviceSynchronize());
  int flag[2] = {0, 0};
  PMC_CUDA_CHECK(cudaMemcpy(flag, c->k1ws.p, sizeof(flag), cudaMemcpyDeviceToHost));
  if (flag[0] != 0) snprintf(buf, size_t(len), "k1_mixture_eval<%d>", c->last_dp);
  else if (flag[1] != 0 && c->last_groups > 0)
    snprintf(buf, size_t(len), "k1_mma_eval<%d, %d, 16, %s>%s", c->last_cb, c->last_nb, c->last_second ? "true" : "false",
             c->last_groups > 1 ? " (component groups)" : "");
  else snprintf(buf, size_t(len), "k1_fast_eval<%d>", c->last_dp);
  return 0;
}

int64_t pmcb200_launch_count(pmcb200_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
