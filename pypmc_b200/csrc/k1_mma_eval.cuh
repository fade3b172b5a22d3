// k1_mma_eval.cuh -- K1, matrix-instruction form: the same fused per-sample x per-component log-pdf + mixture
// log-sum-exp as k1_fast_eval.cuh (same outputs, same reference citations: gauss.pyx:146-151, student_t.pyx:154-164,
// mixture.pyx:112-156, _regularize.pyx:57-83, pmc.pyx:23-43, variational.pyx:774-798), issued as FP64 matrix
// instructions (mma.sync.m8n8k4.f64, SASS DMMA).
//
// Why a third form: the DFMA forms read three 64-bit register operands per FMA and the register file delivers about
// one per clock, which caps them at 66-83 % of the FP64 pipe (profiles/r01_operand_delivery.md; k1_fast_eval measured
// 70 %).  A DMMA moves 256 FMAs with four operand registers per thread (K2 reached 87 % that way).  The triangular
// solve itself maps badly on 8x4 blocks (+25 % FMAs), so the quadratic form is expanded instead:
//     q_nk = (x' - d_k)^T M_k (x' - d_k) = theta_k . phi(x'),      x' = x - c,  d_k = mu_k - c,  M_k = T_k^T T_k
//     phi(x')  = [ 1 | x'_i | x'_i x'_j (j <= i) ]                 F = 1 + D + D(D+1)/2 features (496 at D=30)
//     theta_k  = [ |b_k|^2 | -2 (T_k^T b_k)_i | M_ii, 2 M_ij ]     b_k = T_k d_k       (k1_mma_prepare)
// i.e. a dense (N x F).(F x K) product with F (D(D+1)/2 + D + 1) FMAs per pair -- the algorithmic count, nothing
// padded but the last feature quad -- whose left operand is formed on the fly (one DMUL per fragment).
// Rounding: the expanded terms cancel, so the absolute error of q is about sqrt(F) eps A_k with A_k = sum_f
// |theta_kf| |phi_f(d_k)|, the size of the terms at the component's own centre (= |b_k|^2 for a well-conditioned
// component, up to kappa times more when d_k lies along a wide axis of an ill-conditioned covariance).
// k1_mma_prepare admits this form only while max_k s_k A_k <= kMmaMaxBias2 (s_k = the factor q carries in the
// exponent; error of the log-pdf below ~3e-11), otherwise k1_fast_eval (|b| <= 1e4) or the exact-difference form run
// -- all decided on the device, no host sync.
//
// Mapping: persistent CTAs (one per SM), NW = 16 warps (four per scheduler: a warp issues at most one DMMA per ~32 clk,
// the pipe takes one per 16, and a warp in its epilogue issues none).  theta for ALL components of the launch stays in
// shared memory for the whole kernel ([feature quad][component][4], 127 KB at K=32, D=30), so there is no ring and no
// barrier in the sample loop; mixtures whose theta does not fit are evaluated in component groups (MmaArgs below).
// A warp owns 8 NB samples per tile.  Its private slice holds them TRANSPOSED, [column][sample slot], the two sample
// blocks of a lane interleaved (slot 2g + nb): the x_i operand of a feature quad is then ONE 128-byte wavefront and one
// LDS.128 per lane serves both sample blocks (the row-major slice of round 1 took 2 LDS.64 and 4 wavefronts for it;
// bare loop 86.9 -> 90.6 % of the DMMA peak, profiles/r02b_dmma_feed.md).  The slice is filled by cp.async with the
// rows of the NEXT tile while the epilogue of the current one runs; x - c is applied in place.  The feature quads are
// taken in PAIRS: per pair one table load (LDS.64: byte offsets of both quads' two columns), 4 LDS.128 + 2 NB DMUL for
// phi, CB LDS.128 for theta ([pair][component][feature][quad]: both quads' fragments in one load), 2 NB x CB DMMAs
// into 8x8 (sample x component) accumulators; the y operands of the next pair are fetched before the DMMAs of this
// one and theta fragments are reloaded in place (an LDS instruction of either width costs the sub-partition about
// the same DMMA issue time -- scripts/ubench/k1_feed.cu, k2_feed.cu -- so fewer, wider loads: bare loop 89.8 -> 93.9 %).
// A sample's K log-pdfs end up inside one quad of lanes, so the log-sum-exp is two shuffles deep and the N x K
// outputs leave as 16-byte stores, two full sectors per quad.  The second pass (rho = exp(lp) w_k / (exp(log q) +
// tiny), or the VB soft-max with its sum r log r) happens here too, on the log-pdfs still in registers, one sample
// block at a time (no spills) -- k1_finish, the streaming pass behind the DFMA forms, returns at once.
// The exponentials of the log-sum-exp (one per pair: 5 % of the FP64 pipe's time with the library exp,
// profiles/r02b_k1_diag.md) use a 256-entry table of 2^(j/256) and a degree-4 polynomial -- 9 FP64 instructions
// instead of 17, <= 1 ulp; arguments below -707, where the result leaves the normal range, are redone with the
// library function so that the deep tails keep the reference's subnormal arithmetic.
#pragma once

#include "k1_exp_table.cuh"
#include "k1_fast_eval.cuh"

namespace pmc {

constexpr double kMmaMaxBias2 = 8.0e4;   // above this size of the cancelling terms (A_k >= 4 |b_k|^2) the DFMA forms run instead
constexpr int K1M_SCAL = 7;              // scalars per component kept in shared memory (S0..S4, weight, -tau), [scalar][KP]
constexpr int K1M_TAU = 6;               // row of -tau_k, the significance threshold of a term of the log-sum-exp (see below)
constexpr int K1M_EXPTAB = 256;          // entries of the 2^(j/256) table
constexpr int K1M_LOGTAB = 128 + 128 + 64;   // 1 / c_j, ln c_j (j < 128), e ln 2 (e < 64): log_tab()

struct MmaArgs {
  EvalArgs e;             // e.records = derived records ([T | -b | scalars]); only the scalars are read here
  const double* theta;    // [steps / 2][KP][4][2]  (k1m_theta_index)
  const double* shift;    // [d]
  const int* flag;        // flag[0] != 0: exact-difference form runs; else flag[1] != 0: this form runs; else k1_fast_eval
  int steps;              // feature quads = ceil(F / 4) rounded up to even (k1m_steps)
  int KP;                 // components (of this group) padded to 8 CB
  // Component groups: when theta of all components does not fit shared memory, the components are evaluated in
  // `ngroups` launches of at most KP each (e.records / e.cols / e.kl / theta describe THIS group).  The running
  // (max, weighted sum) of the log-sum-exp travels between the launches in rowstat; the last launch finishes log q
  // and leaves (max, 1 / denominator) there for k1_finish, which then does the second pass (SECOND == false).
  int group, ngroups;
  double* rowstat;        // [n, 2]; null when ngroups == 1 and no k1_finish pass follows
};

__host__ __device__ inline int k1m_features(int d) { return 1 + d + d * (d + 1) / 2; }
// slots per column of a warp's slice: 8 NB samples + 4 of padding.  The column stride is then 20 (NB = 2) or 12
// (NB = 1) doubles, == 4 and 12 (mod 16): the x_j loads of a quad -- 4 columns x 8 NB samples -- touch every bank once.
__host__ __device__ constexpr int k1m_col_stride(int NB) { return 8 * NB + 4; }
// feature quads, rounded up to an even count: the loop takes them in PAIRS (see the mapping note above); theta is stored
// [pair][component][feature in quad][quad of the pair], so that a lane's fragments of both quads are one 16-byte load
__host__ __device__ inline int k1m_steps(int d) { return ((k1m_features(d) + 3) / 4 + 1) & ~1; }
__host__ __device__ inline size_t k1m_theta_index(int f, int KP, int slot) {
  const int s = f >> 2;
  return (size_t(s >> 1) * KP + slot) * 8 + (f & 3) * 2 + (s & 1);
}
inline size_t k1m_smem_bytes(int d, int KP, int NB, int NW) {
  const int steps = k1m_steps(d), dp = (d + 1) & ~1;
  return sizeof(double) * (size_t(steps) * KP * 4 + size_t(KP) * K1M_SCAL + dp + K1M_EXPTAB + K1M_LOGTAB +
                           size_t(NW) * (d + 2) * k1m_col_stride(NB)) +
         sizeof(int) * size_t(steps) * 4 + 16;
}

// lower-triangle index t -> (row, col), row-major
__device__ __forceinline__ void tri_index(int t, int& r, int& c) {
  r = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  c = t - r * (r + 1) / 2;
}

// exp(x) for -707 <= x <= 0 (the caller diverts anything else): x = (256 m + j) ln2/256 + r, |r| <= ln2/512,
// exp(x) = 2^m T[j] e^r, e^r - 1 = r + r^2 (1/2 + r/6 + r^2/24) (truncation 4e-17).  9 FP64 instructions; the scaling
// by 2^m is an integer add on the exponent field (the result stays normal for x >= -707).  Measured against 60-digit
// arithmetic on 2e4 arguments: <= 2.2e-16 relative.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  const double magic = 6755399441055744.0;                         // 1.5 * 2^52: the low word of the sum is round(x 256/ln2)
  const double t = fma(x, 0x1.71547652b82fep+8, magic);            // 256 / ln 2
  const int k = __double2loint(t);
  const double kf = t - magic;
  double r = fma(kf, -0x1.62e42fef80000p-9, x);                    // ln2/256, upper 34 bits: kf * hi is exact
  r = fma(kf, -0x1.1cf79abc9e3b4p-44, r);                          // ln2/256, remainder
  const double r2 = r * r;
  double u = fma(r, 4.16666666666666644e-02, 1.66666666666666657e-01);
  u = fma(r, u, 0.5);
  const double p = fma(r2, u, r);
  const double tj = tab[k & (K1M_EXPTAB - 1)];
  const double v = fma(tj, p, tj);
  return __hiloint2double(__double2hiint(v) + (k >> 8) * 1048576, __double2loint(v));
}
// Order-preserving 32-bit key of a double's high word (signed compare of keys == compare of the doubles' upper 32 bits).
// The log-sum-exp needs a reference point m at or just below the largest log-pdf of the row -- any m gives
// m + log sum_k w_k exp(lp_k - m) -- so the maximum is taken over these keys with integer instructions (the FP64 pipe
// is the one every warp is waiting for) and m is the SMALLEST double with the winning high word: m <= max_k lp_k <=
// m + 2^-20 |m|, i.e. the largest term's exponent lies in [0, 2^-20 |m|] (0.0001 at |lp| = 100).
__device__ __forceinline__ int dkey(double v) {
  const int hi = __double2hiint(v);
  return hi ^ ((hi >> 31) & 0x7fffffff);
}
__device__ __forceinline__ double dkey_floor(int key) {
  return __hiloint2double(key ^ ((key >> 31) & 0x7fffffff), key >> 31);   // low word all ones below zero, zero above
}
// max(v, 0) on the integer pipe (like fmax: a NaN with the sign bit set also becomes 0)
__device__ __forceinline__ double clamp0(double v) { return (__double2hiint(v) < 0) ? 0.0 : v; }

// ln(x) for x in [1, 2^64) (the caller diverts anything else): x = 2^e m, m in [1, 2); c_j = 1 + (j + 1/2)/128 for the
// top seven mantissa bits j; r = m / c_j - 1 (|r| <= 2^-8, one fma with the tabulated reciprocal);
// ln x = e ln 2 + ln c_j + (r - r^2/2 + r^3/3 - r^4/4 + r^5/5 - r^6/6) (truncation 2e-18).  9 FP64 instructions where the
// library logarithm takes ~25; measured against 70-digit arithmetic on 3e4 arguments in [1, 1e18]: absolute error
// <= max(1e-17, 1.7e-16 |ln x|).  tab = [1 / c_j | ln c_j | e ln 2].
__device__ __forceinline__ double log_tab(double x, const double* __restrict__ tab) {
  const int hi = __double2hiint(x);
  const int j = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, tab[j], -1.0);
  double u = fma(r, -1.66666666666666657e-01, 2.00000000000000011e-01);
  u = fma(r, u, -0.25);
  u = fma(r, u, 3.33333333333333315e-01);
  u = fma(r, u, -0.5);
  const double p = fma(r * r, u, r);
  return (p + tab[128 + j]) + tab[256 + ((hi >> 20) - 1023)];
}
static __device__ __noinline__ double log_cold(double x) { return log(x); }
// ln(x) for any positive normal x (else the library function, out of line): log_tab's reduction with the exponent term
// formed by two FMAs (ln 2 = 44-bit head + tail, e * head exact), so arguments below 1 need no table entry -- the sum of
// the log-sum-exp lies in (0, K].  10 FP64 instructions in a dependent chain where the library logarithm has ~25;
// absolute error <= 2e-16 (1 + |ln x|) (tests/test_device_math_cpu.py emulates the same sequence against np.log).
__device__ __forceinline__ double log_any(double x, const double* __restrict__ tab) {
  const int hi = __double2hiint(x);
  if (unsigned(hi) - 0x00100000u >= 0x7fe00000u) return log_cold(x);   // zero, subnormal, negative, inf, nan
  const int j = (hi >> 13) & 127;
  const double e = double((hi >> 20) - 1023);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, tab[j], -1.0);
  double u = fma(r, -1.66666666666666657e-01, 2.00000000000000011e-01);
  u = fma(r, u, -0.25);
  u = fma(r, u, 3.33333333333333315e-01);
  u = fma(r, u, -0.5);
  const double p = fma(r * r, u, r);
  return fma(e, 0x1.62e42fefa38p-1, (p + tab[128 + j]) + e * 0x1.ef35793c7673p-45);
}
// ln(x) with RELATIVE accuracy next to 1 (VB: sum = 1 + eps for a sample with one responsible component, ln r of that
// component is -ln(sum) ~ -eps, and sum_n r ln r adds 1e5 such terms): x - 1 is exact for x in [1, 2), and below 2^-7 the
// series s - s^2/2 + ... + s^7/7 has a truncation of s^7/8 <= 2^-52 relative; everything else goes to log_any, whose
// absolute error of ~1e-16 is then <= 1e-14 of the result.
__device__ __forceinline__ double log_near1(double x, const double* __restrict__ tab) {
  const double s = x - 1.0;
  if (s >= 0.0 && s < 0x1p-7) {
    double u = fma(s, 1.42857142857142849e-01, -1.66666666666666657e-01);
    u = fma(s, u, 2.00000000000000011e-01);
    u = fma(s, u, -0.25);
    u = fma(s, u, 3.33333333333333315e-01);
    u = fma(s, u, -0.5);
    return fma(s * s, u, s);
  }
  return log_any(x, tab);
}
// 1 / x for normal positive x: hardware seed (rcp.approx.ftz.f64, ~2^-23) and two Newton steps, 4 DFMA
__device__ __forceinline__ double rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// the library exponential, out of line: it is only called in the cold paths (deep tails, the reference's literal rho
// formula, component groups), and 16-32 inlined copies of it per kernel made the binaries 130-210 KB -- beyond the
// instruction cache the 16 warps of a CTA, each in a different phase, have to share
static __device__ __noinline__ double exp_cold(double x) { return exp(x); }

constexpr unsigned kExpSlowHi = 0xC0861800u;                        // high word of -707.0: from there on the library exp runs

// ---------------------------------------------------------------------------------------------
// k1_mma_eval<CB, NB, NW, SECOND>: CB blocks of 8 components (KP = 8 CB), NB blocks of 8 samples per warp and tile,
// NW warps per CTA; SECOND: the launch also wants rho / r (the eval-only instantiation keeps no exponentials).
// DIAG (measurement builds only, PMCB200_K1_DIAG): 1 = exponentials replaced by one DFMA, 2 = no sample staging
// (the slice keeps its first tile), 4 = no epilogue beyond one store per sample -- to attribute the idle pipe time.
// ---------------------------------------------------------------------------------------------
template <int CB, int NB, int NW, bool SECOND, int DIAG = 0>
__global__ void __launch_bounds__(NW * 32, 1) k1_mma_eval(const MmaArgs ma) {
  static_assert(NB == 1 || NB == 2, "one LDS.64 / LDS.128 per operand serves the sample blocks of a lane");
  const EvalArgs& a = ma.e;
  if (ma.flag[0] != 0 || ma.flag[1] == 0) return;
  constexpr int RW = 8 * NB, TS = RW * NW, KP = 8 * CB, RS = k1m_col_stride(NB);
  constexpr int UNR = (CB * NB <= 4) ? 2 : 1;                          // pairs of feature quads per loop body (16 DMMAs or more)
  const int D = a.d, steps = ma.steps, dp = (D + 1) & ~1;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* theta_s = reinterpret_cast<double*>(smem_raw);               // [steps / 2][KP][4][2]
  double* scal_s = theta_s + size_t(steps) * KP * 4;                   // [K1M_SCAL][KP]
  double* cs = scal_s + KP * K1M_SCAL;                                 // [dp] shift
  double* etab = cs + dp;                                              // [256] 2^(j/256)
  double* ltab = etab + K1M_EXPTAB;                                    // [128 | 128 | 64] log_tab()
  double* y_all = ltab + K1M_LOGTAB;                                   // [NW][D + 2][RS]
  const int slice = (D + 2) * RS;
  int* tab = reinterpret_cast<int*>(y_all + size_t(NW) * slice);      // [steps / 2][4][2] byte offsets (column i | column j << 16)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;

  // ---- prologue: theta, scalars, shift, exp table, feature table, constant columns ----
  {
    const int n2 = steps * KP * 2;
    const double2* src = reinterpret_cast<const double2*>(ma.theta);
    double2* dst = reinterpret_cast<double2*>(theta_s);
    for (int i = tid; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
    const int rl = record_len(dp), so = tri_len(dp) + dp;
    // VB without the E_nk output (every E-step): variational.pyx:691 + :798 collapse to one FMA per pair,
    //   ln rho~_nk = [S0 + (S1 - S2 - S3) / 2] - (S4 / 2) q_nk,
    // the bracket in the row of S0 and S4 / 2 in the row of S1.  The reference's literal sequence (E_nk = S3 + S4 q first)
    // costs four FP64 instructions per pair, and in the epilogue every FP64 instruction of a warp waits for a turn behind
    // the other warps' DMMAs: with VB scalars the eval-only kernel measured 10.5 ms at C3 against 9.1 ms in Gauss mode
    // (9.0 ms with this form; the fused E-step 11.75 -> 11.30 ms).  The merged constant moves the rounding by ~1e-16 of
    // the terms' size, far inside the 1e-10 contract.
    const bool vb_fast = a.mode == MODE_VB && a.aux_out == nullptr;
    for (int i = tid; i < KP * (K1M_SCAL - 1); i += blockDim.x) {
      const int s = i / KP, k = i - s * KP;
      const double* sc = a.records + size_t(k < a.kl ? k : 0) * rl + so;
      double v = (k < a.kl) ? sc[s] : 0.0;
      if (vb_fast && k < a.kl) {
        if (s == S0) v = sc[S0] + 0.5 * ((sc[S1] - sc[S2]) - sc[S3]);
        if (s == S1) v = 0.5 * sc[S4];
      }
      scal_s[i] = v;
    }
    // Significance thresholds.  A term w_k exp(lp_k - m) below 2^-66 w_min (w_min = the smallest weight) is below
    // 2^-66 of the sum -- the component that attains the maximum contributes at least w_min (m <= max lp) -- so dropping
    // every such term changes the sum by less than K 2^-66 relative, 1e-18 at K = 64.  With -tau_k = ln(w_min / w_k)
    // - 45.75 a term is kept iff lp_k - m > -tau_k.  For samples that belong to one or two components of a
    // separated mixture (the usual case in D >= 10) this leaves one or two exponentials per sample instead of K.
    // The shortcut needs every weight positive and the maximum taken over the evaluated components only: with zero
    // weights, max_init = 0 (pmc.pyx:26-27) or component groups, -tau_k = -inf and every finite term is kept.
    {
      double wmin = INFINITY;
      bool ok = (a.max_init == -DBL_MAX) && ma.ngroups == 1;
      for (int k = 0; k < a.kl; ++k) {
        const double w = a.records[size_t(k) * rl + so + S_WEIGHT];
        ok = ok && (w > 0.0) && isfinite(w);
        wmin = fmin(wmin, w);
      }
      for (int k = tid; k < KP; k += blockDim.x) {
        const double w = (k < a.kl) ? a.records[size_t(k) * rl + so + S_WEIGHT] : 0.0;
        scal_s[K1M_TAU * KP + k] = (ok && k < a.kl) ? log(wmin / w) - 45.75 : -INFINITY;
      }
    }
    for (int j = tid; j < dp; j += blockDim.x) cs[j] = (j < D) ? ma.shift[j] : 0.0;
    for (int j = tid; j < K1M_EXPTAB; j += blockDim.x) etab[j] = kExp2Table[j];
    for (int j = tid; j < 128; j += blockDim.x) { ltab[j] = kLogInvC[j]; ltab[128 + j] = kLogC[j]; }
    for (int j = tid; j < 64; j += blockDim.x) ltab[256 + j] = kLogE[j];
    const int F = k1m_features(D);
    for (int f = tid; f < steps * 4; f += blockDim.x) {
      int oi = D + 1, oj = D + 1;                                       // zero column
      if (f == 0) { oi = D; oj = D; }
      else if (f <= D) { oi = f - 1; oj = D; }
      else if (f < F) tri_index(f - 1 - D, oi, oj);
      tab[(((f >> 3) * 4) + (f & 3)) * 2 + ((f >> 2) & 1)] = (oi * RS * 8) | ((oj * RS * 8) << 16);
    }
    for (int i = tid; i < NW * slice; i += blockDim.x) y_all[i] = ((i % slice) / RS == D) ? 1.0 : 0.0;
  }
  // contiguous output columns (the usual case) allow 16-byte stores
  const bool staged = a.lp_out != nullptr || a.resp_out != nullptr || a.aux_out != nullptr;
  int contig_l = 1;
  if (staged) {
    const int c0 = __ldg(a.cols);
    for (int k = tid; k < a.kl; k += blockDim.x) contig_l &= (__ldg(a.cols + k) == c0 + k);
    contig_l &= ((c0 & 1) == 0) && ((a.k_out & 1) == 0) && ((a.kl & 1) == 0);
    contig_l &= (a.lp_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.lp_out) & 15) == 0);
    contig_l &= (a.resp_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.resp_out) & 15) == 0);
    contig_l &= (a.aux_out == nullptr) || ((reinterpret_cast<uintptr_t>(a.aux_out) & 15) == 0);
  }
  const bool contig = __syncthreads_and(contig_l) != 0;               // also publishes the prologue's stores
  const int col0 = staged ? __ldg(a.cols) : 0;

  double* yw = y_all + size_t(warp) * slice;
  const char* ylane = reinterpret_cast<const char*>(yw + g * NB);      // this lane's sample slot(s) in column 0
  const double2* thl = reinterpret_cast<const double2*>(theta_s) + g * 4 + tq;   // theta[pair][8 cb + g][tq] = (quad 2p, quad 2p + 1)
  const int2* tabl = reinterpret_cast<const int2*>(tab) + tq;
  const double* scl = scal_s + 2 * tq;                                 // scalars of components 8 cb + 2 tq + {0, 1}

  const int64_t num_tiles = (a.n + TS - 1) / TS;
  double part_a = 0.0, part_w = 0.0;
  unsigned wmask = 0u;                                                 // bit 2 cb + e: component 8 cb + 2 tq + e has a nonzero weight
#pragma unroll
  for (int cb = 0; cb < CB; ++cb) {
    const double2 w2 = *reinterpret_cast<const double2*>(scl + S_WEIGHT * KP + 8 * cb);
    wmask |= (w2.x != 0.0 ? 1u : 0u) << (2 * cb);
    wmask |= (w2.y != 0.0 ? 2u : 0u) << (2 * cb);
  }

  // this warp's rows of a tile -> its slice, asynchronously (LDGSTS, lane = column; sample r -> slot NB (r % 8) + r / 8);
  // rows beyond n are zero-filled.  The shift is applied in place when the tile is picked up.
  auto stage_rows = [&](int64_t r0) {
    const bool full = r0 + RW <= a.n;
    for (int j = lane; j < D; j += 32) {
      const uint32_t dst = smem_u32(yw + j * RS);
      const double* src = a.x + r0 * a.ldx + j;
      if (full) {
#pragma unroll
        for (int r = 0; r < RW; ++r)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + uint32_t(NB * (r & 7) + (r >> 3)) * 8u),
                       "l"(src + r * a.ldx)
                       : "memory");
      } else {
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const bool in = r0 + r < a.n;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + uint32_t(NB * (r & 7) + (r >> 3)) * 8u),
                       "l"(in ? src + r * a.ldx : a.x), "r"(in ? 8 : 0)
                       : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (int64_t(blockIdx.x) < num_tiles) stage_rows(int64_t(blockIdx.x) * TS + int64_t(warp) * RW);
  // The warps of a scheduler would run in lock step: all in the DMMA loop (sharing the pipe), then all in the
  // latency-bound epilogue (pipe idle, 18 % of the time in the first profile).  Starting warp group j that many
  // DMMA-loops later keeps them in different phases for the whole kernel: a warp's epilogue hides behind the others' loops.
  if (warp >= 4 && num_tiles >= 4 * int64_t(gridDim.x)) {
    const long long t0 = clock64(), wait = (long long)steps * (NB * CB * 16) * (warp / 4);
    while (clock64() - t0 < wait) {
    }
  }

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TS + int64_t(warp) * RW;
    // ---- pick up the staged rows: y = x - c in place, 16 bytes (two sample slots of one column) per lane and step ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (!(DIAG & 2) || tile == blockIdx.x) {
      constexpr int PPC = RW / 2;                                       // slot pairs per column
      for (int e2 = lane; e2 < D * PPC; e2 += 32) {
        const int col = e2 / PPC, pp = e2 - col * PPC;
        double2* p = reinterpret_cast<double2*>(yw + col * RS + 2 * pp);
        const double c = cs[col];
        double2 v = *p;
        v.x -= c;
        v.y -= c;
        *p = v;
      }
    }
    __syncwarp();

    double acc[NB][CB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) { acc[nb][cb][0] = 0.0; acc[nb][cb][1] = 0.0; }

    // ---- q = phi . theta over the feature quads, two quads per pass (round 2, scripts/ubench/k1_feed.cu: an LDS
    // instruction of either width costs the sub-partition's DMMA issue about the same, so theta travels as ONE LDS.128 per
    // component block and pair of quads, the table word of both quads as one LDS.64).  The y operands of the next pair
    // are fetched before the DMMAs of this one; a theta fragment is reloaded in place right after its last DMMA of the
    // pair -- component blocks go in groups of G, quad 0 then quad 1 of the group, so the two DMMAs on one accumulator
    // stay G NB instructions apart and no second fragment buffer is needed (bare loop at C2 89.8 -> 93.9 % of the pipe).
    // Every accumulator still sees the quads in ascending order: same bits as the quad-at-a-time loop.
    const int pairs = steps >> 1;
    double yi_n[2][NB], yj_n[2][NB];
    double2 th[CB];
    auto ld_y = [&](unsigned off, double (&v)[NB]) {
      if constexpr (NB == 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(ylane + off);
        v[0] = t2.x; v[NB - 1] = t2.y;
      } else {
        v[0] = *reinterpret_cast<const double*>(ylane + off);
      }
    };
    auto fetch_y = [&](int p) {
      const int2 t = tabl[4 * p];
      ld_y(unsigned(t.x) & 0xffffu, yi_n[0]); ld_y(unsigned(t.x) >> 16, yj_n[0]);
      ld_y(unsigned(t.y) & 0xffffu, yi_n[1]); ld_y(unsigned(t.y) >> 16, yj_n[1]);
    };
    fetch_y(0);
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) th[cb] = thl[cb * 8 * 4];
    const uint32_t th_addr = smem_u32(thl);
    constexpr int G = 2;                                                // (the last group of an odd CB holds one block)
#pragma unroll UNR
    for (int p = 0; p < pairs; ++p) {
      double ph[2][NB];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) ph[q][nb] = yi_n[q][nb] * yj_n[q][nb];
      const int pn = min(p + 1, pairs - 1);
      fetch_y(pn);
#pragma unroll
      for (int c0 = 0; c0 < CB; c0 += G) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int cb = c0; cb < c0 + G && cb < CB; ++cb)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(acc[nb][cb][0]), "+d"(acc[nb][cb][1])
                           : "d"(ph[q][nb]), "d"(q ? th[cb].y : th[cb].x));
        // (volatile: the compiler otherwise sinks these loads to the end of the loop body, in front of their first use)
#pragma unroll
        for (int cb = c0; cb < c0 + G && cb < CB; ++cb)
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                       : "=d"(th[cb].x), "=d"(th[cb].y)
                       : "r"(th_addr + uint32_t(pn * KP * 64 + cb * 512)));
      }
    }
    // ---- the slice is free: start copying this warp's rows of the CTA's next tile behind the epilogue ----
    __syncwarp();
    if (!(DIAG & 2) && tile + gridDim.x < num_tiles) stage_rows(row0 + int64_t(gridDim.x) * TS);
    if constexpr ((DIAG & 4) != 0) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        double t = 0.0;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) t += acc[nb][cb][0] + acc[nb][cb][1];
        const int64_t row = row0 + 8 * nb + g;
        if (row < a.n && tq == 0 && a.logq) a.logq[row] = t;
      }
      continue;
    }

    // ---- epilogue: lane holds q[sample 8 nb + g][component 8 cb + 2 tq + e] ----
    // Phase 1, whole tile: log-pdfs (they replace q in acc), gamma / E, running maxima, stores of the log-pdfs.
    double mx[NB], m_prev[NB], s_prev[NB];
    int mkey[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      m_prev[nb] = a.max_init;
      s_prev[nb] = 0.0;
      if (ma.group > 0) {                                               // running (reference point, sum) of the earlier groups
        const int64_t row = row0 + 8 * nb + g;
        if (row < a.n) {
          m_prev[nb] = ma.rowstat[2 * row];
          s_prev[nb] = ma.rowstat[2 * row + 1];
        }
      }
      mkey[nb] = dkey(m_prev[nb]);
    }
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      const int k = 8 * cb + 2 * tq;
      const bool pad0 = k >= a.kl, pad1 = k + 1 >= a.kl;
      const double2 c0 = *reinterpret_cast<const double2*>(scl + S0 * KP + 8 * cb);
      if (a.mode == MODE_GAUSS) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const double l0 = c0.x - 0.5 * clamp0(acc[nb][cb][0]);                         // gauss.pyx:151
          const double l1 = c0.y - 0.5 * clamp0(acc[nb][cb][1]);
          acc[nb][cb][0] = pad0 ? -INFINITY : l0;
          acc[nb][cb][1] = pad1 ? -INFINITY : l1;
        }
      } else if (a.mode == MODE_VB && a.aux_out == nullptr) {
        const double2 h2 = *reinterpret_cast<const double2*>(scl + S1 * KP + 8 * cb);   // S4 / 2 (see the prologue)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const double l0 = fma(-h2.x, clamp0(acc[nb][cb][0]), c0.x);                    // variational.pyx:691, :798
          const double l1 = fma(-h2.y, clamp0(acc[nb][cb][1]), c0.y);
          acc[nb][cb][0] = pad0 ? -INFINITY : l0;
          acc[nb][cb][1] = pad1 ? -INFINITY : l1;
        }
      } else {
        const double2 c1 = *reinterpret_cast<const double2*>(scl + S1 * KP + 8 * cb);
        const double2 c2 = *reinterpret_cast<const double2*>(scl + S2 * KP + 8 * cb);
        const double2 c3 = *reinterpret_cast<const double2*>(scl + S3 * KP + 8 * cb);
        const double2 c4 = *reinterpret_cast<const double2*>(scl + S4 * KP + 8 * cb);
        const bool student = a.mode == MODE_STUDENT_T, want_aux = a.aux_out != nullptr;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          double l[2], ax[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double c0e = e ? c0.y : c0.x, c1e = e ? c1.y : c1.x, c2e = e ? c2.y : c2.x, c3e = e ? c3.y : c3.x,
                         c4e = e ? c4.y : c4.x;
            const double q = clamp0(acc[nb][cb][e]);
            if (student) {
              double t = fma(q, c2e, 1.0);                                               // student_t.pyx:159-164
              // 1 <= t < 2^64 always, unless q is non-finite or absurd: then the library logarithm
              t = (unsigned(__double2hiint(t)) - 0x3ff00000u < 0x04000000u) ? log_tab(t, ltab) : log_cold(t);
              l[e] = fma(t, c1e, c0e);
              ax[e] = want_aux ? c4e * rcp_pos(c3e + q) : 0.0;                           // gamma_nk, pmc.pyx:610
            } else {
              ax[e] = c3e + c4e * q;                                                     // variational.pyx:798
              l[e] = c0e + 0.5 * (c1e - c2e - ax[e]);                                    // variational.pyx:691
            }
          }
          acc[nb][cb][0] = pad0 ? -INFINITY : l[0];
          acc[nb][cb][1] = pad1 ? -INFINITY : l[1];
          const int64_t row = row0 + 8 * nb + g;
          if (a.aux_out && row < a.n) {
            if (contig) {
              if (!pad0) *reinterpret_cast<double2*>(a.aux_out + size_t(row) * a.k_out + col0 + k) = make_double2(ax[0], ax[1]);
            } else {
              if (!pad0) a.aux_out[size_t(row) * a.k_out + __ldg(a.cols + k)] = ax[0];
              if (!pad1) a.aux_out[size_t(row) * a.k_out + __ldg(a.cols + k + 1)] = ax[1];
            }
          }
        }
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) mkey[nb] = max(mkey[nb], max(dkey(acc[nb][cb][0]), dkey(acc[nb][cb][1])));
    }
    // N x K values leave as 16-byte stores when the output columns are contiguous
    auto store_pairs = [&](double* out, int nb, const double (&v)[CB][2]) {
      const int64_t row = row0 + 8 * nb + g;
      if (row >= a.n) return;
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) {
        const int k = 8 * cb + 2 * tq;
        if (contig) {
          if (k < a.kl) *reinterpret_cast<double2*>(out + size_t(row) * a.k_out + col0 + k) = make_double2(v[cb][0], v[cb][1]);
        } else {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (k + e < a.kl) out[size_t(row) * a.k_out + __ldg(a.cols + k + e)] = v[cb][e];
        }
      }
    };
    {
      // mixtures: `individual`; without a fused second pass the raw log-pdfs also wait in resp_out for k1_finish
      double* const lp_dst = SECOND ? ((a.mode != MODE_VB) ? a.lp_out : nullptr) : (a.lp_out ? a.lp_out : a.resp_out);
      if (lp_dst) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) store_pairs(lp_dst, nb, acc[nb]);
      }
    }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      mkey[nb] = max(mkey[nb], __shfl_xor_sync(0xffffffffu, mkey[nb], 1));
      mkey[nb] = max(mkey[nb], __shfl_xor_sync(0xffffffffu, mkey[nb], 2));
      mx[nb] = dkey_floor(mkey[nb]);                                    // reference point of the log-sum-exp (see dkey)
      if ((mkey[nb] ^ (mkey[nb] >> 31)) >= 0x41B00000) {                 // |m| >= 2^28: m could sit hundreds below the maximum
        double e = (ma.group > 0) ? m_prev[nb] : mx[nb];                // and exp(lp - m) overflow -- take the exact maximum
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) e = fmax(e, fmax(acc[nb][cb][0], acc[nb][cb][1]));
        const unsigned quad = 0xFu << (lane & ~3);                      // (the four lanes of a row take this branch together)
        e = fmax(e, __shfl_xor_sync(quad, e, 1));
        e = fmax(e, __shfl_xor_sync(quad, e, 2));
        mx[nb] = e;
      }
    }

    // Phase 2: the terms w_k exp(lp - max) of a sample block and their sum over the quad (_regularize.pyx:72-81 up to
    // rounding).  Eval-only: both blocks back to back (independent exponentials); with the fused second pass the
    // terms of one block are kept for rho / r while the other block waits in acc.
    auto term = [&](double d, double w, bool live, bool& deep) -> double {
      if constexpr ((DIAG & 1) != 0) return w * fma(d, 1e-3, 1.0);
      const bool slow = unsigned(__double2hiint(d)) >= kExpSlowHi;     // d <= -707 (padding: -inf, weight 0)
      deep |= slow && live;
      return __dmul_rn(w, exp_tab(slow ? 0.0 : d, etab));               // (no contraction: the same bits in every instantiation)
    };
    auto quad_sum = [&](int nb, double sum) -> double {
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (ma.group > 0) sum = fma(s_prev[nb], exp_cold(m_prev[nb] - mx[nb]), sum);
      return sum;
    };
    // bit 2 cb + e: the term of component 8 cb + 2 tq + e is significant (d > -tau_k, or a NaN that must propagate)
    auto significant = [&](int nb, double (&d)[CB][2]) -> unsigned {
      unsigned sig = 0u;
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) {
        const double2 t2 = *reinterpret_cast<const double2*>(scl + K1M_TAU * KP + 8 * cb);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          d[cb][e] = acc[nb][cb][e] - mx[nb];
          const unsigned hi = unsigned(__double2hiint(d[cb][e])), thr = unsigned(__double2hiint(e ? t2.y : t2.x));
          sig |= ((hi < thr || hi > 0xFFF00000u) ? 1u : 0u) << (2 * cb + e);
        }
      }
      return sig;
    };
    // eval-only: nothing is kept.  Rounds over the significant terms of the warp: in every round each lane evaluates
    // its next significant term (none: a zero), so a tile costs max-over-lanes(#significant) exponentials per lane
    // instead of 2 CB.  The terms are added in component order, like the dense sum of the second-pass instantiation
    // with the insignificant ones left out -- the two give the same bits.  A lane that met the deep tail (d <= -707:
    // subnormal results and zeros, as the library exp gives them; only possible with tau = inf) redoes its sum.
    auto terms_sum = [&](int nb) -> double {
      double d[CB][2];
      unsigned sig = significant(nb, d);
      const unsigned sig0 = sig;
      bool deep = false;
      double sum = 0.0;
      if (__any_sync(0xffffffffu, __popc(sig) > (CB + 1) / 2)) {        // many overlapping components: all terms, straight-line
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const double2 w2 = *reinterpret_cast<const double2*>(scl + S_WEIGHT * KP + 8 * cb);
          const double t0 = term(d[cb][0], w2.x, (wmask >> (2 * cb)) & 1u, deep);
          const double t1 = term(d[cb][1], w2.y, (wmask >> (2 * cb + 1)) & 1u, deep);
          if (sig & (1u << (2 * cb))) sum += t0;
          if (sig & (2u << (2 * cb))) sum += t1;
        }
        sig = 0u;
      }
      while (__any_sync(0xffffffffu, sig != 0u)) {
        const int idx = __ffs(int(sig)) - 1;                            // -1: this lane has nothing left
        double dv = 0.0;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          if (idx == 2 * cb) dv = d[cb][0];
          if (idx == 2 * cb + 1) dv = d[cb][1];
        }
        const double wv = (idx >= 0) ? scl[S_WEIGHT * KP + 8 * (idx >> 1) + (idx & 1)] : 0.0;
        sum += term(dv, wv, idx >= 0 && ((wmask >> idx) & 1u), deep);
        sig &= sig - 1u;
      }
      if (deep) {
        sum = 0.0;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const double2 w2 = *reinterpret_cast<const double2*>(scl + S_WEIGHT * KP + 8 * cb);
          if (sig0 & (1u << (2 * cb))) sum += __dmul_rn(w2.x, exp_cold(d[cb][0]));
          if (sig0 & (2u << (2 * cb))) sum += __dmul_rn(w2.y, exp_cold(d[cb][1]));
        }
      }
      return quad_sum(nb, sum);
    };
    // fused second pass: every term is needed for rho / r; the sum runs over the significant ones (see terms_sum)
    auto terms_keep = [&](int nb, double (&ex)[CB][2]) -> double {
      double d[CB][2];
      const unsigned sig = significant(nb, d);
      bool deep = false;
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) {
        const double2 w2 = *reinterpret_cast<const double2*>(scl + S_WEIGHT * KP + 8 * cb);
        ex[cb][0] = term(d[cb][0], w2.x, (wmask >> (2 * cb)) & 1u, deep);
        ex[cb][1] = term(d[cb][1], w2.y, (wmask >> (2 * cb + 1)) & 1u, deep);
      }
      if (deep) {
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const double2 w2 = *reinterpret_cast<const double2*>(scl + S_WEIGHT * KP + 8 * cb);
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (unsigned(__double2hiint(d[cb][e])) >= kExpSlowHi) ex[cb][e] = __dmul_rn(e ? w2.y : w2.x, exp_cold(d[cb][e]));
        }
      }
      double sum = 0.0;
#pragma unroll
      for (int cb = 0; cb < CB; ++cb) {
        if (sig & (1u << (2 * cb))) sum += ex[cb][0];
        if (sig & (2u << (2 * cb))) sum += ex[cb][1];
      }
      return quad_sum(nb, sum);
    };
    // per-sample results for one row: log q and the sums; returns log q, f0 = 1 / sum and f1 = -log(sum) for the second pass
    auto finish_row = [&](int64_t row, bool writer, double sum, double m, double& lq, double& f0, double& f1) {
      lq = 0.0;
      f0 = 0.0;
      f1 = 0.0;
      if (!SECOND && ma.group + 1 < ma.ngroups) {                       // more components to come
        if (writer) {
          ma.rowstat[2 * row] = m;
          ma.rowstat[2 * row + 1] = sum;
        }
        return;
      }
      // ln(sum): the table logarithm in the mixture modes (the same function in every instantiation: same bits of log q).
      // VB: sum = 1 + eps for a sample with one responsible component, ln r of that component IS -ln(sum) ~ -eps, and
      // sum_n r ln r adds 1e5 such terms -- the table form's absolute error of ~1e-16 is systematic near 1 (one table
      // entry) and showed up as 2e-12 in the bound's q_Z term (contract 1e-10 relative): log_near1 has a series there.
      const double ls = (a.mode == MODE_VB) ? log_near1(sum, ltab) : log_any(sum, ltab);
      lq = ls + m;                                                      // _regularize.pyx:81
      if (writer) {
        const double w_n = a.sw ? __ldg(a.sw + row) : 1.0;
        part_w += w_n;
        if (a.logq) a.logq[row] = lq;
        if (a.mode != MODE_VB) part_a += w_n * lq;                      // pmc.pyx:388-391
      }
      if constexpr (SECOND) {
        // 1 / sum: hardware seed + two Newton steps for a positive normal sum (5 instructions, <= 1 ulp), else the division
        f0 = (unsigned(__double2hiint(sum)) - 0x00100000u < 0x7fe00000u) ? rcp_pos(sum) : 1.0 / sum;   // variational.pyx:728-755 (and rho below)
        f1 = -ls;
      } else if (ma.rowstat && writer) {                                // for k1_finish
        ma.rowstat[2 * row] = m;
        ma.rowstat[2 * row + 1] = (a.mode != MODE_VB) ? 1.0 / (exp_cold(lq) + kTiny) : 1.0 / sum;   // pmc.pyx:39-41
      }
    };
    if constexpr (!SECOND) {
      double sum[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) sum[nb] = terms_sum(nb);
      // one logarithm per lane: lanes with odd tq finish the second sample block, lanes tq = 0 / 1 write
      const int sel = (NB == 2) ? (tq & 1) : 0;
      const int64_t row = row0 + 8 * sel + g;
      double lq, f0, f1;
      finish_row(row, tq < NB && row < a.n, sel ? sum[NB - 1] : sum[0], sel ? mx[NB - 1] : mx[0], lq, f0, f1);
    } else {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        double ex[CB][2], lq, f0, f1;
        const double sum = terms_keep(nb, ex);
        finish_row(row0 + 8 * nb + g, tq == nb && row0 + 8 * nb + g < a.n, sum, mx[nb], lq, f0, f1);
        if (a.mode != MODE_VB) {
          // rho_nk = exp(lp_nk) w_k / (exp(log q_n) + tiny)  (pmc.pyx:39-41)  = [w_k exp(lp_nk - m)] / sum_j [w_j exp(lp_nj - m)]
          // as long as neither exponential of the reference's formula leaves the normal range and tiny is negligible
          // beside exp(log q); where exp(lp_nk) is subnormal (lp < -700), exp(log q) overflows or log q < -650 (tiny /
          // exp(log q) > 1e-26), the literal formula reproduces the reference's own rounding
          const bool literal_row = mx[nb] > 700.0 || lq < -650.0;
          bool literal = literal_row;
#pragma unroll
          for (int cb = 0; cb < CB; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              literal |= unsigned(__double2hiint(acc[nb][cb][e])) >= 0xC085E000u;      // lp <= -700 (padding: -inf)
              ex[cb][e] *= f0;
            }
          if (literal) {
            const double g0 = 1.0 / (exp_cold(lq) + kTiny);
#pragma unroll
            for (int cb = 0; cb < CB; ++cb)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                if (acc[nb][cb][e] < -700.0 || literal_row)
                  ex[cb][e] = exp_cold(acc[nb][cb][e]) * scl[S_WEIGHT * KP + 8 * cb + e] * g0;
          }
          store_pairs(a.resp_out, nb, ex);
        } else {
          const int64_t row = row0 + 8 * nb + g;
          const double w_r = (a.sw && row < a.n) ? __ldg(a.sw + row) : 1.0;
          double s_rl = 0.0;
#pragma unroll
          for (int cb = 0; cb < CB; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double rv = ex[cb][e] * f0;
              if (((__double2hiint(rv) & 0x7fffffff) | __double2loint(rv)) == 0) rv = kTiny;   // r == 0 -> tiny, variational.pyx:753-754
              const double lrn = (acc[nb][cb][e] - mx[nb]) + f1;                      // variational.pyx:741,755
              ex[cb][e] = rv;
              acc[nb][cb][e] = lrn;
              if (8 * cb + 2 * tq + e < a.kl) s_rl = fma(rv, lrn, s_rl);              // variational.pyx:1003-1013
            }
          if (row < a.n) part_a = fma(w_r, s_rl, part_a);
          if (a.resp_out) store_pairs(a.resp_out, nb, ex);
          if (a.lp_out) store_pairs(a.lp_out, nb, acc[nb]);
        }
      }
    }
  }

  if (a.partials) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      part_a += __shfl_xor_sync(0xffffffffu, part_a, o);
      part_w += __shfl_xor_sync(0xffffffffu, part_w, o);
    }
    if (lane == 0) {
      a.partials[(size_t(blockIdx.x) * PMC_MAX_WARPS + warp) * 2 + 0] = part_a;
      a.partials[(size_t(blockIdx.x) * PMC_MAX_WARPS + warp) * 2 + 1] = part_w;
    }
  }
}

}  // namespace pmc
