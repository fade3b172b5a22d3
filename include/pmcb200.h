/*
 * pmcb200.h -- C ABI of the B200-native (sm_100a) mixture-density / proposal-update hot path.
 *
 * The reference (pypmc v1.2.6 @ 9e0ab49) has no C/FFI boundary for this path: the loops live in Cython
 * modules behind Python classes.  Each entry point below names the reference loops it replaces
 * (file:line relative to the reference tree); pypmc_b200/ (Python, ctypes) is the host-side mirror of the
 * reference classes on top of this ABI, and INTEGRATION.md shows the binding a pypmc maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pmcb200_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - `*_dev` pointers are device pointers on the context's device, `*_host` pointers are host memory
 *     (pinned memory makes the copies asynchronous and fast; pageable memory works).
 *   - a context is bound to one device and is not thread-safe: one context per host thread (the Python mirror keeps
 *     one per device and holds the GIL across calls); several contexts / processes may share a GPU.
 *   - all matrices are float64, row-major; `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - N x K outputs have row stride k_out.
 */
#ifndef PMCB200_H
#define PMCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMCB200_VERSION 101

#define PMCB200_MODE_GAUSS 0     /* Gauss components      density/gauss.pyx:132-153      */
#define PMCB200_MODE_STUDENT_T 1 /* StudentT components   density/student_t.pyx:135-166  */
#define PMCB200_MODE_VB 2        /* GaussianInference E-step   mix_adapt/variational.pyx:116-127 */

#define PMCB200_NUM_SCALARS 8
#define PMCB200_MAX_DIM 64

typedef struct pmcb200_ctx pmcb200_ctx;

int pmcb200_version(void);
const char* pmcb200_last_error(void);

/* number of CUDA devices visible (0 without a GPU; never fails) */
int pmcb200_device_count(void);

/* Create / destroy a context bound to one CUDA device (owns scratch buffers and two copy streams). */
int pmcb200_create(int device, pmcb200_ctx** out);
int pmcb200_destroy(pmcb200_ctx* ctx);

/* ---- component records (host side, no GPU needed) ------------------------------------------------
 * One packed record per component: the lower-triangular factor T (T^T T = Sigma^-1, or = W_k for VB), the
 * centre and 8 mode-dependent scalars.  Layout: pypmc_b200/csrc/pmc_common.cuh.  These replace the
 * per-component state the reference keeps in LocalGauss / StudentT objects (density/gauss.pyx:23-56,
 * density/student_t.pyx:78-117) and GaussianInference's (m, W, nu, beta) (mix_adapt/variational.pyx:774-798).
 */
int pmcb200_record_len(int d);             /* doubles per record for dimension d, -1 if d unsupported */
int pmcb200_pack_record(int d,
                        const double* t_lower_host, /* [d, d] row-major, upper part ignored */
                        const double* center_host,  /* [d] */
                        const double* scalars_host, /* [PMCB200_NUM_SCALARS] */
                        double* record_host);       /* [pmcb200_record_len(d)] */

/* ---- K1: fused log-pdf + mixture log-sum-exp + responsibilities ------------------------------------
 * Replaces, in one pass over the samples:
 *   MixtureDensity.multi_evaluate   density/mixture.pyx:112-156   (+ Gauss/StudentT.multi_evaluate,
 *                                   tools/_linalg.pyx:10-39 bilinear_sym, tools/_regularize.pyx:57-83 logsumexp2D)
 *   calculate_rho_rb                mix_adapt/pmc.pyx:23-43
 *   gamma_nk of student_t_pmc       mix_adapt/pmc.pyx:602-610
 *   E-step of GaussianInference     mix_adapt/variational.pyx:774-798, 675-691, 711-757
 *   PMC.log_likelihood reduction    mix_adapt/pmc.pyx:371-391;  E[log q(Z)]  mix_adapt/variational.pyx:1003-1013
 *
 * records_dev/cols_dev describe the `kl` components to evaluate (record i writes column cols[i]).
 * Outputs (any may be NULL):
 *   logq_dev [n]            log q(x_n) = LSE_k(lp_nk + ln w_k)        (VB: LSE of the unnormalised log rho)
 *   lp_dev   [n, k_out]     component log-pdfs `individual` (mixture) / normalised log_rho (VB)
 *   resp_dev [n, k_out]     rho_nk (PMC, pmc.pyx:39-41) / r_nk (VB, zeros replaced by tiny)
 *   aux_dev  [n, k_out]     gamma_nk (Student-t) / expectation_gauss_exponent (VB)
 *   sums_dev [2]            { sum_n w_n log q_n  (VB: sum_n w_n sum_k r_nk log r_nk),  sum_n w_n }
 * weights_dev [n] are the per-sample weights used in sums_dev only (NULL = 1).
 * max_init is the starting value of the running maximum: -DBL_MAX normally, 0.0 to reproduce the reference
 * when dead columns (holding 0) take part in logsumexp2D's maximum (pmc.pyx:26-27, _regularize.pyx:72-76).
 *
 * Three kernel forms sit behind this call; which one works is decided on the device from the component parameters
 * alone (never from the set of outputs, so log q has the same bits whichever outputs are requested): the FP64
 * matrix-instruction form (kl >= 9, d >= 8, max_k |T_k (mu_k - c)|^2 <= 2e4; in groups of components when their
 * parameters exceed shared memory), the DFMA form
 * q = |T x' - b|^2 (|b| <= 1e4), and the exact-difference form y = x - mu_k for anything further out.
 * Environment (tuning / comparison runs only, read per call): PMCB200_K1_FORM=dfma disables the first form.
 */
int pmcb200_mixture_eval(pmcb200_ctx* ctx,
                         const double* x_dev, int64_t n, int64_t ldx, int d,
                         const double* records_dev, const int* cols_dev, int kl,
                         int k_out, int mode, double max_init,
                         double* logq_dev, double* lp_dev, double* resp_dev, double* aux_dev,
                         const double* weights_dev, double* sums_dev,
                         void* stream);

/* ---- K2: weighted sufficient statistics -----------------------------------------------------------
 * Replaces the einsum / triple loops of
 *   gaussian_pmc   mix_adapt/pmc.pyx:191-222      student_t_pmc   mix_adapt/pmc.pyx:612-650
 *   GaussianInference._update_N_comp/_update_x_mean_comp/_update_S[_weighted]  mix_adapt/variational.pyx:699-709, 806-932
 * out_dev [k, 3 + d + d(d+1)/2] per component:
 *   [0] A = sum w rho   [1] B = sum w rho gamma   [2..2+d) m = sum w rho gamma (x - shift)
 *   [2+d..2+d+d(d+1)/2)  R = sum w rho gamma (x-shift)(x-shift)^T, lower triangle row-major (i(i+1)/2 + j)
 *   [last] sum w rho ln(gamma)  (0 without gamma; the N-sized part of the dof condition, pmc.pyx:654-691).
 * gamma_dev and weights_dev may be NULL (= 1).  These K rows are what ranks all-reduce (sum).
 */
int pmcb200_suffstats(pmcb200_ctx* ctx,
                      const double* x_dev, int64_t n, int64_t ldx, int d,
                      const double* shift_dev,
                      const double* rho_dev, const double* gamma_dev, int k, int ld_rho,
                      const double* weights_dev,
                      double* out_dev,
                      void* stream);

/* ---- host-buffer (end-to-end) form of K1 ------------------------------------------------------------
 * Same computation as pmcb200_mixture_eval with HOST buffers: samples are streamed to the device in
 * chunks of `chunk_rows` rows (0 = library default) on two streams so copies overlap the kernel; results are
 * copied back per chunk.  This is the call behind MixtureDensity.multi_evaluate(ndarray) in pypmc_b200.
 */
int pmcb200_mixture_eval_host(pmcb200_ctx* ctx,
                              const double* x_host, int64_t n, int64_t ldx, int d,
                              const double* records_host, const int* cols_host, int kl,
                              int k_out, int mode, double max_init,
                              double* logq_host, double* lp_host, double* resp_host, double* aux_host,
                              const double* weights_host, double* sums_host,
                              int64_t chunk_rows);

/* ---- host -> device upload of a sample matrix --------------------------------------------------------
 * Copies rows x d doubles (row stride ld_src) from host memory into a contiguous device matrix.  Pageable host
 * memory is moved by several host threads through pinned bounce buffers on two streams (3x the throughput of a
 * plain cudaMemcpy from pageable memory); page-locked memory is copied directly.  Synchronous.  This is what
 * PMC / GaussianInference use to make the caller's ndarray device resident once (pmc.pyx:362, 447 keeps the
 * samples fixed over all EM steps).
 */
int pmcb200_upload(pmcb200_ctx* ctx, double* dst_dev, const double* src_host, int64_t rows, int d, int64_t ld_src);

/* ---- K3: draw samples from the mixture on the device ------------------------------------------------
 * Replaces MixtureDensity.propose density/mixture.pyx:159-212 with Gauss.propose density/gauss.pyx:159-163 /
 * StudentT.propose density/student_t.pyx:49-55,172-176 underneath (one Python-level draw per sample there).
 * starts_host [k+1]: first row of each component's block, from the caller's multinomial counts
 * (mixture.pyx:193) -- rows [starts[c], starts[c+1]) are x = mean_c + L_c z [* sqrt(dof_c / chi2(dof_c))].
 * Variates: Philox4x32-10, subsequence = index0 + row, so the result does not depend on the launch geometry;
 * give each rank index0 = its global row offset.  latent_dev (int32 [n]) may be NULL.  chol is the lower
 * Cholesky factor of every covariance ([k, d, d] row-major), dofs_dev NULL for a Gaussian mixture.
 */
int pmcb200_mixture_propose(pmcb200_ctx* ctx, int64_t n, int d, int k,
                            const double* means_dev, const double* chol_dev, const double* dofs_dev,
                            const int64_t* starts_host, uint64_t seed, uint64_t index0,
                            double* x_dev, int64_t ldx, int* latent_dev, void* stream);

/* ---- K4: importance weights of a run and the weight-vector reductions ---------------------------------
 * Replaces ImportanceSampler._calculate_weights sampler/importance_sampling.py:197-215 (w_n = exp(log target(x_n) -
 * log q(x_n)), one Python-level evaluation per sample there) and the N-sized passes of perp / ess
 * tools/convergence.py:6-72 and of the weighted log-likelihood mix_adapt/pmc.pyx:388-391.
 * log_target_dev [n] may be NULL: logq_dev then already holds log w.  w_dev [n] may be NULL (sums only).
 * sums_dev [5] = { sum w, sum w log q, sum w^2, sum w log w, number of nonzero w }  (zero weights contribute nothing to
 * the logarithmic sums, convergence.py:30-34); with them
 *   perp = exp(log(S0) - S3 / S0) / n,   ess = S0^2 / (n S2).
 */
int pmcb200_importance_weights(pmcb200_ctx* ctx, const double* log_target_dev, const double* logq_dev, int64_t n,
                               double* w_dev, double* sums_dev, void* stream);

/* ---- measurement helpers ---------------------------------------------------------------------------
 * FP64 FMA throughput of this device (register-resident DFMA chains on every SM), the roof that bounds
 * K1/K2.  which: 0 = DFMA only, 1 = DFMA + one broadcast LDS.128 per 4 DFMA,
 * 2 = DFMA + one broadcast LDS.128 per 2 DFMA, 3 = FP64 mma.sync m8n8k4 (DMMA).
 */
int pmcb200_fp64_peak(pmcb200_ctx* ctx, int which, int iters, double* gflops_out, double* ms_out);

/* Name of the K1 kernel that did the work in the last pmcb200_mixture_eval call on this context, e.g.
 * "k1_mma_eval<4, 2, 16, false>", "k1_fast_eval<30>" or "k1_mixture_eval<30>" -- the three forms are chosen on the
 * device (k1_prepare / k1_mma_prepare), so this reads the decision flags back (synchronises the device).  Writes a
 * NUL-terminated string of at most len - 1 characters; for bench.py's roofline record. */
int pmcb200_last_k1_kernel(pmcb200_ctx* ctx, char* buf, int len);

/* kernels launched by this library since the context was created (for bench.py's gpu_launches) */
int64_t pmcb200_launch_count(pmcb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PMCB200_H */
