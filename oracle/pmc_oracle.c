/*
 * oracle/pmc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded, CPU restatement of the reference's (pypmc v1.2.6,
 * commit 9e0ab49) mixture-density / proposal-update hot path.  Every function
 * cites the reference file:line whose arithmetic (operation order included,
 * as far as C allows) it restates.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (pypmc_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks these functions against
 * (a) the hand-computed golden numbers in the reference's own unit tests and
 * (b) fixtures under tests/golden/ produced by the compiled, unmodified
 * reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o libpmc_oracle.so pmc_oracle.c -lm
 * (-ffp-contract=off: the reference is built by gcc for baseline x86-64 and
 *  therefore never fuses a multiply with an add.)
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define ORC_TINY 2.2250738585072014e-308 /* numpy.finfo('d').tiny */

/* pypmc/tools/_linalg.pyx:10-39  bilinear_sym: x^T M x over the lower triangle,
 * diagonal term first, each off-diagonal term counted twice. */
double orc_bilinear_sym(const double *m, ptrdiff_t ld, const double *v, int d)
{
    double res = 0.0;
    for (int i = 0; i < d; ++i) {
        res += v[i] * v[i] * m[i * ld + i];
        for (int j = 0; j < i; ++j)
            res += 2. * v[i] * v[j] * m[i * ld + j];
    }
    return res;
}

/* pypmc/density/gauss.pyx:132-153  Gauss.multi_evaluate:
 * out[n] = log_normalization - 0.5 * bilinear_sym(inv_sigma, x[n] - mu).
 * `out` may be a strided column (mixture.pyx:145 passes individual[:,k]). */
void orc_gauss_multi_evaluate(const double *x, int64_t n_samples, ptrdiff_t ldx, int d,
                              const double *mu, const double *inv_sigma, double log_norm,
                              double *out, ptrdiff_t out_stride)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int64_t n = 0; n < n_samples; ++n) {
        for (int i = 0; i < d; ++i)
            diff[i] = x[n * ldx + i] - mu[i];
        out[n * out_stride] = log_norm - 0.5 * orc_bilinear_sym(inv_sigma, d, diff, d);
    }
    free(diff);
}

/* pypmc/density/student_t.pyx:135-166  StudentT.multi_evaluate, step by step:
 * r = bilinear; r *= inv_dof; r += 1; r = log(r); r *= prefactor; r += log_norm. */
void orc_student_t_multi_evaluate(const double *x, int64_t n_samples, ptrdiff_t ldx, int d,
                                  const double *mu, const double *inv_sigma, double log_norm,
                                  double prefactor, double inv_dof,
                                  double *out, ptrdiff_t out_stride)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int64_t n = 0; n < n_samples; ++n) {
        for (int i = 0; i < d; ++i)
            diff[i] = x[n * ldx + i] - mu[i];
        double r = orc_bilinear_sym(inv_sigma, d, diff, d);
        r *= inv_dof;
        r += 1.;
        r = log(r);
        r *= prefactor;
        r += log_norm;
        out[n * out_stride] = r;
    }
    free(diff);
}

/* pypmc/tools/_regularize.pyx:19-55  logsumexp (1-D, weighted). */
double orc_logsumexp(const double *a, const double *w, int64_t len)
{
    double max_val = -DBL_MAX, res = 0.0;
    for (int64_t i = 0; i < len; ++i)
        if (a[i] > max_val)
            max_val = a[i];
    for (int64_t i = 0; i < len; ++i)
        res += w[i] * exp(a[i] - max_val);
    return log(res) + max_val;
}

/* pypmc/tools/_regularize.pyx:57-83  logsumexp2D: row-wise weighted LSE; the
 * maximum runs over ALL K columns (also columns whose weight is zero). */
void orc_logsumexp2D(const double *a, int64_t n_rows, int k_cols, ptrdiff_t lda,
                     const double *w, double *res)
{
    for (int64_t n = 0; n < n_rows; ++n) {
        double max_val = -DBL_MAX, acc = 0.0;
        for (int k = 0; k < k_cols; ++k)
            if (a[n * lda + k] > max_val)
                max_val = a[n * lda + k];
        for (int k = 0; k < k_cols; ++k)
            acc += w[k] * exp(a[n * lda + k] - max_val);
        res[n] = log(acc) + max_val;
    }
}

/* pypmc/mix_adapt/pmc.pyx:23-43  calculate_rho_rb, the part after
 * multi_evaluate: on entry rho[n,k] holds the component log-pdf for live k and
 * 0 for dead k; on exit rho[n,k] = exp(lp)*w_k / (exp(log_den[n]) + tiny). */
void orc_rho_rb_inplace(double *rho, int64_t n_samples, int k_comp, const double *w,
                        const int *live, int n_live, double *log_den)
{
    orc_logsumexp2D(rho, n_samples, k_comp, k_comp, w, log_den);
    for (int l = 0; l < n_live; ++l) {
        int k = live[l];
        for (int64_t n = 0; n < n_samples; ++n) {
            double v = exp(rho[n * k_comp + k]) * w[k];
            v /= exp(log_den[n]) + ORC_TINY;
            rho[n * k_comp + k] = v;
        }
    }
}

/* pypmc/mix_adapt/pmc.pyx:602-610  Student-t gamma_nk with the OLD parameters:
 * gamma[n,k] = (nu_k + D) / (nu_k + bilinear_sym(inv_sigma_k, x_n - mu_k)). */
void orc_student_t_gamma(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                         const double *mu, const double *inv_sigma, const double *dof,
                         const int *live, int n_live, double *gamma)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int l = 0; l < n_live; ++l) {
        int k = live[l];
        for (int64_t n = 0; n < n_samples; ++n) {
            for (int i = 0; i < d; ++i) {
                diff[i] = x[n * ldx + i];
                diff[i] -= mu[k * d + i];
            }
            gamma[n * k_comp + k] =
                (dof[k] + (double)d) / (dof[k] + orc_bilinear_sym(inv_sigma + (size_t)k * d * d, d, diff, d));
        }
    }
    free(diff);
}

/* pypmc/mix_adapt/pmc.pyx:188-222 (Gaussian) and :612-650 (Student-t):
 *   alpha_un[k] = sum_n w_n rho_nk                  ('n,nk->k')
 *   mu_num[k,:] = sum_n w_n rho_nk gamma_nk x_n     ('n,nk,nk,ni->ki')
 *   mu_norm[k]  = sum_n w_n rho_nk gamma_nk         (== alpha_un when gamma is NULL)
 * `w` and `gamma` may be NULL (unweighted / Gaussian).  The reference evaluates
 * these with numpy.einsum, whose internal summation order is unspecified; the
 * oracle sums in sample order. */
void orc_pmc_first_moments(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                           const double *w, const double *rho, const double *gamma,
                           double *alpha_un, double *mu_norm, double *mu_num)
{
    for (int k = 0; k < k_comp; ++k) {
        alpha_un[k] = 0.0;
        mu_norm[k] = 0.0;
        for (int i = 0; i < d; ++i)
            mu_num[k * d + i] = 0.0;
    }
    for (int64_t n = 0; n < n_samples; ++n) {
        double wn = w ? w[n] : 1.0;
        for (int k = 0; k < k_comp; ++k) {
            double u = wn * rho[n * k_comp + k];
            alpha_un[k] += u;
            if (gamma)
                u *= gamma[n * k_comp + k];
            mu_norm[k] += u;
            for (int i = 0; i < d; ++i)
                mu_num[k * d + i] += u * x[n * ldx + i];
        }
    }
}

/* pypmc/mix_adapt/pmc.pyx:199-204 / :218-222 / :625-630 / :646-650: second pass,
 * centred on the NEW mean:
 *   cov_un[k] = sum_n w_n rho_nk [gamma_nk] (x_n - mu_k)(x_n - mu_k)^T, live k only.
 * The caller scales by 1/regularize(alpha_un[k]). */
void orc_pmc_second_moments(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                            const double *w, const double *rho, const double *gamma,
                            const double *mu_new, const int *live, int n_live, double *cov_un)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int l = 0; l < n_live; ++l) {
        int k = live[l];
        double *c = cov_un + (size_t)k * d * d;
        for (int i = 0; i < d * d; ++i)
            c[i] = 0.0;
        for (int64_t n = 0; n < n_samples; ++n) {
            double u = (w ? w[n] : 1.0) * rho[n * k_comp + k];
            if (gamma)
                u *= gamma[n * k_comp + k];
            for (int i = 0; i < d; ++i)
                diff[i] = x[n * ldx + i] - mu_new[k * d + i];
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    c[i * d + j] += u * diff[i] * diff[j];
        }
    }
    free(diff);
}

/* pypmc/mix_adapt/pmc.pyx:654-691  degree-of-freedom statistic:
 *   t = log(.5 (q + nu)); t -= psi(.5 (D + nu)); t *= rho;
 *   t += (1 - rho)(log(.5 nu) - psi(.5 nu)); t += rho (D + nu)/(q + nu); t += 1 - rho
 *   out[k] = sum_n [w_n] t_nk      (the caller forms 1 - out/weight_normalization)
 * psi_half_d_plus_nu[k] = digamma(.5(D+nu_k)), psi_half_nu[k] = digamma(.5 nu_k) are
 * computed by the caller with scipy.special.digamma as in the reference. */
void orc_student_t_dof_stat(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                            const double *mu, const double *inv_sigma, const double *dof,
                            const double *psi_half_d_plus_nu, const double *psi_half_nu,
                            const double *w, const double *rho, const int *live, int n_live,
                            double *out)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    double dd = (double)d;
    for (int l = 0; l < n_live; ++l) {
        int k = live[l];
        double nu = dof[k], acc = 0.0;
        for (int64_t n = 0; n < n_samples; ++n) {
            for (int i = 0; i < d; ++i) {
                diff[i] = x[n * ldx + i];
                diff[i] -= mu[k * d + i];
            }
            double q = orc_bilinear_sym(inv_sigma + (size_t)k * d * d, d, diff, d);
            double r = rho[n * k_comp + k];
            double t = log(.5 * (q + nu));
            t -= psi_half_d_plus_nu[k];
            t *= r;
            t += (1. - r) * (log(.5 * nu) - psi_half_nu[k]);
            t += r * (dd + nu) / (q + nu);
            t += (1. - r);
            acc += w ? t * w[n] : t;
        }
        out[k] = acc;
    }
    free(diff);
}

/* pypmc/mix_adapt/variational.pyx:774-798  _update_expectation_gauss_exponent:
 * E[n,k] = D/beta_k + nu_k * bilinear_sym(W_k, x_n - m_k). */
void orc_vb_gauss_exponent(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                           const double *m, const double *w_mat, const double *beta,
                           const double *nu, double *e_out)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int k = 0; k < k_comp; ++k)
        for (int64_t n = 0; n < n_samples; ++n) {
            for (int i = 0; i < d; ++i)
                diff[i] = x[n * ldx + i] - m[k * d + i];
            e_out[n * k_comp + k] =
                (double)d / beta[k] + nu[k] * orc_bilinear_sym(w_mat + (size_t)k * d * d, d, diff, d);
        }
    free(diff);
}

/* pypmc/mix_adapt/variational.pyx:675-691 (_update_log_rho) followed by
 * :711-757 (_update_r): log_rho = E[ln pi] + 0.5(E[ln det Lambda] - D ln 2pi - E);
 * r = softmax over k (max-shifted, multiplied by 1/norm), exact zeros -> tiny,
 * log_rho overwritten by log_rho - max + log(1/norm). */
void orc_vb_update_r(const double *e_gauss, int64_t n_samples, int k_comp, int d,
                     const double *e_ln_pi, const double *e_det_ln_lambda,
                     double *log_rho, double *r)
{
    double dlog = (double)d * log(2. * 3.141592653589793);
    for (int64_t n = 0; n < n_samples; ++n) {
        double *lr = log_rho + n * k_comp, *rr = r + n * k_comp;
        for (int k = 0; k < k_comp; ++k)
            lr[k] = e_ln_pi[k] + 0.5 * (e_det_ln_lambda[k] - dlog - e_gauss[n * k_comp + k]);
        double max = lr[0];
        for (int k = 1; k < k_comp; ++k)
            if (lr[k] > max)
                max = lr[k];
        double norm = 0.0;
        for (int k = 0; k < k_comp; ++k) {
            lr[k] -= max;
            rr[k] = exp(lr[k]);
            norm += rr[k];
        }
        double norm_inv = 1. / norm;
        double log_norm_inv = log(norm_inv);
        for (int k = 0; k < k_comp; ++k) {
            rr[k] *= norm_inv;
            if (rr[k] == 0.0)
                rr[k] = ORC_TINY;
            lr[k] += log_norm_inv;
        }
    }
}

/* pypmc/mix_adapt/variational.pyx:699-709 (N_comp), :806-853 (x_mean_comp),
 * :855-932 (S), unweighted and weighted variants:
 *   N_k = sum_n [w_n] r_nk ; inv = 1/regularize(N_k)
 *   xbar_k = (sum_n [w_n] r_nk x_n) * inv
 *   S_k = (sum_n [w_n] r_nk (x_n - xbar_k)(x_n - xbar_k)^T) * inv, lower triangle
 *         accumulated then mirrored.
 * N_comp is returned un-regularised, as in the reference (regularize acts on the
 * attribute in place; a zero becomes tiny there too, so we do the same). */
void orc_vb_statistics(const double *x, int64_t n_samples, ptrdiff_t ldx, int d, int k_comp,
                       const double *w, const double *r,
                       double *n_comp, double *x_mean, double *s_out)
{
    double *diff = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
    for (int k = 0; k < k_comp; ++k) {
        double acc = 0.0;
        for (int64_t n = 0; n < n_samples; ++n)
            acc += (w ? w[n] : 1.0) * r[n * k_comp + k];
        if (acc == 0.0)
            acc = ORC_TINY; /* regularize(), _regularize.pyx:6-17 */
        n_comp[k] = acc;
        double inv = 1. / acc;

        double *xm = x_mean + (size_t)k * d;
        for (int i = 0; i < d; ++i)
            xm[i] = 0.0;
        for (int64_t n = 0; n < n_samples; ++n) {
            double u = w ? w[n] * r[n * k_comp + k] : r[n * k_comp + k];
            for (int i = 0; i < d; ++i)
                xm[i] += u * x[n * ldx + i];
        }
        for (int i = 0; i < d; ++i)
            xm[i] *= inv;

        double *s = s_out + (size_t)k * d * d;
        for (int i = 0; i < d * d; ++i)
            s[i] = 0.0;
        for (int64_t n = 0; n < n_samples; ++n) {
            for (int i = 0; i < d; ++i)
                diff[i] = x[n * ldx + i] - xm[i];
            double u = w ? w[n] * r[n * k_comp + k] : r[n * k_comp + k];
            for (int i = 0; i < d; ++i)
                for (int j = 0; j <= i; ++j)
                    s[i * d + j] += u * diff[i] * diff[j];
        }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j <= i; ++j) {
                s[i * d + j] *= inv;
                s[j * d + i] = s[i * d + j];
            }
    }
    free(diff);
}

/* pypmc/mix_adapt/variational.pyx:1003-1013  E[log q(Z)] = sum_nk [w_n] r_nk log_rho_nk. */
double orc_vb_log_q_z(const double *r, const double *log_rho, const double *w,
                      int64_t n_samples, int k_comp)
{
    double acc = 0.0;
    for (int64_t n = 0; n < n_samples; ++n) {
        double row = 0.0;
        for (int k = 0; k < k_comp; ++k)
            row += r[n * k_comp + k] * log_rho[n * k_comp + k];
        acc += w ? w[n] * row : row;
    }
    return acc;
}
