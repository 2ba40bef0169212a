/*
 * oracle/bfsm_oracle.c -- TEST INFRASTRUCTURE ("port" oracle), not product code.
 *
 * Plain-C restatement of the reference CPU algorithm
 *   BoltzmannOperator<FFTW_Backend>::computeCollision
 *   (/root/reference/Collisions/FFTWBoltzmannOperator.cpp:147-334)
 * with the same formulas, the same loop nest per quadrature pair and the same
 * unnormalised FFT conventions, but STREAMING: each thread owns four grid-sized
 * scratch arrays instead of the reference's six P-sized batch arrays
 * (FFTWBoltzmannOperator.cpp:30-37, 96*N*P bytes = 154.6 GB at 64^3 / 32x192),
 * so every BASELINE.json configuration fits in host memory.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *  (1) oracle/_ref/libbfsm_ref.so = the unmodified reference sources, and
 *  (2) the known answers printed in /root/reference/Results/
 *      maxwell_bkw_fftw_atomics.txt:19-21, 371-373 (committed under tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (CUDA) path never does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "shim/fftw3.h"
#include "shim/gsl/gsl_integration.h"

static const double pi = 3.14159265358979323846; /* Utilities/constants.hpp:7 */

/* FFTWBoltzmannOperator.hpp:17-21 */
static double sincc(double x)
{
    const double eps = 2.220446049250313e-16; /* std::numeric_limits<double>::epsilon() */
    return sin(x + eps) / (x + eps);
}

/* Fourier mode table, FFTWBoltzmannOperator.cpp:50-57 */
static int *mode_table(int n)
{
    int *l = (int *)malloc(sizeof(int) * (size_t)n);
    int c = 0;
    for (int i = 0; i < n / 2; ++i) l[c++] = i;
    for (int i = -n / 2; i < 0; ++i) l[c++] = i;
    return l;
}

typedef struct {
    int nvx, nvy, nvz, n_r, n_s;
    const double *rho, *w_r, *sx, *sy, *sz, *w_s;
    double gamma, b_gamma, L;
} oracle_cfg;

/*
 * Gain spectrum Q_gain_hat restricted to pairs b = r*n_s + s in
 * [pair_begin, pair_end) -- steps 0..5 of the reference (cpp:168-276).
 * f_hat_out (N complex) receives FFT3(f) for the caller's loss/combine step.
 */
static int gain_hat_range(const oracle_cfg *c, const double *f_in, int pair_begin, int pair_end,
                          fftw_complex *f_hat, fftw_complex *Q_gain_hat)
{
    const int Nvx = c->nvx, Nvy = c->nvy, Nvz = c->nvz;
    const int N_spherical = c->n_s;
    const int grid_size = Nvx * Nvy * Nvz;
    const double fft_scale = 1.0 / grid_size; /* cpp:162 */
    const double L = c->L;

    int *lx = mode_table(Nvx), *ly = mode_table(Nvy), *lz = mode_table(Nvz);
    fftw_complex *f = fftw_alloc_complex((size_t)grid_size);
    fftw_plan fft_plan = fftw_plan_dft_3d(Nvx, Nvy, Nvz, f, f_hat, FFTW_FORWARD, FFTW_ESTIMATE);
    fftw_plan ifft_plan = fftw_plan_dft_3d(Nvx, Nvy, Nvz, f_hat, f, FFTW_BACKWARD, FFTW_ESTIMATE);
    if (!lx || !ly || !lz || !f || !fft_plan || !ifft_plan) return 1;

    /* cpp:168-186 */
    for (int idx3 = 0; idx3 < grid_size; ++idx3) {
        f[idx3][0] = f_in[idx3];
        f[idx3][1] = 0;
        Q_gain_hat[idx3][0] = 0;
        Q_gain_hat[idx3][1] = 0;
    }
    fftw_execute_dft(fft_plan, f, f_hat);

    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* per-thread partial accumulators, summed in thread order afterwards
     * (the reference uses `omp atomic`, cpp:268-271: same sum, order unspecified) */
    fftw_complex *partial = fftw_alloc_complex((size_t)grid_size * (size_t)nthreads);
    if (!partial) return 1;
    memset(partial, 0, sizeof(fftw_complex) * (size_t)grid_size * (size_t)nthreads);
    int failed = 0;

#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        fftw_complex *acc = partial + (size_t)tid * grid_size;
        fftw_complex *a1_hat = fftw_alloc_complex((size_t)grid_size);
        fftw_complex *a2_hat = fftw_alloc_complex((size_t)grid_size);
        fftw_complex *a1 = fftw_alloc_complex((size_t)grid_size);
        fftw_complex *a2 = fftw_alloc_complex((size_t)grid_size);
        if (!a1_hat || !a2_hat || !a1 || !a2) {
#pragma omp atomic write
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int b = pair_begin; b < pair_end; ++b) {
                const int r = b / N_spherical, s = b % N_spherical; /* cpp:196 */
                const double rho = c->rho[r];

                /* step 1, cpp:198-225 */
                for (int i = 0; i < Nvx; ++i)
                    for (int j = 0; j < Nvy; ++j)
                        for (int k = 0; k < Nvz; ++k) {
                            const int idx3 = (i * Nvy + j) * Nvz + k;
                            const double l_dot_sigma =
                                lx[i] * c->sx[s] + ly[j] * c->sy[s] + lz[k] * c->sz[s];
                            const double tmp = -(pi / (2 * L)) * rho * l_dot_sigma;
                            const double a_re = cos(tmp), a_im = sin(tmp);
                            const double b_re = f_hat[idx3][0], b_im = f_hat[idx3][1];
                            a1_hat[idx3][0] = fft_scale * (a_re * b_re - a_im * b_im);
                            a1_hat[idx3][1] = fft_scale * (a_re * b_im + a_im * b_re);
                            a2_hat[idx3][0] = fft_scale * (a_re * b_re + a_im * b_im);
                            a2_hat[idx3][1] = fft_scale * (a_re * b_im - a_im * b_re);
                        }

                /* step 2, cpp:229-230 */
                fftw_execute_dft(ifft_plan, a1_hat, a1);
                fftw_execute_dft(ifft_plan, a2_hat, a2);

                /* step 3, cpp:233-246 (product written over a1_hat: scratch reuse) */
                for (int idx = 0; idx < grid_size; ++idx) {
                    const double a_re = a1[idx][0], a_im = a1[idx][1];
                    const double b_re = a2[idx][0], b_im = a2[idx][1];
                    a1_hat[idx][0] = a_re * b_re - (a_im * b_im);
                    a1_hat[idx][1] = a_re * b_im + (a_im * b_re);
                }

                /* step 4, cpp:249 */
                fftw_execute_dft(fft_plan, a1_hat, a2_hat);

                /* step 5, cpp:252-273 */
                const double weight =
                    fft_scale * c->w_r[r] * c->w_s[s] * pow(rho, c->gamma + 2);
                for (int i = 0; i < Nvx; ++i)
                    for (int j = 0; j < Nvy; ++j)
                        for (int k = 0; k < Nvz; ++k) {
                            const int idx3 = (i * Nvy + j) * Nvz + k;
                            const double norm_l =
                                sqrt((double)(lx[i] * lx[i] + ly[j] * ly[j] + lz[k] * lz[k]));
                            const double beta1 =
                                4 * pi * c->b_gamma * sincc(pi * rho * norm_l / (2 * L));
                            acc[idx3][0] += weight * beta1 * a2_hat[idx3][0];
                            acc[idx3][1] += weight * beta1 * a2_hat[idx3][1];
                        }
            }
        }
        fftw_free(a1_hat);
        fftw_free(a2_hat);
        fftw_free(a1);
        fftw_free(a2);
    }

    for (int t = 0; t < nthreads; ++t)
        for (int idx = 0; idx < grid_size; ++idx) {
            Q_gain_hat[idx][0] += partial[(size_t)t * grid_size + idx][0];
            Q_gain_hat[idx][1] += partial[(size_t)t * grid_size + idx][1];
        }

    fftw_free(partial);
    fftw_free(f);
    fftw_destroy_plan(fft_plan);
    fftw_destroy_plan(ifft_plan);
    free(lx);
    free(ly);
    free(lz);
    return failed;
}

/* Steps 6-7 of the reference (cpp:281-330): loss term and combine. */
static int loss_and_combine(const oracle_cfg *c, const double *f_in, fftw_complex *f_hat,
                            fftw_complex *Q_gain_hat, double *Q)
{
    const int Nvx = c->nvx, Nvy = c->nvy, Nvz = c->nvz;
    const int grid_size = Nvx * Nvy * Nvz;
    const double fft_scale = 1.0 / grid_size;
    const double L = c->L;
    int *lx = mode_table(Nvx), *ly = mode_table(Nvy), *lz = mode_table(Nvz);
    fftw_complex *b2f_hat = fftw_alloc_complex((size_t)grid_size);
    fftw_complex *b2f = fftw_alloc_complex((size_t)grid_size);
    fftw_complex *Q_gain = fftw_alloc_complex((size_t)grid_size);
    fftw_plan ifft_plan =
        fftw_plan_dft_3d(Nvx, Nvy, Nvz, b2f_hat, b2f, FFTW_BACKWARD, FFTW_ESTIMATE);
    if (!lx || !ly || !lz || !b2f_hat || !b2f || !Q_gain || !ifft_plan) return 1;

    /* cpp:281-299 */
#pragma omp parallel for collapse(2)
    for (int i = 0; i < Nvx; ++i)
        for (int j = 0; j < Nvy; ++j)
            for (int k = 0; k < Nvz; ++k) {
                const int idx3 = (i * Nvy + j) * Nvz + k;
                double beta2 = 0.0;
                const double norm_l =
                    sqrt((double)(lx[i] * lx[i] + ly[j] * ly[j] + lz[k] * lz[k]));
                for (int r = 0; r < c->n_r; ++r)
                    beta2 += 16 * pi * pi * c->b_gamma * c->w_r[r] *
                             pow(c->rho[r], c->gamma + 2) * sincc(pi * c->rho[r] * norm_l / L);
                b2f_hat[idx3][0] = fft_scale * beta2 * f_hat[idx3][0];
                b2f_hat[idx3][1] = fft_scale * beta2 * f_hat[idx3][1];
            }

    /* cpp:304-309 */
    fftw_execute_dft(ifft_plan, Q_gain_hat, Q_gain);
    fftw_execute_dft(ifft_plan, b2f_hat, b2f);

    /* cpp:314-330; f = (f_in, 0) */
    for (int idx3 = 0; idx3 < grid_size; ++idx3) {
        const double a_re = b2f[idx3][0], a_im = b2f[idx3][1];
        const double b_re = f_in[idx3], b_im = 0.0;
        const double loss_re = a_re * b_re - (a_im * b_im);
        Q[idx3] = Q_gain[idx3][0] - loss_re;
    }

    fftw_free(b2f_hat);
    fftw_free(b2f);
    fftw_free(Q_gain);
    fftw_destroy_plan(ifft_plan);
    free(lx);
    free(ly);
    free(lz);
    return 0;
}

/* ------------------------------------------------------------ public API */

void bfsm_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int bfsm_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Gauss-Legendre rule on [a,b], nodes ascending -- Quadratures/GaussLegendre.hpp:10-24. */
int bfsm_oracle_gauss_legendre(int n, double a, double b, double *nodes, double *weights)
{
    gsl_integration_glfixed_table *t = gsl_integration_glfixed_table_alloc((size_t)n);
    if (!t) return 1;
    for (int i = 0; i < n; ++i)
        gsl_integration_glfixed_point(a, b, (size_t)i, &nodes[i], &weights[i], t);
    gsl_integration_glfixed_table_free(t);
    return 0;
}

/* Full evaluation Q = Q(f,f); one call == one reference operator()(Q, f_in). */
int bfsm_oracle_collide(int nvx, int nvy, int nvz, int n_r, const double *rho, const double *w_r,
                        int n_s, const double *sx, const double *sy, const double *sz,
                        const double *w_s, double gamma, double b_gamma, double L,
                        const double *f_in, double *Q)
{
    oracle_cfg c = {nvx, nvy, nvz, n_r, n_s, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L};
    const size_t N = (size_t)nvx * nvy * nvz;
    fftw_complex *f_hat = fftw_alloc_complex(N);
    fftw_complex *Qg_hat = fftw_alloc_complex(N);
    if (!f_hat || !Qg_hat) return 1;
    int rc = gain_hat_range(&c, f_in, 0, n_r * n_s, f_hat, Qg_hat);
    if (!rc) rc = loss_and_combine(&c, f_in, f_hat, Qg_hat, Q);
    fftw_free(f_hat);
    fftw_free(Qg_hat);
    return rc;
}

/* Partial gain spectrum over pairs [pair_begin, pair_end): 2*N doubles (re,im). */
int bfsm_oracle_gain_hat(int nvx, int nvy, int nvz, int n_r, const double *rho,
                         const double *w_r, int n_s, const double *sx, const double *sy,
                         const double *sz, const double *w_s, double gamma, double b_gamma,
                         double L, const double *f_in, int pair_begin, int pair_end,
                         double *Qhat_out)
{
    oracle_cfg c = {nvx, nvy, nvz, n_r, n_s, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L};
    const size_t N = (size_t)nvx * nvy * nvz;
    fftw_complex *f_hat = fftw_alloc_complex(N);
    if (!f_hat) return 1;
    int rc = gain_hat_range(&c, f_in, pair_begin, pair_end, f_hat, (fftw_complex *)Qhat_out);
    fftw_free(f_hat);
    return rc;
}

/* Loss + combine from an externally summed gain spectrum (shard-sum tests). */
int bfsm_oracle_finish(int nvx, int nvy, int nvz, int n_r, const double *rho, const double *w_r,
                       double gamma, double b_gamma, double L, const double *f_in,
                       const double *Qhat_in, double *Q)
{
    oracle_cfg c = {nvx, nvy, nvz, n_r, 0, rho, w_r, NULL, NULL, NULL, NULL, gamma, b_gamma, L};
    const size_t N = (size_t)nvx * nvy * nvz;
    fftw_complex *f = fftw_alloc_complex(N), *f_hat = fftw_alloc_complex(N);
    fftw_complex *Qg_hat = fftw_alloc_complex(N);
    if (!f || !f_hat || !Qg_hat) return 1;
    for (size_t i = 0; i < N; ++i) {
        f[i][0] = f_in[i];
        f[i][1] = 0;
    }
    memcpy(Qg_hat, Qhat_in, sizeof(fftw_complex) * N);
    fftw_plan p = fftw_plan_dft_3d(nvx, nvy, nvz, f, f_hat, FFTW_FORWARD, FFTW_ESTIMATE);
    fftw_execute_dft(p, f, f_hat);
    fftw_destroy_plan(p);
    int rc = loss_and_combine(&c, f_in, f_hat, Qg_hat, Q);
    fftw_free(f);
    fftw_free(f_hat);
    fftw_free(Qg_hat);
    return rc;
}

/* Raw 3-D transform of the FFT stand-in, exposed so tests can pin it against
 * numpy.fft (an independent implementation). sign = -1 forward, +1 backward. */
int bfsm_oracle_fft3(int n0, int n1, int n2, int sign, const double *in, double *out)
{
    fftw_plan p = fftw_plan_dft_3d(n0, n1, n2, NULL, NULL, sign, FFTW_ESTIMATE);
    if (!p) return 1;
    fftw_execute_dft(p, (fftw_complex *)in, (fftw_complex *)out);
    fftw_destroy_plan(p);
    return 0;
}
