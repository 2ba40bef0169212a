/*
 * oracle/shim/fftw3_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for the FFTW3 entry points the reference CPU operator calls
 * (see fftw3.h in this directory for the list and the reference call sites).
 * Nothing in the product path (the CUDA library) links or calls this file.
 *
 * Algorithm: 3-D complex DFT by three 1-D passes.  Each pass gathers LANES
 * lines into a split re/im scratch block [n][LANES], runs a radix-4 Stockham
 * autosort FFT down the first index (inner loop over the LANES columns, which
 * the compiler vectorises), and scatters the block back.  Twiddles are
 * tabulated once per plan with cosl/sinl.  Lengths that are not a power of two
 * fall back to a direct O(n^2) DFT so that any grid the reference accepts
 * still produces a correct answer.
 *
 * fftw_execute_dft() only reads the plan, and uses thread-local scratch, so a
 * single plan may be executed concurrently from many OpenMP threads exactly
 * as the reference does (FFTWBoltzmannOperator.cpp:229-230, 249).
 */
#define _POSIX_C_SOURCE 200112L
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LANES 8

struct bfsm_shim_plan_s {
    int n[3];
    int sign;
    double *wr[3]; /* wr[a][k] = cos(2 pi k / n[a])        */
    double *wi[3]; /* wi[a][k] = sign * sin(2 pi k / n[a]) */
    fftw_complex *in, *out;
};

/* ---------------------------------------------------------------- memory */

void *fftw_malloc(size_t n)
{
    void *p = NULL;
    if (n == 0) n = 64;
    if (posix_memalign(&p, 64, n) != 0) return NULL;
    return p;
}
double *fftw_alloc_real(size_t n) { return (double *)fftw_malloc(n * sizeof(double)); }
fftw_complex *fftw_alloc_complex(size_t n)
{
    return (fftw_complex *)fftw_malloc(n * sizeof(fftw_complex));
}
void fftw_free(void *p) { free(p); }

int fftw_import_wisdom_from_filename(const char *filename)
{
    (void)filename;
    return 0;
}
int fftw_export_wisdom_to_filename(const char *filename)
{
    (void)filename;
    return 1;
}

/* ------------------------------------------------------------------ plan */

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in,
                           fftw_complex *out, int sign, unsigned flags)
{
    (void)flags;
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return NULL;
    struct bfsm_shim_plan_s *p = (struct bfsm_shim_plan_s *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->n[0] = n0;
    p->n[1] = n1;
    p->n[2] = n2;
    p->sign = (sign < 0) ? -1 : +1;
    p->in = in;
    p->out = out;
    for (int a = 0; a < 3; ++a) {
        int n = p->n[a];
        p->wr[a] = (double *)malloc(sizeof(double) * (size_t)n);
        p->wi[a] = (double *)malloc(sizeof(double) * (size_t)n);
        for (int k = 0; k < n; ++k) {
            long double ang = 2.0L * 3.14159265358979323846264338327950288L *
                              (long double)k / (long double)n;
            p->wr[a][k] = (double)cosl(ang);
            p->wi[a][k] = (double)(p->sign * sinl(ang));
        }
    }
    return p;
}

void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    for (int a = 0; a < 3; ++a) {
        free(p->wr[a]);
        free(p->wi[a]);
    }
    free(p);
}

/* ------------------------------------------------ column-batched 1-D FFT */

/* One radix-4 Stockham pass: sequence length ncur, stride s (ncur*s == n). */
static void pass_radix4(int ncur, int s, int sign, const double *restrict wr,
                        const double *restrict wi, const double *restrict xr,
                        const double *restrict xi, double *restrict yr,
                        double *restrict yi)
{
    const int n1 = ncur / 4;
    const int sl = s * LANES;
    for (int p = 0; p < n1; ++p) {
        const double w1r = wr[p * s], w1i = wi[p * s];
        const double w2r = wr[2 * p * s], w2i = wi[2 * p * s];
        const double w3r = wr[3 * p * s], w3i = wi[3 * p * s];
        const double *ar = xr + (size_t)sl * (p), *ai = xi + (size_t)sl * (p);
        const double *br = xr + (size_t)sl * (p + n1), *bi = xi + (size_t)sl * (p + n1);
        const double *cr = xr + (size_t)sl * (p + 2 * n1), *ci = xi + (size_t)sl * (p + 2 * n1);
        const double *dr = xr + (size_t)sl * (p + 3 * n1), *di = xi + (size_t)sl * (p + 3 * n1);
        double *o0r = yr + (size_t)sl * (4 * p + 0), *o0i = yi + (size_t)sl * (4 * p + 0);
        double *o1r = yr + (size_t)sl * (4 * p + 1), *o1i = yi + (size_t)sl * (4 * p + 1);
        double *o2r = yr + (size_t)sl * (4 * p + 2), *o2i = yi + (size_t)sl * (4 * p + 2);
        double *o3r = yr + (size_t)sl * (4 * p + 3), *o3i = yi + (size_t)sl * (4 * p + 3);
        for (int q = 0; q < sl; ++q) {
            const double apcr = ar[q] + cr[q], apci = ai[q] + ci[q];
            const double amcr = ar[q] - cr[q], amci = ai[q] - ci[q];
            const double bpdr = br[q] + dr[q], bpdi = bi[q] + di[q];
            const double bmdr = br[q] - dr[q], bmdi = bi[q] - di[q];
            /* (sign*i)*(b-d) */
            const double jr = -sign * bmdi, ji = sign * bmdr;
            const double t1r = amcr + jr, t1i = amci + ji;
            const double t2r = apcr - bpdr, t2i = apci - bpdi;
            const double t3r = amcr - jr, t3i = amci - ji;
            o0r[q] = apcr + bpdr;
            o0i[q] = apci + bpdi;
            o1r[q] = w1r * t1r - w1i * t1i;
            o1i[q] = w1r * t1i + w1i * t1r;
            o2r[q] = w2r * t2r - w2i * t2i;
            o2i[q] = w2r * t2i + w2i * t2r;
            o3r[q] = w3r * t3r - w3i * t3i;
            o3i[q] = w3r * t3i + w3i * t3r;
        }
    }
}

/* Final radix-2 pass (ncur == 2, no twiddles). */
static void pass_radix2_last(int s, const double *restrict xr, const double *restrict xi,
                             double *restrict yr, double *restrict yi)
{
    const int sl = s * LANES;
    for (int q = 0; q < sl; ++q) {
        const double ar = xr[q], ai = xi[q];
        const double br = xr[q + sl], bi = xi[q + sl];
        yr[q] = ar + br;
        yi[q] = ai + bi;
        yr[q + sl] = ar - br;
        yi[q + sl] = ai - bi;
    }
}

/* Direct DFT for non power-of-two lengths (slow, correctness only). */
static void dft_direct(int n, const double *wr, const double *wi, const double *xr,
                       const double *xi, double *yr, double *yi)
{
    for (int k = 0; k < n; ++k) {
        for (int c = 0; c < LANES; ++c) {
            yr[k * LANES + c] = 0.0;
            yi[k * LANES + c] = 0.0;
        }
        for (int j = 0; j < n; ++j) {
            const int t = (int)(((long long)j * k) % n);
            const double cr = wr[t], ci = wi[t];
            for (int c = 0; c < LANES; ++c) {
                const double ar = xr[j * LANES + c], ai = xi[j * LANES + c];
                yr[k * LANES + c] += cr * ar - ci * ai;
                yi[k * LANES + c] += cr * ai + ci * ar;
            }
        }
    }
}

/* Transforms the LANES columns held in (b0r,b0i); (b1r,b1i) is work space.
 * Returns 0 if the result is in buffer 0, 1 if it is in buffer 1. */
static int fft_columns(int n, int sign, const double *wr, const double *wi, double *b0r,
                       double *b0i, double *b1r, double *b1i)
{
    if (n == 1) return 0;
    if (!is_pow2(n)) {
        dft_direct(n, wr, wi, b0r, b0i, b1r, b1i);
        return 1;
    }
    int which = 0;
    int ncur = n, s = 1;
    while (ncur >= 4) {
        if (which == 0)
            pass_radix4(ncur, s, sign, wr, wi, b0r, b0i, b1r, b1i);
        else
            pass_radix4(ncur, s, sign, wr, wi, b1r, b1i, b0r, b0i);
        which ^= 1;
        ncur /= 4;
        s *= 4;
    }
    if (ncur == 2) {
        if (which == 0)
            pass_radix2_last(s, b0r, b0i, b1r, b1i);
        else
            pass_radix2_last(s, b1r, b1i, b0r, b0i);
        which ^= 1;
    }
    return which;
}

/* --------------------------------------------------------------- execute */

static __thread double *tls_buf = NULL;
static __thread size_t tls_len = 0;

static double *scratch(size_t doubles)
{
    if (tls_len < doubles) {
        free(tls_buf);
        tls_buf = (double *)fftw_malloc(doubles * sizeof(double));
        tls_len = tls_buf ? doubles : 0;
    }
    return tls_buf;
}

/* 1-D pass along one axis.  `nlines` lines of length n; line l starts at
 * base(l) = (l / inner) * outer_stride + (l % inner) and has element stride
 * `estride` (all in units of complex elements). src may equal dst. */
static void axis_pass(const struct bfsm_shim_plan_s *p, int axis, const fftw_complex *src,
                      fftw_complex *dst, size_t nlines, size_t inner, size_t outer_stride,
                      size_t estride)
{
    const int n = p->n[axis];
    double *buf = scratch((size_t)4 * n * LANES);
    double *b0r = buf, *b0i = buf + (size_t)n * LANES;
    double *b1r = buf + (size_t)2 * n * LANES, *b1i = buf + (size_t)3 * n * LANES;

    for (size_t l0 = 0; l0 < nlines; l0 += LANES) {
        const int nl = (int)((nlines - l0 < LANES) ? (nlines - l0) : LANES);
        size_t base[LANES];
        for (int c = 0; c < LANES; ++c) {
            size_t l = l0 + (size_t)((c < nl) ? c : 0);
            base[c] = (l / inner) * outer_stride + (l % inner);
        }
        for (int j = 0; j < n; ++j)
            for (int c = 0; c < LANES; ++c) {
                const double *e = src[base[c] + (size_t)j * estride];
                b0r[j * LANES + c] = e[0];
                b0i[j * LANES + c] = e[1];
            }
        const int w = fft_columns(n, p->sign, p->wr[axis], p->wi[axis], b0r, b0i, b1r, b1i);
        const double *rr = w ? b1r : b0r, *ri = w ? b1i : b0i;
        for (int j = 0; j < n; ++j)
            for (int c = 0; c < nl; ++c) {
                double *e = dst[base[c] + (size_t)j * estride];
                e[0] = rr[j * LANES + c];
                e[1] = ri[j * LANES + c];
            }
    }
}

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out)
{
    const size_t n0 = (size_t)p->n[0], n1 = (size_t)p->n[1], n2 = (size_t)p->n[2];
    /* axis 2 (contiguous): line l = (i,j), base = l*n2, element stride 1 */
    axis_pass(p, 2, (const fftw_complex *)in, out, n0 * n1, 1, n2, 1);
    /* axis 1: lines (i,k): base = i*n1*n2 + k, element stride n2 */
    axis_pass(p, 1, (const fftw_complex *)out, out, n0 * n2, n2, n1 * n2, n2);
    /* axis 0: lines (j,k): base = j*n2 + k, element stride n1*n2 */
    axis_pass(p, 0, (const fftw_complex *)out, out, n1 * n2, n1 * n2, 0, n1 * n2);
}

void fftw_execute(const fftw_plan p) { fftw_execute_dft(p, p->in, p->out); }
