/*
 * oracle/shim/gsl/gsl_integration.h -- TEST INFRASTRUCTURE, not product code.
 *
 * GSL is not installed in this image.  The reference's Gauss-Legendre class
 * (/root/reference/Quadratures/GaussLegendre.hpp:5,14-23) needs exactly three
 * GSL calls; this header-only stand-in provides them with GSL's documented
 * behaviour:
 *   - the table holds the non-negative nodes of the n-point rule on [-1,1]
 *     in ascending order with their weights ((n+1)/2 entries);
 *   - gsl_integration_glfixed_point(a,b,i,...) unpacks them into the i-th
 *     node/weight of the rule on [a,b], nodes ascending in i, with
 *     x = B -/+ A*xhat, w = A*what, A=(b-a)/2, B=(a+b)/2.
 * Nodes are found by Newton iteration on P_n in long double (any correctly
 * rounded Gauss-Legendre rule agrees with GSL's tables to ~1 ulp).
 */
#ifndef BFSM_ORACLE_GSL_INTEGRATION_SHIM_H
#define BFSM_ORACLE_GSL_INTEGRATION_SHIM_H

#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    size_t n;
    double *x;
    double *w;
    int precomputed;
} gsl_integration_glfixed_table;

static inline gsl_integration_glfixed_table *gsl_integration_glfixed_table_alloc(size_t n)
{
    if (n == 0) return NULL;
    gsl_integration_glfixed_table *t =
        (gsl_integration_glfixed_table *)malloc(sizeof(gsl_integration_glfixed_table));
    if (!t) return NULL;
    const size_t m = (n + 1) / 2;
    t->n = n;
    t->precomputed = 0;
    t->x = (double *)malloc(sizeof(double) * m);
    t->w = (double *)malloc(sizeof(double) * m);
    if (!t->x || !t->w) {
        free(t->x);
        free(t->w);
        free(t);
        return NULL;
    }
    const long double PI_L = 3.14159265358979323846264338327950288L;
    /* k-th largest root (k = 1..m) -> stored descending then reversed. */
    for (size_t k = 1; k <= m; ++k) {
        long double z = cosl(PI_L * ((long double)k - 0.25L) / ((long double)n + 0.5L));
        long double pp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1.0L, p1 = z;
            if (n == 1) {
                p1 = z;
                p0 = 1.0L;
            }
            for (size_t j = 2; j <= n; ++j) {
                long double p2 = ((2.0L * j - 1.0L) * z * p1 - (j - 1.0L) * p0) / (long double)j;
                p0 = p1;
                p1 = p2;
            }
            /* p1 = P_n(z), p0 = P_{n-1}(z) */
            pp = (long double)n * (z * p1 - p0) / (z * z - 1.0L);
            long double dz = p1 / pp;
            z -= dz;
            if (fabsl(dz) < 1e-20L) break;
        }
        {
            /* recompute derivative at the converged root for the weight */
            long double p0 = 1.0L, p1 = z;
            for (size_t j = 2; j <= n; ++j) {
                long double p2 = ((2.0L * j - 1.0L) * z * p1 - (j - 1.0L) * p0) / (long double)j;
                p0 = p1;
                p1 = p2;
            }
            pp = (long double)n * (z * p1 - p0) / (z * z - 1.0L);
        }
        long double w = 2.0L / ((1.0L - z * z) * pp * pp);
        if ((n & 1) && k == m) z = 0.0L; /* middle root of an odd rule is exactly 0 */
        t->x[m - k] = (double)z;
        t->w[m - k] = (double)w;
    }
    return t;
}

static inline int gsl_integration_glfixed_point(double a, double b, size_t i, double *xi,
                                                double *wi,
                                                const gsl_integration_glfixed_table *t)
{
    const double A = (b - a) / 2;
    const double B = (a + b) / 2;
    if (i >= t->n) return 1;
    if (t->n & 1) {
        const long k = (long)i - (long)(t->n / 2);
        if (k < 0) {
            *xi = B - A * t->x[-k];
            *wi = A * t->w[-k];
        } else {
            *xi = B + A * t->x[k];
            *wi = A * t->w[k];
        }
    } else if (i < t->n / 2) {
        const size_t k = (t->n / 2) - 1 - i;
        *xi = B - A * t->x[k];
        *wi = A * t->w[k];
    } else {
        const size_t k = i - t->n / 2;
        *xi = B + A * t->x[k];
        *wi = A * t->w[k];
    }
    return 0;
}

static inline void gsl_integration_glfixed_table_free(gsl_integration_glfixed_table *t)
{
    if (!t) return;
    free(t->x);
    free(t->w);
    free(t);
}

#ifdef __cplusplus
}
#endif

#endif /* BFSM_ORACLE_GSL_INTEGRATION_SHIM_H */
