/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE, not product code.
 *
 * FFTW3 is not installed in this image (no fftw3.h, no libfftw3, no network).
 * This header declares the exact subset of the FFTW3 API that the reference's
 * CPU operator uses, so that /root/reference/Collisions/FFTWBoltzmannOperator.cpp
 * and /root/reference/maxwell_bkw_fftw.cpp compile UNMODIFIED against it
 * (call sites: FFTWBoltzmannOperator.cpp:27-68, 186, 229-230, 249, 305, 309,
 * 341-361; maxwell_bkw_fftw.cpp:78-79, 106, 169-171).
 *
 * The implementation (fftw3_shim.c) is an independent power-of-two
 * Stockham FFT written for this repo; it follows FFTW's documented semantics:
 * unnormalised transforms, FFTW_FORWARD = exp(-i...), FFTW_BACKWARD = exp(+i...),
 * row-major n0 x n1 x n2, fftw_execute_dft is thread-safe for a shared plan.
 */
#ifndef BFSM_ORACLE_FFTW3_SHIM_H
#define BFSM_ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct bfsm_shim_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)

#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
double *fftw_alloc_real(size_t n);
fftw_complex *fftw_alloc_complex(size_t n);
void fftw_free(void *p);

fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in,
                           fftw_complex *out, int sign, unsigned flags);
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

/* Wisdom is meaningless for the shim: import reports "nothing imported" (0),
 * export reports success without touching the filesystem. */
int fftw_import_wisdom_from_filename(const char *filename);
int fftw_export_wisdom_to_filename(const char *filename);

#ifdef __cplusplus
}
#endif

#endif /* BFSM_ORACLE_FFTW3_SHIM_H */
