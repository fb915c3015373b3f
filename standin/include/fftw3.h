/*
 * Stand-in for <fftw3.h> (single precision subset).
 *
 * FFTW3 is not installed in this image (no headers, no libfftw3f, no network).
 * This header declares exactly the fftwf_* entry points that the reference
 * calls (cli/phase-rotate.cc:137-155,164,193,205,1009 and
 * src/phaserotate.c:138,164,188-217,307-308,333,345-346,361-364,380,400,648,657)
 * so that the UNMODIFIED reference sources compile.  The implementation is in
 * standin/src/fftw3_standin.c: an unnormalised real DFT with FFTW's r2c/c2r
 * conventions, computed in double precision internally (exact DFT rounded once
 * to float) unless built with -DSTANDIN_FFT_FLOAT.
 *
 * It is test/oracle infrastructure only.  The product never links it.
 */
#ifndef STANDIN_FFTW3_H
#define STANDIN_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct standin_fftwf_plan_s* fftwf_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void* fftwf_malloc (size_t n);
void  fftwf_free (void* p);

fftwf_plan fftwf_plan_dft_r2c_1d (int n, float* in, fftwf_complex* out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d (int n, fftwf_complex* in, float* out, unsigned flags);

void fftwf_execute_dft_r2c (const fftwf_plan p, float* in, fftwf_complex* out);
void fftwf_execute_dft_c2r (const fftwf_plan p, fftwf_complex* in, float* out);

void fftwf_destroy_plan (fftwf_plan p);
void fftwf_cleanup (void);

#ifdef __cplusplus
}
#endif
#endif
