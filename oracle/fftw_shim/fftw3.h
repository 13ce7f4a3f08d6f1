/*
 * TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
 *
 * Minimal stand-in for the slice of the FFTW3 single-precision API that the
 * reference calls (fft_transform.c:87-105,135-147,208,218 in /root/reference).
 * FFTW3f is an un-vendored third-party dependency of the reference
 * (meson.build:94, no version pin) and is absent from this image, so the
 * reference sources are compiled against this header when building
 * oracle/_ref/.  The transforms are evaluated in double precision and rounded
 * to float once (see fftw_shim.c), i.e. at least as close to the exact DFT as
 * FFTW's float codelets.
 */
#ifndef SPECBLEACH_ORACLE_FFTW3_SHIM_H
#define SPECBLEACH_ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FFTW_R2HC = 0,
  FFTW_HC2R = 1
} fftwf_r2r_kind;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

typedef struct sb_shim_plan_s* fftwf_plan;

void* fftwf_malloc(size_t n);
void fftwf_free(void* p);
fftwf_plan fftwf_plan_r2r_1d(int n, float* in, float* out, fftwf_r2r_kind kind,
                             unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

#ifdef __cplusplus
}
#endif
#endif
