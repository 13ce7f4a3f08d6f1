/*
 * specbleach-b200 -- drop-in C ABI for the 1D spectral denoiser.
 *
 * Every declaration below is ABI-identical to the reference's public header
 * (reference: include/specbleach_denoiser.h:33-141): same symbol names, same
 * argument order and types, same 44-byte parameter struct passed BY VALUE.
 * Only the implementation behind the handle differs: frames are processed in
 * batches by CUDA kernels on a B200.
 *
 * Error behaviour mirrors the reference: NULL / false / 0, never errno.
 *
 * Attribution: the interface declared here (names, types, struct layout) is that of
 * libspecbleach - A spectral processing library, Copyright 2022 Luciano Dato
 * <lucianodato@gmail.com>, distributed under the GNU Lesser General Public License,
 * version 2.1 or (at your option) any later version, WITHOUT ANY WARRANTY; see
 * <https://www.gnu.org/licenses/old-licenses/lgpl-2.1.html>.  This header restates that
 * interface so that programs written against libspecbleach link unchanged; it is
 * distributed under the same licence.
 */
#ifndef SPECBLEACH_DENOISER_H_INCLUDED
#define SPECBLEACH_DENOISER_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

#include <stdbool.h>
#include <stdint.h>

typedef void* SpectralBleachHandle;

/* Replaces: reference include/specbleach_denoiser.h:33-79 (field order, types
 * and units are part of the ABI). */
typedef struct SpectralBleachDenoiserParameters {
  int learn_noise;             /* >0: capture all four noise profiles          */
  bool residual_listen;        /* output = input - denoised                    */
  float reduction_amount;      /* dB, 0..40                                    */
  float smoothing_factor;      /* %, temporal release smoothing of the spectrum */
  float whitening_factor;      /* %, 0..100                                    */
  int adaptive_noise;          /* !=0: time-recursive noise tracking           */
  int noise_estimation_method; /* 0 SPP-MMSE, 1 Brandt, 2 Martin               */
  float masking_depth;         /* 0..1, psychoacoustic veto depth              */
  float suppression_strength;  /* %, Berouti oversubtraction strength          */
  float aggressiveness;        /* -1 (median) .. 0 (mean) .. +1 (max) profile  */
  float tonal_reduction;       /* dB, extra reduction at tonal noise bins      */
} SpectralBleachDenoiserParameters;

/* reference specbleach_denoiser.h:86 / src/processors/specbleach_denoiser.c:39-82.
 * NULL unless 4000 <= sample_rate <= 192000 and frame_size (ms) > 0. */
SpectralBleachHandle specbleach_initialize(uint32_t sample_rate, float frame_size);

/* reference specbleach_denoiser.h:91 */
void specbleach_free(SpectralBleachHandle instance);

/* reference specbleach_denoiser.h:97 / specbleach_denoiser.c:199-229 */
bool specbleach_load_parameters(SpectralBleachHandle instance,
                                SpectralBleachDenoiserParameters parameters);

/* reference specbleach_denoiser.h:103 / specbleach_denoiser.c:114-127.
 * input/output are HOST buffers of number_of_samples floats and may alias. */
bool specbleach_process(SpectralBleachHandle instance, uint32_t number_of_samples,
                        const float* input, float* output);

/* reference specbleach_denoiser.h:109: latency == frame size in samples */
uint32_t specbleach_get_latency(SpectralBleachHandle instance);

/* reference specbleach_denoiser.h:114: fft_size/2+1 for the 1D denoiser */
uint32_t specbleach_get_noise_profile_size(SpectralBleachHandle instance);

/* reference specbleach_denoiser.h:119; mode 1..4 = mean, median, max, minimum */
bool specbleach_load_noise_profile_for_mode(SpectralBleachHandle instance,
                                            const float* restored_profile,
                                            uint32_t profile_size,
                                            uint32_t block_count, int mode);

/* reference specbleach_denoiser.h:126 */
bool specbleach_reset_noise_profile(SpectralBleachHandle instance);

/* reference specbleach_denoiser.h:132 */
uint32_t specbleach_get_noise_profile_block_count_for_mode(
    SpectralBleachHandle instance, int mode);

/* reference specbleach_denoiser.h:137: borrowed pointer to a host mirror that
 * is refreshed after every learn-mode process() call. */
float* specbleach_get_noise_profile_for_mode(SpectralBleachHandle instance, int mode);

/* reference specbleach_denoiser.h:141 */
bool specbleach_noise_profile_available_for_mode(SpectralBleachHandle instance,
                                                 int mode);

#ifdef __cplusplus
}
#endif
#endif /* SPECBLEACH_DENOISER_H_INCLUDED */
