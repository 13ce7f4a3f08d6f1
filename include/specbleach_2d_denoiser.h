/*
 * specbleach-b200 -- drop-in C ABI for the 2D (Non-Local-Means) denoiser.
 *
 * ABI-identical to the reference's public header
 * (reference: include/specbleach_2d_denoiser.h:37-171).  See
 * specbleach_denoiser.h for the conventions.
 *
 * Attribution: the interface declared here (names, types, struct layout) is that of
 * libspecbleach - A spectral processing library, Copyright 2022 Luciano Dato
 * <lucianodato@gmail.com>, distributed under the GNU Lesser General Public License,
 * version 2.1 or (at your option) any later version, WITHOUT ANY WARRANTY; see
 * <https://www.gnu.org/licenses/old-licenses/lgpl-2.1.html>.  This header restates that
 * interface so that programs written against libspecbleach link unchanged; it is
 * distributed under the same licence.
 */
#ifndef SPECBLEACH_2D_DENOISER_H_INCLUDED
#define SPECBLEACH_2D_DENOISER_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

#include <stdbool.h>
#include <stdint.h>

typedef void* SpectralBleachHandle;

/* Replaces: reference include/specbleach_2d_denoiser.h:37-97 */
typedef struct SpectralBleach2DDenoiserParameters {
  int learn_noise;              /* >0: capture all four noise profiles         */
  bool residual_listen;         /* output = input - denoised                   */
  float reduction_amount;       /* dB, 0..40                                   */
  float smoothing_factor;       /* NLM smoothing strength h (typ. 0.5..3)      */
  float whitening_factor;       /* %, 0..100                                   */
  int adaptive_noise;           /* !=0: time-recursive noise tracking          */
  int noise_estimation_method;  /* 0 SPP-MMSE, 1 Brandt, 2 Martin              */
  float nlm_masking_protection; /* 0..1, psychoacoustic veto depth             */
  float suppression_strength;   /* %, Berouti oversubtraction strength         */
  float aggressiveness;         /* -1 (median) .. 0 (mean) .. +1 (max) profile */
  float tonal_reduction;        /* dB, extra reduction at tonal noise bins     */
} SpectralBleach2DDenoiserParameters;

/* reference specbleach_2d_denoiser.h:107 / src/processors/specbleach_2d_denoiser.c:40-80.
 * NULL if sample_rate == 0 or frame_size (ms) <= 0. */
SpectralBleachHandle specbleach_2d_initialize(uint32_t sample_rate, float frame_size);

/* reference specbleach_2d_denoiser.h:113 */
void specbleach_2d_free(SpectralBleachHandle instance);

/* reference specbleach_2d_denoiser.h:119 / specbleach_2d_denoiser.c:223-255 */
bool specbleach_2d_load_parameters(SpectralBleachHandle instance,
                                   SpectralBleach2DDenoiserParameters parameters);

/* reference specbleach_2d_denoiser.h:126 / specbleach_2d_denoiser.c:120-133 */
bool specbleach_2d_process(SpectralBleachHandle instance, uint32_t number_of_samples,
                           const float* input, float* output);

/* reference specbleach_2d_denoiser.h:134 / specbleach_2d_denoiser.c:102-118:
 * frame + 4*(fft_size/4) samples (the value the reference REPORTS). */
uint32_t specbleach_2d_get_latency(SpectralBleachHandle instance);

/* reference specbleach_2d_denoiser.h:139: fft_size (sic) for the 2D denoiser */
uint32_t specbleach_2d_get_noise_profile_size(SpectralBleachHandle instance);

/* reference specbleach_2d_denoiser.h:144 */
bool specbleach_2d_load_noise_profile_for_mode(SpectralBleachHandle instance,
                                               const float* restored_profile,
                                               uint32_t profile_size,
                                               uint32_t block_count, int mode);

/* reference specbleach_2d_denoiser.h:152 */
bool specbleach_2d_reset_noise_profile(SpectralBleachHandle instance);

/* reference specbleach_2d_denoiser.h:158 */
uint32_t specbleach_2d_get_noise_profile_block_count_for_mode(
    SpectralBleachHandle instance, int mode);

/* reference specbleach_2d_denoiser.h:164 */
float* specbleach_2d_get_noise_profile_for_mode(SpectralBleachHandle instance,
                                                int mode);

/* reference specbleach_2d_denoiser.h:170 */
bool specbleach_2d_noise_profile_available_for_mode(SpectralBleachHandle instance,
                                                    int mode);

#ifdef __cplusplus
}
#endif
#endif /* SPECBLEACH_2D_DENOISER_H_INCLUDED */
