/* TEST INFRASTRUCTURE ONLY -- a build VARIANT of the reference's 1D processor.
 *
 * The reference selects its gain rule, its oversubtraction strategy and its temporal smoother at
 * compile time: GAIN_ESTIMATION_TYPE_1D (configurations.h:195), the SuppressionType literal in
 * spectral_denoiser_run (spectral_denoiser.c:294-297) and the TimeSmoothingType literal in
 * spectral_denoiser_initialize (:83).  The shipped build only ever takes WIENER /
 * SUPPRESSION_BEROUTI_PER_BIN / FIXED; the other branches are reachable by
 * rebuilding.  This file does exactly that without touching or copying the reference source:
 * it pins the three selectors through the preprocessor and then includes the reference's
 * translation unit from where it lies.  oracle/Makefile passes
 *   -DSB_VARIANT_GAIN=<GainCalculationType> -DSB_VARIANT_SUPPRESSION=<SuppressionType>
 *   -DSB_VARIANT_SMOOTHING=<TimeSmoothingType>
 * and links the result in place of spectral_denoiser.o into oracle/_ref/variants/<name>.so. */
#include "shared/configurations.h"
#include "shared/denoiser_logic/processing/gain_calculator.h"
#include "shared/denoiser_logic/processing/suppression_engine.h"
#include "shared/utils/spectral_smoother.h"

#ifdef SB_VARIANT_GAIN
#undef GAIN_ESTIMATION_TYPE_1D
#define GAIN_ESTIMATION_TYPE_1D SB_VARIANT_GAIN
#endif
#ifdef SB_VARIANT_SUPPRESSION
#define SUPPRESSION_BEROUTI_PER_BIN SB_VARIANT_SUPPRESSION
#endif
#ifdef SB_VARIANT_SMOOTHING
#define FIXED SB_VARIANT_SMOOTHING
#endif

#include "processors/denoiser/spectral_denoiser.c"
