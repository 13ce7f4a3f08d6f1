/* TEST INFRASTRUCTURE ONLY -- a build VARIANT of the reference's 2D processor: see
 * spectral_denoiser_variant.c.  Selectors: GAIN_ESTIMATION_TYPE_2D (configurations.h:218) and the
 * SuppressionType literal in spectral_2d_denoiser_run (spectral_2d_denoiser.c:397-400); the 2D
 * processor has no temporal smoother. */
#include "shared/configurations.h"
#include "shared/denoiser_logic/processing/gain_calculator.h"
#include "shared/denoiser_logic/processing/suppression_engine.h"

#ifdef SB_VARIANT_GAIN
#undef GAIN_ESTIMATION_TYPE_2D
#define GAIN_ESTIMATION_TYPE_2D SB_VARIANT_GAIN
#endif
#ifdef SB_VARIANT_SUPPRESSION
#define SUPPRESSION_BEROUTI_PER_BIN SB_VARIANT_SUPPRESSION
#endif

#include "processors/denoiser2d/spectral_2d_denoiser.c"
