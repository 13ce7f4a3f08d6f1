/*
 * specbleach-b200 -- EXTRA entry points that the reference does not have.
 *
 * The 22 drop-in symbols live in specbleach_denoiser.h and
 * specbleach_2d_denoiser.h and are untouched.  Everything here is additive
 * ("Allowed extension"): device selection, pinned host
 * buffers, device-resident and batched processing, and instrumentation for
 * bench.py / the parity tests.  Plain C ABI: pointers and sizes only.
 */
#ifndef SPECBLEACH_B200_H_INCLUDED
#define SPECBLEACH_B200_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef void* SpectralBleachHandle;

/* Number of CUDA devices visible to the process (0 if none / no driver). */
int specbleach_b200_device_count(void);

/* Selects the device used by handles created afterwards on this thread
 * (one process per GPU is the intended deployment). Returns false on error. */
bool specbleach_b200_set_device(int device);

/* Last error message of this library (empty string if none). */
const char* specbleach_b200_last_error(void);

/* Page-locked host buffers.  specbleach[_2d]_process and the batch call accept any host
 * memory: page-locked buffers (these, or cudaHostAlloc / cudaHostRegister'ed memory) are
 * DMA-copied directly under the kernels of the neighbouring sub-calls; pageable buffers
 * (malloc, the stack) of large calls are staged by the calling thread through two page-locked
 * bounce buffers of the lane, piece by piece, also under the kernels. */
void* specbleach_b200_host_alloc(size_t bytes);
void specbleach_b200_host_free(void* p);

/* Engine tuning (call before the first handle is created):
 * frames per internal sub-call (0 = automatic, the default: a whole number of NLM
 * chunks whose grids fill whole waves of SMs, 6 x 2176 = 13056 frames at
 * 48 kHz / 50 ms, ~1.5 GB of scratch per lane) and number of lanes, i.e. CUDA
 * streams with private scratch spectrograms (default 4, allocated on first use). */
bool specbleach_b200_configure(uint32_t chunk_frames, uint32_t lanes);

/* Same contract as specbleach[_2d]_process, but `input`/`output` are DEVICE
 * pointers on the handle's device; no host copies, returns after the work has
 * completed.  Works for handles of either denoiser. */
bool specbleach_b200_process_device(SpectralBleachHandle instance, uint32_t number_of_samples,
                                    const float* d_input, float* d_output);

/* Processes `count` independent handles (streams/channels), equivalent to
 * calling specbleach[_2d]_process on each, with host<->device copies and
 * kernels of different handles overlapped across lanes.  inputs[i]/outputs[i]
 * are HOST buffers of number_of_samples[i] floats (may alias each other
 * pairwise). Returns false if any handle failed. */
bool specbleach_b200_process_batch(SpectralBleachHandle* instances, uint32_t count,
                                   const uint32_t* number_of_samples,
                                   const float* const* inputs, float* const* outputs);

/* Device-resident variant of the batch call. */
bool specbleach_b200_process_batch_device(SpectralBleachHandle* instances, uint32_t count,
                                          const uint32_t* number_of_samples,
                                          const float* const* d_inputs,
                                          float* const* d_outputs);

/* Time-tiled processing of ONE long stream: same contract as
 * specbleach[_2d]_process on HOST buffers, but the `number_of_samples` are cut into `tiles`
 * hop-aligned time tiles that run concurrently on the engine's lanes and, when `devices` names
 * more GPUs of this process, on those as well (no data-path collective: every tile reads its
 * slice of `input` and writes its slice of `output`).  Tile 0 continues the handle; every other
 * tile runs on a fresh internal handle that received the finalised noise profile and the
 * parameters and starts `warmup_frames` hops early (its warm-up output is dropped); afterwards
 * the handle continues from the state of the last tile.  The handle must be in denoise mode.
 * Manual-profile mode: with warmup_frames >= 512 the output equals specbleach[_2d]_process bit
 * for bit in the tests (the carried recursions decay by >= 0.09 per frame).  Adaptive estimators
 * are only approximated (long memories): refused unless `flags` has
 * SPECBLEACH_B200_TILED_ALLOW_ADAPTIVE.  Pageable buffers are page-locked for the call. */
enum { SPECBLEACH_B200_TILED_ALLOW_ADAPTIVE = 1 };
bool specbleach_b200_process_tiled(SpectralBleachHandle instance, uint32_t number_of_samples,
                                   const float* input, float* output, uint32_t tiles,
                                   uint32_t warmup_frames, const int* devices,
                                   uint32_t device_count, uint32_t flags);

/* The same for several handles at once (the channels of one file): all tiles of all handles are
 * enqueued before anything is waited for, so tiles of different channels overlap on every GPU. */
bool specbleach_b200_process_tiled_batch(SpectralBleachHandle* instances, uint32_t count,
                                         const uint32_t* number_of_samples, const float* const* inputs,
                                         float* const* outputs, uint32_t tiles, uint32_t warmup_frames,
                                         const int* devices, uint32_t device_count, uint32_t flags);

/* Strategy branches of the reference that its shipped build never selects.  The
 * reference fixes them at compile time: GAIN_ESTIMATION_TYPE_1D / _2D (configurations.h:195,218),
 * the SuppressionType literal in spectral_denoiser.c:294-297 / spectral_2d_denoiser.c:397-400 and
 * the TimeSmoothingType literal in spectral_denoiser.c:83.  Here they are per-handle settings
 * with the reference's enum values; the defaults are the shipped build's (WIENER, BEROUTI_PER_BIN,
 * FIXED), so a handle that never calls this is the drop-in path.  Parity is held against
 * rebuilds of the reference with the selector pinned (oracle/variants/).
 *   gain_type        0 WIENER, 1 GATES, 2 GENERALIZED_SPECTRALSUBTRACTION   (gain_calculator.c:71-139)
 *   suppression_type 0 BEROUTI_PER_BIN, 1 GLOBAL_SNR, 2 CRITICAL_BANDS_SNR, 3 MASKING_THRESHOLDS,
 *                    4 NONE                                               (suppression_engine.c:135-275)
 *   smoothing_type   1 FIXED, 2 TRANSIENT_AWARE (1D denoiser only; ignored by the 2D denoiser,
 *                    which has no temporal smoother)    (spectral_smoother.c:141-181, transient_detector.c:66-111)
 * May be changed between any two process() calls, like the parameters. */
bool specbleach_b200_set_strategy(SpectralBleachHandle instance, int gain_type, int suppression_type,
                                  int smoothing_type);

/* Noise-profile persistence ------------------------------------------------
 * The reference can only hand out / take in raw profile arrays
 * (specbleach_2d_denoiser.c:145-206: get_noise_profile_for_mode,
 * load_noise_profile_for_mode).  These two wrap them in a small checksummed file
 * ("SBNP" v1: sample rate, frame size, denoiser kind, and for each of the four
 * modes block count, availability and the array) so that one learned profile can
 * be shared by thousands of stream handles.  load fails (false, last_error set)
 * when the file was saved for another (sample_rate, frame_ms, 1D/2D). */
bool specbleach_b200_profile_save(SpectralBleachHandle instance, const char* path);
bool specbleach_b200_profile_load(SpectralBleachHandle instance, const char* path);

/* Streaming-state snapshot: parameters, profiles, FIFO / overlap-add tails,
 * 2D delay line and NLM ring, veto recursions, estimator state -- everything a
 * handle carries between process() calls -- as one opaque,
 * checksummed host blob.  A handle of the same geometry restored from it (on any
 * GPU / process) continues the stream bit-exactly. */
size_t specbleach_b200_state_size(SpectralBleachHandle instance);
bool specbleach_b200_state_save(SpectralBleachHandle instance, void* blob, size_t bytes);
bool specbleach_b200_state_load(SpectralBleachHandle instance, const void* blob, size_t bytes);

/* Interleaved multichannel PCM through one handle per channel: replaces the
 * host-side de-interleave / sf_readf_float / sf_writef_float loop of
 * examples/denoiser_demo.c:217-300.  `input` / `output` are HOST buffers of
 * frames * channels samples (may alias); conversion and (de)interleaving run on
 * the device, the channels as one batch. */
enum {
  SPECBLEACH_B200_PCM_F32 = 0, /* IEEE float, as is                           */
  SPECBLEACH_B200_PCM_S16 = 1, /* in: x / 2^15; out: round(x * (2^15-1)), sat. */
  SPECBLEACH_B200_PCM_S24 = 2, /* packed 3-byte little endian                  */
  SPECBLEACH_B200_PCM_S32 = 3
};
bool specbleach_b200_process_interleaved(SpectralBleachHandle* instances, uint32_t channels,
                                         uint32_t frames, int format, const void* input,
                                         void* output);

/* Instrumentation ---------------------------------------------------------- */
typedef struct SpecbleachB200Stats {
  uint64_t kernel_launches;    /* kernels launched by this library since reset      */
  uint64_t nlm_launches;       /* NLM main-kernel launches timed with CUDA events   */
  double nlm_ms;               /* summed device time of those launches (ms)         */
  uint64_t nlm_frame_blocks;   /* target frames x 8-bin paste blocks they processed */
  uint64_t frames;             /* STFT frames processed                             */
} SpecbleachB200Stats;

/* enable != 0: bracket every kernel launch with CUDA events on its stream. */
void specbleach_b200_profile(int enable);
void specbleach_b200_stats_reset(void);
/* Synchronises the device, folds pending event pairs into the totals. */
bool specbleach_b200_stats_get(SpecbleachB200Stats* out);

/* Per-kernel-class device time (profile mode): every launch of the library is
 * bracketed by CUDA events on its stream.  Classes are enumerated 0..count-1. */
int specbleach_b200_stats_classes(void);
const char* specbleach_b200_stats_class_name(int cls);
bool specbleach_b200_stats_class(int cls, double* ms, uint64_t* launches);

/* Geometry of a handle: frame, fft size, hop, bins (for tests and tools). */
bool specbleach_b200_get_geometry(SpectralBleachHandle instance, uint32_t* frame,
                                  uint32_t* fft_size, uint32_t* hop, uint32_t* bins);

/* FFMA micro-benchmark on the current device: sustained FP32 TFLOP/s (FMA = 2
 * FLOP).  It is the denominator of the NLM kernel's roofline, which is bound by
 * the FP32 pipe and not by HBM (MEASURED_PEAKS.json has no FP32 figure). */
double specbleach_b200_debug_fp32_peak_tflops(void);

/* With SB_GUARD=1 in the environment when the library is loaded, every device allocation
 * carries a 4 KB zone of 0xFF bytes on each side (an out-of-bounds float load reads a NaN, an
 * out-of-bounds store damages the zone).  Returns the number of allocations, live or freed since
 * the previous call, whose zones were written; -1 when guarding is off.  Synchronises all
 * devices.  SB_GUARD=selftest additionally stores one byte behind the first live allocation
 * before checking, so the call must then return >= 1 (tests/test_gpu_guard.py).
 * (compute-sanitizer's memcheck is the better tool where it can be run.) */
int specbleach_b200_debug_check_guards(void);

/* Single-kernel entry points used by the parity tests (host buffers). ------ */
/* NLM over a whole SNR map snr[frames][bins] (row-major, pitch = bins):
 * out[j] = filtered frame j-4 as the reference returns it after pushing frame
 * j (rows j < 20 are left untouched). */
bool specbleach_b200_debug_nlm(const float* snr, uint32_t frames, uint32_t bins, float h,
                               float* out);
/* Forward real FFT of `frames` rows of fft_size samples each -> FFTW
 * halfcomplex rows (r0..r_{N/2}, i_{N/2-1}..i_1), then back (unnormalised). */
bool specbleach_b200_debug_rfft(uint32_t fft_size, uint32_t frames, const float* in,
                                float* halfcomplex, float* roundtrip);

#ifdef __cplusplus
}
#endif
#endif /* SPECBLEACH_B200_H_INCLUDED */
