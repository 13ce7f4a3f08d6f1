// specbleach-b200: shared declarations for the CUDA engine behind the C ABI.
//
// Naming follows the reference's domain: frames, bins, hops, (Bark) bands,
// noise profiles, SNR spectrogram.  Reference citations are file:line relative
// to the reference tree (lucianodato/libspecbleach v0.3.0).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace sb {

// ---------------------------------------------------------------------------
// Compile-time constants of the reference (src/shared/configurations.h)
// ---------------------------------------------------------------------------
constexpr float kSpectralEps = 1e-12f;          // :50  SPECTRAL_EPSILON
constexpr float kDbFsToSpl = 96.0f;             // :59
constexpr float kPowerLawExp = 0.6f;            // :60
constexpr float kAdditivityExp = 0.4f;          // :64  PEAQ exponent
constexpr float kVetoSmoothing = 0.5f;          // :90
constexpr float kVetoNmrRange = 20.0f;          // :91
constexpr float kAlphaMax = 4.0f;               // :101
constexpr float kAlphaMaxTonal = 10.0f;         // :102
constexpr float kAlphaMin = 1.0f;               // :103
constexpr float kSnrLowDb = -5.0f;              // :106
constexpr float kSnrHighDb = 20.0f;             // :107
constexpr float kSilenceThreshold = 1e-10f;     // :110
constexpr int kMartinSubwinCount = 8;           // :113
constexpr int kMartinSubwinLen = 12;            // :114
constexpr float kMartinBias = 1.5f;             // :115
constexpr float kMartinAlpha = 0.8f;            // :116
constexpr float kSppXi = 31.62f;                // :121
constexpr float kSppAlphaPow = 0.8f;            // :122
constexpr float kSppSmooth = 0.9f;              // :123
constexpr float kSppCurrent = 0.1f;             // :124
constexpr float kSppCap = 0.99f;                // :125
constexpr float kPeakThreshold = 1.41f;         // :142
constexpr float kMinPeakProminence = 1e-7f;     // :143
constexpr int kTonalWindow = 51;                // :145
constexpr int kMinWindowsAveraged = 5;          // :155
constexpr int kMedianFrames = 25;               // :156
constexpr float kInterpThreshold = 1e-9f;       // :157
constexpr float kNoiseSmoothing = 0.5f;         // :158/159
constexpr int kNlmPatch = 8;                    // :162
constexpr int kNlmPaste = 8;                    // :163
constexpr int kNlmRangeFreq = 8;                // :164
constexpr int kNlmPast = 16;                    // :165
constexpr int kNlmFuture = 4;                   // :166
constexpr int kNlmRing = kNlmPast + kNlmFuture + 1;  // 21 (spectral_2d_denoiser.c:191)
constexpr float kNlmMinWeight = 1e-10f;         // :168
constexpr float kNlmNoiseFloorMin = 1e-9f;      // :169
constexpr float kNlmFreqScale = 3.0f;           // :170
constexpr float kNlmThrMul = 4.0f;              // :171
constexpr int kOverlap = 4;                     // :182 / :204
constexpr int kZeroPad = 800;                   // :188 / :210
constexpr int kMaxBands = 25;                   // critical_bands.c:25-28
constexpr int kBandPitch = 32;                  // pitch of per-band arrays
constexpr int kBandStateRows = 5;               // Handle::d_band_state rows
constexpr int kMaxBandTasks = 96;               // band-sum work units per frame

constexpr int kDelayFrames = kNlmFuture;        // 2D look-ahead (nlm_filter.c:447-452)
constexpr int kSnrHist = kNlmRing - 1;          // 20 SNR frames carried between calls

// ---------------------------------------------------------------------------
// Geometry: everything that depends only on (sample_rate, frame_ms)
// ---------------------------------------------------------------------------
struct GeoK {            // POD copied into kernel parameters (constant bank)
  int sr, frame, N, hop, K, Kp, cp, M;   // M = N/2 complex points
  int frame_p;                           // synthesis row pitch
  int nfac;
  int fac[14];
  // per Stockham pass (host-computed: the kernels never divide): butterflies' stride l = len/r,
  // s = product of the previous radices, twiddle step M/len, nbf = M/r and ceil(2^32/s), nbf
  int pass_l[14], pass_s[14], pass_tstep[14], pass_nbf[14];
  unsigned pass_smagic[14], pass_nbfmagic[14];
  int fft_threads;            // CTA size of the FFT kernels (fill_passes: fewest idle warp rounds)
  // Length the Stockham passes above work on: M, or -- when M has a large prime factor (N = 3006 =
  // 2 3^2 167 at 44.1 kHz / 50 ms; the rule is build_fft_plan's work estimate) -- the power of two L >= 2M - 1 of Bluestein's chirp-z form,
  // in which the M-point DFT is a cyclic convolution of length L: two L-point FFTs and three
  // pointwise products instead of an O(r) sum per output of the radix-r pass.
  int fft_len;                // M or L
  int bl_L;                   // 0: plain mixed radix
  const float2* bl_w;         // [M] chirp exp(-i pi n^2 / M)
  const float2* bl_B;         // [L] FFT_L of the wrapped conjugate chirp, divided by L
  int nb;                                // number of Bark bands (<= 25)
  int band_start[kMaxBands + 1];         // band b = [band_start[b], band_start[b+1])
  // Bark bands differ wildly in width (3 bins ... two thirds of the spectrum): per-frame band
  // sums are cut into tasks of <= task_bins bins, dealt round-robin to warps and added per band
  // in task order (fixed order: deterministic)
  int ntask;
  short task_band[kMaxBandTasks];
  int task_k[kMaxBandTasks + 1];         // task t covers bins [task_k[t], min(task_k[t]+task_bins, band end))
  short band_task0[kMaxBands + 1];       // first task of band b
  float centre[kMaxBands];               // masking_veto.c:103-110
  float fwd_pow[kMaxBands];              // forward_decay[b]^0.6 (masking_estimator.c:148-157)
  float bwd;                             // backward_decay (masking_estimator.c:160)
  float spl_ref;                         // absolute_hearing_thresholds.c:118-139
  float scale;                           // stft_windows.c:103-118
  const float* win;                      // [frame] periodic Hann
  const float2* tw;                      // [fft_len] exp(-2 pi i k / fft_len)
  const float2* twn;                     // [M+1] exp(-2 pi i k / N)
  const double2* twg;                    // per generic-radix pass: r roots of unity in double
  int twg_off[14];                       // offset of pass f inside twg (-1: radix 2/4 pass)
  const float* abs_thr_spl;              // [Kp] absolute threshold, dB SPL
  const float* abs_thr_lin;              // [Kp] the same mapped back to power
  const int* cband;                      // [Kp] interpolation band of a bin
  const float* veto_t;                   // [Kp] position of the bin between the centres of cband, cband + 1, clamped to [0, 1]
  const int* band_of_bin;                // [Kp] Bark band containing a bin
};

struct Geometry {
  GeoK k;
  uint32_t sample_rate;
  float frame_ms;
  std::vector<void*> dev_allocs;
};

// Cached, shared by all handles on the current device with the same geometry.
const Geometry* get_geometry(uint32_t sample_rate, float frame_ms);

// ---------------------------------------------------------------------------
// Lane: one CUDA stream plus the scratch spectrograms of one sub-call
// ---------------------------------------------------------------------------
struct Lane {
  cudaStream_t stream = nullptr;
  int cap_frames = 0;     // frames per sub-call the scratch is sized for
  int cap_Kp = 0;
  int cap_frame_p = 0;
  size_t cap_samples = 0;
  float* xbuf = nullptr;       // [frame-hop+carry + n] input samples
  float* ybuf = nullptr;       // [n] output samples (host API staging)
  float* Xre = nullptr;        // [4+F][Kp]
  float* Xim = nullptr;
  float* P = nullptr;          // [F][Kp] power spectrum
  float* noise = nullptr;      // [4+F][Kp]
  float* S = nullptr;          // [20+F][Kp] SNR spectrogram
  float* Shat = nullptr;       // [F][Kp] NLM output
  float* Mg = nullptr;         // [F][Kp] smoothed magnitude / 1D smoothed power
  float* alpha = nullptr;      // [F][Kp]
  float* stable = nullptr;     // [F][Kp]
  float* wt = nullptr;         // [F][Kp] whitening weights (or 1 row)
  float* mask = nullptr;       // [F][Kp] tonal mask (or 1 row)
  float* Yre = nullptr;        // [F][Kp]
  float* Yim = nullptr;
  float* synth = nullptr;      // [F][frame_p]
  float* band_a = nullptr;     // [F][32]
  float* band_T = nullptr;     // [F][32]
  float* band_p = nullptr;     // [F][32]
  float* frame_flag = nullptr; // [F] frame energy (adaptive estimators)
  // host I/O pipeline: H2D of sub-call k+1 and D2H of sub-call k-1 overlap the kernels of
  // sub-call k (two staging buffers each way, dedicated copy streams, events)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  float* in_stage[2] = {nullptr, nullptr};   // [samples] incoming PCM
  float* out_stage[2] = {nullptr, nullptr};  // [samples] outgoing PCM
  cudaEvent_t ev_in[2] = {nullptr, nullptr};    // H2D into in_stage[b] finished
  cudaEvent_t ev_used[2] = {nullptr, nullptr};  // in_stage[b] consumed by the compute stream
  cudaEvent_t ev_out[2] = {nullptr, nullptr};   // out_stage[b] written by the compute stream
  cudaEvent_t ev_done[2] = {nullptr, nullptr};  // D2H from out_stage[b] finished
  unsigned io_seq = 0;                          // host sub-calls enqueued on this lane
  // page-locked bounce buffers for callers that pass pageable memory (allocated on first use)
  float* pin_in[2] = {nullptr, nullptr};
  float* pin_out[2] = {nullptr, nullptr};
  size_t pin_cap = 0;
  std::vector<cudaEvent_t> ev; // profiling event pairs (NLM)
};

// ---------------------------------------------------------------------------
// Engine-internal parameters (already converted: specbleach_2d_denoiser.c:234-248)
// ---------------------------------------------------------------------------
struct Params {
  int learn_noise = 0;
  bool residual_listen = false;
  float reduction = 1.0f;        // linear gain
  float smoothing = 0.0f;        // 1D: remapped release factor; 2D: NLM h
  float whitening = 0.0f;        // 0..1
  int adaptive = 0;
  int method = 0;
  float depth = 0.0f;            // masking depth / nlm_masking_protection
  float strength = 0.0f;         // 0..1
  float aggressiveness = 0.0f;
  float tonal = 1.0f;            // linear gain
};

enum Method { kSpp = 0, kBrandt = 1, kMartin = 2 };

struct Handle {
  bool two_d = true;
  const Geometry* geo = nullptr;
  int device = 0;
  Params prm;
  bool params_loaded = false;

  // --- noise profile, host mirror (noise_profile.c:27-32) ---
  uint32_t profile_size = 0;            // 2D: N (sic), 1D: K
  std::vector<float> prof_host[4];
  uint32_t block_count[4] = {0, 0, 0, 0};
  bool available[4] = {false, false, false, false};
  bool host_dirty = true;               // host mirror newer than device copy
  bool profile_learned = false;         // any median bin > 0 (tonal_reducer.c:62-69)
  uint64_t noise_version = 1;           // bumped whenever the manual noise vector may change

  // --- device state ---
  float* d_prof = nullptr;        // [4][Kp] mean, median, max, min
  float* d_med_ring = nullptr;    // [25][Kp]
  int med_write = 0;
  float* d_hist = nullptr;        // [frame] input FIFO tail
  float* d_tail[2] = {nullptr, nullptr};  // [frame] OLA accumulator, ping-pong
  int tail_cur = 0;
  uint32_t carry = 0;             // samples since the last frame boundary
  bool was_learning = false;
  bool profile_dl_pending = false; // host mirror refreshed by an in-flight copy
  // 2D delay line + NLM ring
  float* d_X_hist_re = nullptr;   // [4][Kp]
  float* d_X_hist_im = nullptr;
  float* d_noise_hist = nullptr;  // [4][Kp]
  uint64_t noise_hist_version[4] = {0, 0, 0, 0};
  float* d_S_hist = nullptr;      // [20][Kp]
  uint64_t dn_frames = 0;         // denoise-mode frames seen (ring fill)
  // veto / smoothing recursions
  float* d_stable = nullptr;      // [Kp]
  float* d_sp_prev = nullptr;     // [Kp] 1D release smoother
  float* d_band_state = nullptr;  // [5][32]: u = T_prev^0.6, protection memory; N4: u of the suppression
                                  // engine's masking estimator, transient detector energies, its flag
  int gain_type = 0, suppression_type = 0, smoothing_type = 1;   // N4 strategy selectors (GainArgs)
  // static-noise caches: tonal mask and whitening weights of the manual profile
  float* d_aux_mask = nullptr;    // [Kp]
  float* d_aux_wt = nullptr;      // [Kp]
  uint64_t aux_mask_version = 0, aux_wt_version = 0;
  float aux_wt_whitening = -1.0f;
  // noise for the current call
  float* d_noise_cur = nullptr;   // [Kp]
  float* d_floor = nullptr;       // [Kp] manual floor (adaptive mode)
  // adaptive estimator
  int est_method = -1;            // -1: none created yet
  bool est_first = true;
  float nlm_h = 1.0f;             // NLM_DEFAULT_H_PARAMETER until smoothing_factor > 0 is loaded
  int last_adaptive_state = 0;
  float aggressiveness_state = 0.0f;
  float* d_est = nullptr;         // estimator state, layout depends on method
  size_t est_floats = 0;
  uint32_t martin_frame_count = 0, martin_subwin = 0;
};

// ---------------------------------------------------------------------------
// Error plumbing and counters
// ---------------------------------------------------------------------------
void set_error(const char* what, cudaError_t e, const char* file, int line);
const char* last_error();
extern uint64_t g_launches;     // kernels launched by this library

// Every device allocation of the library goes through these two.  With SB_GUARD=1 in the
// environment each allocation gets a 4 KB zone of 0xFF bytes (float NaNs) on both sides:
// an out-of-bounds store damages a zone (check_guards / specbleach_b200_debug_check_guards
// counts them), an out-of-bounds float load drags a NaN into the output.  compute-sanitizer
// is not available on the GPU boxes of this project; this is the stand-in.
cudaError_t dev_malloc(void** p, size_t bytes);
void dev_free(void* p);
int check_guards();             // allocations (live, or freed since the last call) with a damaged zone; -1 on error
template <typename T>
static inline cudaError_t dev_malloc(T** p, size_t bytes) { return dev_malloc((void**)p, bytes); }
extern bool g_profile;          // record NLM events

#define SB_CUDA(expr)                                              \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) {                                       \
      ::sb::set_error(#expr, _e, __FILE__, __LINE__);              \
      return false;                                                \
    }                                                              \
  } while (0)

#define SB_LAUNCH_CHECK()                                          \
  do {                                                             \
    ::sb::g_launches++;                                            \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) {                                       \
      ::sb::set_error("kernel launch", _e, __FILE__, __LINE__);    \
      return false;                                                \
    }                                                              \
  } while (0)

// ---------------------------------------------------------------------------
// Per-kernel-class device timing (profile mode): CUDA events on the launching
// stream around every launch, folded into per-class totals by stats_collect().
// ---------------------------------------------------------------------------
enum KClass {
  kcForward = 0, kcInverse, kcOla, kcNlm, kcNlmTail, kcSnr, kcGainPre, kcBinScan, kcBand,
  kcTScan, kcNmr, kcPScan, kcFinal, kcTonal, kcWhiten, kcLearn, kcInterp, kcAdaptive, kcMisc,
  kcCount
};
const char* kclass_name(int c);
void ktimer_begin(int cls, cudaStream_t st);
void ktimer_end(int cls, cudaStream_t st);
struct KTimer {
  int cls;
  cudaStream_t st;
  KTimer(int c, cudaStream_t s) : cls(c), st(s) { if (g_profile) ktimer_begin(cls, st); }
  ~KTimer() { if (g_profile) ktimer_end(cls, st); }
};

// ---------------------------------------------------------------------------
// Kernel launchers (one translation unit per kernel family)
// ---------------------------------------------------------------------------
// sb_stft.cu
bool launch_forward(const GeoK& g, const float* xbuf, int frames, float* Xre, float* Xim,
                    float* P, cudaStream_t st);
bool launch_inverse(const GeoK& g, const float* Yre, const float* Yim, int frames,
                    float* synth, cudaStream_t st);
bool launch_ola(const GeoK& g, const float* synth, int frames, uint32_t carry, uint32_t n,
                const float* tail_in, float* tail_out, float* y, cudaStream_t st);

// sb_nlm.cu
bool launch_nlm(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                float* Shat, cudaStream_t st);

// sb_nlm3.cu (full paste blocks only; launch_nlm adds the tail kernel)
bool launch_nlm3(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                 float* Shat, cudaStream_t st);
int nlm3_chunk_targets(int K);   // targets per NLM chunk: whole waves of CTAs on this device

// sb_gain.cu
struct GainArgs {
  int frames;        // frames in this sub-call (M)
  int first_active;  // first frame index with gains (2D: ring ready), 0 for 1D
  bool two_d;
  float strength, tonal, reduction, whitening, depth, smoothing;
  bool residual;
  // N4: strategy branches the shipped reference build never selects (enum values of
  // gain_calculator.h:27-31, suppression_engine.h:39-46, spectral_smoother.h:28-31)
  int gain_type = 0;         // 0 Wiener, 1 gates, 2 generalized spectral subtraction
  int suppression_type = 0;  // 0 Berouti per bin, 1 global SNR, 2 critical-band SNR, 3 masking thresholds, 4 none
  int smoothing_type = 1;    // 1D: 1 fixed, 2 transient aware
  int mask_stride;   // 0: one tonal mask row shared by all frames, else Kp
  int wt_stride;     // 0: one whitening-weight row shared by all frames, else Kp
  const float* mask; // tonal mask rows (indexed by frame index * mask_stride)
  const float* wt;   // whitening weight rows
};
bool launch_morph_profile(const GeoK& g, const float* prof, float aggressiveness, float* out,
                          cudaStream_t st);
bool launch_fill_rows(const GeoK& g, const float* row, float* dst, int rows, cudaStream_t st);
bool launch_snr(const GeoK& g, const float* P, const float* noise, float* S, int frames,
                cudaStream_t st);
bool launch_tonal_mask(const GeoK& g, const float* det, int det_stride, int rows, float* mask,
                       cudaStream_t st);
bool launch_whitening(const GeoK& g, const float* noise, int noise_stride, int rows,
                      float whitening, float* wt, cudaStream_t st);
bool launch_gain_chain(const GeoK& g, const GainArgs& a, Lane& L, Handle& h, cudaStream_t st);

// sb_learn.cu
bool launch_learn(const GeoK& g, const float* P, int frames, uint32_t count0, float* prof,
                  float* med_ring, int med_write, cudaStream_t st);
bool launch_interp_smooth(const GeoK& g, float* rows, int row_stride, int nrows,
                          unsigned enable_mask, float threshold, float factor,
                          const float* floorv, cudaStream_t st);

// sb_adaptive.cu
size_t estimator_floats(const GeoK& g, int method);
bool launch_estimator_seed(const GeoK& g, int method, const float* floorv, float* est,
                           cudaStream_t st);
// idx_scratch: >= frames + 1 ints of lane scratch (Brandt: compacted non-silent frames);
// fresh_seed: launch_estimator_seed ran for this very call.
bool launch_adaptive(const GeoK& g, int method, const float* P, int frames, const float* floorv,
                     float* est, float* energy, int* idx_scratch, bool fresh_seed,
                     float* noise_rows, cudaStream_t st);

// sb_brandt.cu
void set_error_msg(const char* msg);   // sb_engine.cu
size_t brandt_estimator_floats(const GeoK& g);
int brandt_history(const GeoK& g);
bool brandt_supported(const GeoK& g);
bool launch_brandt_seed(const GeoK& g, const float* floorv, float* est, cudaStream_t st);
bool launch_brandt(const GeoK& g, const float* P, int frames, const float* floorv, float* est,
                   const float* energy, int* idx, bool fresh_seed, float* noise_rows,
                   cudaStream_t st);

}  // namespace sb
