// Engine: handle state, lanes (stream + scratch), and the per-call pipeline that
// turns the reference's sample-by-sample loop (stft_processor.c:126-175 calling
// spectral_2d_denoiser_run / spectral_denoiser_run once per hop) into batched
// kernel launches over all frames that complete inside a process() call.
//
// Equivalence with the streaming reference is kept by carrying exactly the
// state the reference carries: input FIFO tail, OLA
// accumulator, 4 delayed FFT/noise frames, 20 SNR frames, the veto/smoother
// recursions, the estimator state and the noise profiles.
#include "sb_engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <system_error>
#include <thread>

namespace sb {

uint64_t g_launches = 0;
bool g_profile = false;

static thread_local std::string t_err;
static std::string g_err;
static std::mutex g_err_mu;

void set_error(const char* what, cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  std::lock_guard<std::mutex> lock(g_err_mu);
  g_err = buf;
}
void set_error_msg(const char* msg) {
  std::lock_guard<std::mutex> lock(g_err_mu);
  g_err = msg;
}
const char* last_error() {
  std::lock_guard<std::mutex> lock(g_err_mu);
  t_err = g_err;
  return t_err.c_str();
}

// ---------------------------------------------------------------------------
// Per-kernel-class timing
// ---------------------------------------------------------------------------
struct KEvent { int cls; cudaEvent_t e0, e1; cudaStream_t st; };
static std::vector<KEvent> g_kevents;
static std::mutex g_kev_mu;
static const char* const kNames[kcCount] = {
    "forward_fft", "inverse_fft", "ola", "nlm", "nlm_tail", "snr", "gain_pre", "bin_scan", "band",
    "tscan", "nmr", "pscan", "final", "tonal_mask", "whitening", "learn", "interp_smooth",
    "adaptive", "misc"};
const char* kclass_name(int c) { return (c >= 0 && c < kcCount) ? kNames[c] : "?"; }
void ktimer_begin(int cls, cudaStream_t st) {
  KEvent k;
  k.cls = cls;
  k.st = st;
  if (cudaEventCreate(&k.e0) != cudaSuccess || cudaEventCreate(&k.e1) != cudaSuccess) return;
  cudaEventRecord(k.e0, st);
  std::lock_guard<std::mutex> lock(g_kev_mu);
  g_kevents.push_back(k);
}
void ktimer_end(int cls, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_kev_mu);
  for (size_t i = g_kevents.size(); i-- > 0;)
    if (g_kevents[i].cls == cls) { cudaEventRecord(g_kevents[i].e1, st); break; }
}

// ---------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------
static Context g_ctx;
Context& ctx() { return g_ctx; }

// The lanes (streams + scratch spectrograms) and the geometry cache are shared by all handles of
// the process.  Distinct handles may be driven from different host threads (handles
// are independent, the reference has no globals); every entry point that touches the device
// takes this lock, so concurrent callers are serialised instead of racing on a lane.
static std::recursive_mutex g_engine_mu;
std::recursive_mutex& engine_mutex() { return g_engine_mu; }

// ---- guarded allocator (sb_common.h) ----
namespace {
constexpr size_t kGuardBytes = 4096;
struct GuardRec { int device; size_t bytes; };
std::mutex g_guard_mu;
std::map<void*, GuardRec> g_guarded;      // user pointer -> record
int g_guard_damaged_freed = 0;
bool guard_enabled() {
  static const bool on = [] { const char* e = getenv("SB_GUARD"); return e && *e && *e != '0'; }();
  return on;
}
// true when both zones of the allocation still hold 0xFF
bool guard_intact(void* user, const GuardRec& r) {
  static thread_local std::vector<unsigned char> host(2 * kGuardBytes);
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(r.device);
  char* base = (char*)user - kGuardBytes;
  cudaError_t e = cudaMemcpy(host.data(), base, kGuardBytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess)
    e = cudaMemcpy(host.data() + kGuardBytes, (char*)user + r.bytes, kGuardBytes, cudaMemcpyDeviceToHost);
  cudaSetDevice(cur);
  if (e != cudaSuccess) return false;
  for (unsigned char b : host)
    if (b != 0xFF) return false;
  return true;
}
}  // namespace

cudaError_t dev_malloc(void** p, size_t bytes) {
  if (!guard_enabled()) return cudaMalloc(p, bytes);
  const size_t padded = (bytes + 255) / 256 * 256;   // the trailing zone starts right after the last byte
  char* base = nullptr;
  cudaError_t e = cudaMalloc((void**)&base, padded + 2 * kGuardBytes);
  if (e != cudaSuccess) return e;
  e = cudaMemset(base, 0xFF, padded + 2 * kGuardBytes);
  if (e != cudaSuccess) { cudaFree(base); return e; }
  int dev = 0;
  cudaGetDevice(&dev);
  *p = base + kGuardBytes;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guarded[*p] = GuardRec{dev, bytes};
  return cudaSuccess;
}

void dev_free(void* p) {
  if (!p) return;
  if (!guard_enabled()) { cudaFree(p); return; }
  GuardRec r{};
  {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guarded.find(p);
    if (it == g_guarded.end()) { cudaFree(p); return; }
    r = it->second;
    g_guarded.erase(it);
  }
  cudaDeviceSynchronize();
  if (!guard_intact(p, r)) {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guard_damaged_freed++;
  }
  cudaFree((char*)p - kGuardBytes);
}

int check_guards() {
  if (!guard_enabled()) return -1;
  std::map<void*, GuardRec> live;
  int damaged;
  {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    live = g_guarded;
    damaged = g_guard_damaged_freed;
    g_guard_damaged_freed = 0;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  for (int d = 0; d < ndev; d++) { cudaSetDevice(d); cudaDeviceSynchronize(); }
  cudaSetDevice(cur);
  const char* mode = getenv("SB_GUARD");
  if (mode && !strcmp(mode, "selftest") && !live.empty()) {
    // proves the detector: one byte stored right behind the first live allocation
    auto& kv = *live.begin();
    cudaSetDevice(kv.second.device);
    cudaMemset((char*)kv.first + kv.second.bytes, 0, 1);
    cudaSetDevice(cur);
  }
  for (auto& kv : live)
    if (!guard_intact(kv.first, kv.second)) damaged++;
  return damaged;
}

static bool dev_alloc(float** p, size_t floats) {
  SB_CUDA(dev_malloc((void**)p, floats * sizeof(float)));
  return true;
}
static bool dev_zalloc(float** p, size_t floats) {
  if (!dev_alloc(p, floats)) return false;
  SB_CUDA(cudaMemset(*p, 0, floats * sizeof(float)));
  return true;
}

static void lane_free(Lane& L) {
  float** bufs[] = {&L.xbuf, &L.ybuf, &L.Xre, &L.Xim, &L.P, &L.noise, &L.S, &L.Shat, &L.Mg,
                    &L.alpha, &L.stable, &L.wt, &L.mask, &L.Yre, &L.Yim, &L.synth,
                    &L.band_a, &L.band_T, &L.band_p, &L.frame_flag, &L.in_stage[0],
                    &L.in_stage[1], &L.out_stage[0], &L.out_stage[1]};
  for (float** b : bufs) {
    if (*b) dev_free(*b);
    *b = nullptr;
  }
  L.cap_frames = 0;
}

// Sizes the scratch of a lane for sub-calls of up to F frames of geometry g.
static bool lane_ensure(Lane& L, const GeoK& g, int F) {
  // Blocking streams on purpose: an event recorded on the legacy default stream
  // (what torch.cuda.Event uses) then brackets the work of every lane, while
  // lanes still overlap with each other.
  if (!L.stream) SB_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamDefault));
  if (!L.s_in) {
    SB_CUDA(cudaStreamCreateWithFlags(&L.s_in, cudaStreamDefault));
    SB_CUDA(cudaStreamCreateWithFlags(&L.s_out, cudaStreamDefault));
    for (int b = 0; b < 2; b++) {
      SB_CUDA(cudaEventCreateWithFlags(&L.ev_in[b], cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&L.ev_used[b], cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&L.ev_out[b], cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&L.ev_done[b], cudaEventDisableTiming));
    }
  }
  const size_t samples = (size_t)g.frame + (size_t)(F + 1) * g.hop;
  if (L.cap_frames >= F && L.cap_Kp >= g.Kp && L.cap_frame_p >= g.frame_p &&
      L.cap_samples >= samples)
    return true;
  SB_CUDA(cudaStreamSynchronize(L.stream));
  SB_CUDA(cudaStreamSynchronize(L.s_in));
  SB_CUDA(cudaStreamSynchronize(L.s_out));
  lane_free(L);
  L.io_seq = 0;
  const size_t Kp = g.Kp, R = (size_t)F;
  bool ok = dev_alloc(&L.xbuf, samples) && dev_alloc(&L.ybuf, samples) &&
            dev_alloc(&L.in_stage[0], samples) && dev_alloc(&L.in_stage[1], samples) &&
            dev_alloc(&L.out_stage[0], samples) && dev_alloc(&L.out_stage[1], samples) &&
            dev_alloc(&L.Xre, (R + 4) * Kp) && dev_alloc(&L.Xim, (R + 4) * Kp) &&
            dev_alloc(&L.P, R * Kp) && dev_alloc(&L.noise, (R + 4) * Kp) &&
            dev_alloc(&L.S, (R + kSnrHist) * Kp) && dev_alloc(&L.Shat, R * Kp) &&
            dev_alloc(&L.Mg, R * Kp) && dev_alloc(&L.alpha, R * Kp) &&
            dev_alloc(&L.stable, R * Kp) && dev_alloc(&L.wt, R * Kp) &&
            dev_alloc(&L.mask, R * Kp) && dev_alloc(&L.Yre, R * Kp) &&
            dev_alloc(&L.Yim, R * Kp) && dev_alloc(&L.synth, R * (size_t)g.frame_p) &&
            dev_alloc(&L.band_a, R * kBandPitch) && dev_alloc(&L.band_T, R * kBandPitch) &&
            dev_alloc(&L.band_p, R * kBandPitch) && dev_alloc(&L.frame_flag, R);
  if (!ok) {
    lane_free(L);
    return false;
  }
  L.cap_frames = F;
  L.cap_Kp = g.Kp;
  L.cap_frame_p = g.frame_p;
  L.cap_samples = samples;
  return true;
}

int sub_call_frames(const GeoK& g) {
  const int c = ctx().chunk_frames;
  if (c > 0) return c;
  // default: six full NLM chunks per sub-call (measured on configs[1]: 2 chunks 22.6k x
  // realtime, 4: 23.0k, 6: 23.05k, 8: 23.04k, 12: 22.8k), fewer when a lane's buffers
  // (~20 rows of Kp floats per frame) would pass 4 GB
  const int chunk = nlm3_chunk_targets(g.K);
  const long long max_frames = 4000000000ll / (80ll * (long long)g.Kp);
  int nchunks = (int)(max_frames / chunk);
  if (nchunks > 6) nchunks = 6;
  if (nchunks < 1) nchunks = 1;
  const int f = nchunks * chunk;
  return f < 32 ? 32 : f;
}

Lane* get_lane(int idx, const GeoK& g) {
  Context& c = ctx();
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;   // callers select the handle's device first
  std::vector<Lane>& lanes = c.lanes[dev];
  if ((int)lanes.size() < c.num_lanes) lanes.resize(c.num_lanes);
  Lane& L = lanes[idx % c.num_lanes];
  if (!lane_ensure(L, g, sub_call_frames(g))) return nullptr;
  return &L;
}

// ---------------------------------------------------------------------------
// Handle lifecycle
// ---------------------------------------------------------------------------
Handle* handle_create(bool two_d, uint32_t sample_rate, float frame_ms) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error_msg("specbleach-b200: no CUDA device available (this build has no CPU fallback)");
    return nullptr;
  }
  const Geometry* geo = get_geometry(sample_rate, frame_ms);
  if (!geo) return nullptr;
  const GeoK& g = geo->k;
  Handle* h = new Handle();
  h->two_d = two_d;
  h->geo = geo;
  cudaGetDevice(&h->device);
  h->profile_size = two_d ? (uint32_t)g.N : (uint32_t)g.K;   // specbleach_2d_denoiser.c:65 / specbleach_denoiser.c:66
  for (int m = 0; m < 4; m++) h->prof_host[m].assign(h->profile_size, 0.f);
  const size_t Kp = g.Kp;
  bool ok = dev_zalloc(&h->d_prof, 4 * Kp) && dev_zalloc(&h->d_med_ring, kMedianFrames * Kp) &&
            dev_zalloc(&h->d_hist, (size_t)g.frame + g.hop) &&
            dev_zalloc(&h->d_tail[0], g.frame) && dev_zalloc(&h->d_tail[1], g.frame) &&
            dev_zalloc(&h->d_stable, Kp) && dev_zalloc(&h->d_sp_prev, Kp) &&
            dev_zalloc(&h->d_band_state, kBandStateRows * kBandPitch) && dev_zalloc(&h->d_noise_cur, Kp) &&
            dev_zalloc(&h->d_floor, Kp) && dev_zalloc(&h->d_aux_mask, Kp) &&
            dev_zalloc(&h->d_aux_wt, Kp);
  if (ok && two_d) {
    ok = dev_zalloc(&h->d_X_hist_re, 4 * Kp) && dev_zalloc(&h->d_X_hist_im, 4 * Kp) &&
         dev_zalloc(&h->d_noise_hist, 4 * Kp) && dev_zalloc(&h->d_S_hist, kSnrHist * Kp);
  }
  if (!ok) {
    handle_destroy(h);
    return nullptr;
  }
  h->host_dirty = false;   // device and host profiles are both all-zero
  return h;
}

void handle_destroy(Handle* h) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  if (!h) return;
  DeviceGuard restore_device;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  float* bufs[] = {h->d_prof, h->d_med_ring, h->d_hist, h->d_tail[0], h->d_tail[1],
                   h->d_X_hist_re, h->d_X_hist_im, h->d_noise_hist, h->d_S_hist, h->d_stable,
                   h->d_sp_prev, h->d_band_state, h->d_noise_cur, h->d_floor, h->d_est,
                   h->d_aux_mask, h->d_aux_wt};
  for (float* b : bufs)
    if (b) dev_free(b);
  delete h;
}

static void refresh_profile_learned(Handle* h) {
  // tonal_reducer.c:62-69: "learned" == any median bin > 0 (first K entries)
  const int K = h->geo->k.K;
  bool learned = false;
  for (int k = 0; k < K && !learned; k++) learned = h->prof_host[1][k] > 0.0f;
  h->profile_learned = learned;
}

bool handle_load_parameters(Handle* h, const Params& p) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  // load_2d_reduction_parameters / load_reduction_parameters
  // (spectral_2d_denoiser.c:287-322, spectral_denoiser.c:236-256)
  if (p.adaptive) {
    int requested = p.method;
    if (requested != kSpp && requested != kBrandt && requested != kMartin) requested = kMartin;
    if (requested == kBrandt && !brandt_supported(h->geo->k)) {
      set_error_msg("specbleach-b200: Brandt history too long for this frame size");
      return false;
    }
    if (h->est_method != requested) {
      const GeoK& g = h->geo->k;
      DeviceGuard restore_device;           // the estimator state lives on the handle's GPU
      SB_CUDA(cudaSetDevice(h->device));
      SB_CUDA(cudaDeviceSynchronize());     // a lane may still read the old estimator
      if (h->d_est) dev_free(h->d_est);
      h->d_est = nullptr;
      h->est_floats = estimator_floats(g, requested);
      if (!dev_zalloc(&h->d_est, h->est_floats)) return false;
      const float one = 1.0f;   // is_first_frame = true
      SB_CUDA(cudaMemcpy(h->d_est, &one, sizeof(float), cudaMemcpyHostToDevice));
      h->est_method = requested;
      h->last_adaptive_state = 0;
    }
  }
  if (p.aggressiveness != h->prm.aggressiveness || p.adaptive != h->prm.adaptive ||
      !h->params_loaded)
    h->noise_version++;
  h->prm = p;
  h->aggressiveness_state = p.aggressiveness;
  if (h->two_d && p.smoothing > 0.0f) h->nlm_h = p.smoothing;   // spectral_2d_denoiser.c:317-319
  h->params_loaded = true;
  return true;
}

bool handle_load_profile(Handle* h, const float* prof, uint32_t size, uint32_t blocks, int mode) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  if (size != h->profile_size || mode < 1 || mode > 4) return false;
  memcpy(h->prof_host[mode - 1].data(), prof, size * sizeof(float));
  h->block_count[mode - 1] = blocks;
  h->available[mode - 1] = true;
  h->host_dirty = true;
  h->noise_version++;
  refresh_profile_learned(h);
  return true;
}

void handle_reset_profile(Handle* h) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  for (int m = 0; m < 4; m++) {
    std::fill(h->prof_host[m].begin(), h->prof_host[m].end(), 0.f);
    h->block_count[m] = 0;
    h->available[m] = false;
  }
  h->host_dirty = true;
  h->noise_version++;
  h->profile_learned = false;
}

// host mirror -> device rows (only the first K entries are ever used)
static bool upload_profiles(Handle* h, cudaStream_t st) {
  const GeoK& g = h->geo->k;
  for (int m = 0; m < 4; m++)
    SB_CUDA(cudaMemcpyAsync(h->d_prof + (size_t)m * g.Kp, h->prof_host[m].data(),
                            g.K * sizeof(float), cudaMemcpyHostToDevice, st));
  h->host_dirty = false;
  return true;
}
static bool download_profiles(Handle* h, cudaStream_t st) {
  const GeoK& g = h->geo->k;
  for (int m = 0; m < 4; m++)
    SB_CUDA(cudaMemcpyAsync(h->prof_host[m].data(), h->d_prof + (size_t)m * g.Kp,
                            g.K * sizeof(float), cudaMemcpyDeviceToHost, st));
  return true;
}

// ---------------------------------------------------------------------------
// One sub-call: `m` new samples are already at L.xbuf[hist_len ...]
// ---------------------------------------------------------------------------
static bool d2d(void* dst, const void* src, size_t floats, cudaStream_t st) {
  if (floats == 0) return true;
  SB_CUDA(cudaMemcpyAsync(dst, src, floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return true;
}

uint32_t hist_len(const Handle* h) {
  const GeoK& g = h->geo->k;
  return (uint32_t)(g.frame - g.hop) + h->carry;
}

bool sub_begin(Handle* h, Lane& L) {
  // the FIFO tail goes in front of the new samples (stft_buffer.c:88-103)
  return d2d(L.xbuf, h->d_hist, hist_len(h), L.stream);
}

// learn -> denoise transition (denoiser_profile_core.c:42-48): every available profile is refined
// once (gap interpolation + smoothing, noise_estimator.c:143-162) and the host mirror follows.
static bool finalize_profiles(Handle* h, cudaStream_t st) {
  if (!h->was_learning) return true;
  const GeoK& g = h->geo->k;
  if (h->host_dirty && !upload_profiles(h, st)) return false;
  unsigned mask = 0;
  for (int m = 0; m < 4; m++)
    if (h->available[m]) mask |= 1u << m;
  if (!launch_interp_smooth(g, h->d_prof, g.Kp, 4, mask, kInterpThreshold, kNoiseSmoothing,
                            nullptr, st))
    return false;
  if (!download_profiles(h, st)) return false;
  SB_CUDA(cudaStreamSynchronize(st));
  refresh_profile_learned(h);
  h->was_learning = false;
  h->noise_version++;
  return true;
}

static bool denoise_frames(Handle* h, Lane& L, int M) {
  const GeoK& g = h->geo->k;
  cudaStream_t st = L.stream;
  const size_t Kp = g.Kp;
  const bool two_d = h->two_d;
  const Params& p = h->prm;
  const int lead = two_d ? kDelayFrames : 0;   // history rows in front of X / noise

  // -- learn -> denoise transition: refine every available profile once --
  if (!finalize_profiles(h, st)) return false;
  if (h->host_dirty && !upload_profiles(h, st)) return false;

  // -- spectra of the new frames (rows lead.. of the delayed buffers) --
  if (!launch_forward(g, L.xbuf, M, L.Xre + lead * Kp, L.Xim + lead * Kp, L.P, st)) return false;
  if (two_d) {
    if (!d2d(L.Xre, h->d_X_hist_re, 4 * Kp, st) || !d2d(L.Xim, h->d_X_hist_im, 4 * Kp, st) ||
        !d2d(L.noise, h->d_noise_hist, 4 * Kp, st) || !d2d(L.S, h->d_S_hist, kSnrHist * Kp, st))
      return false;
  }

  // -- noise rows for the new frames (denoiser_profile_core.c:53-102) --
  float* noise_new = L.noise + lead * Kp;
  const bool adaptive = p.adaptive && h->d_est;
  if (adaptive) {
    const bool state_changed = h->last_adaptive_state == 0;
    const bool mode_changed = fabsf(h->aggressiveness_state - p.aggressiveness) > 0.01f;
    const bool fresh_seed = state_changed || mode_changed;
    if (fresh_seed) {
      if (!launch_morph_profile(g, h->d_prof, p.aggressiveness, h->d_floor, st)) return false;
      if (!launch_estimator_seed(g, h->est_method, h->d_floor, h->d_est, st)) return false;
      h->last_adaptive_state = 1;
    }
    // L.band_a ([F][32] floats) is free until the gain chain: scratch for the frame index list
    if (!launch_adaptive(g, h->est_method, L.P, M, h->d_floor, h->d_est, L.frame_flag,
                         reinterpret_cast<int*>(L.band_a), fresh_seed, noise_new, st))
      return false;
    h->aggressiveness_state = p.aggressiveness;
  } else {
    h->last_adaptive_state = 0;
    if (!launch_morph_profile(g, h->d_prof, p.aggressiveness, h->d_noise_cur, st)) return false;
    if (!launch_fill_rows(g, h->d_noise_cur, noise_new, M, st)) return false;
  }

  // -- 2D: SNR spectrogram and NLM over every target whose ring is full --
  int first_active = 0;
  if (two_d) {
    if (!launch_snr(g, L.P, noise_new, L.S + kSnrHist * Kp, M, st)) return false;
    const uint64_t need = (uint64_t)kSnrHist;            // frame j is filtered once j >= 20
    first_active = h->dn_frames >= need ? 0 : (int)(need - h->dn_frames);
    if (first_active > M) first_active = M;
    const int targets = M - first_active;
    if (targets > 0) {
      // target of frame i is frame i-4: row (20+i)-4 of S
      if (!launch_nlm(g, L.S, kNlmPast + first_active, targets, h->nlm_h,
                      L.Shat + (size_t)first_active * Kp, st))
        return false;
      if (g_profile) ctx().nlm_frame_blocks += (uint64_t)targets * (uint64_t)((g.K + 7) / 8);
    }
  }

  // -- tonal mask and whitening weights: shared row when the noise is static --
  const int active = M - first_active;
  GainArgs ga;
  ga.frames = M;
  ga.first_active = first_active;
  ga.two_d = two_d;
  ga.strength = p.strength;
  ga.tonal = p.tonal;
  ga.reduction = p.reduction;
  ga.whitening = p.whitening;
  ga.depth = p.depth;
  ga.smoothing = p.smoothing;
  ga.residual = p.residual_listen;
  ga.gain_type = h->gain_type;
  ga.suppression_type = h->suppression_type;
  ga.smoothing_type = h->smoothing_type;
  ga.mask_stride = 0;
  ga.wt_stride = 0;
  ga.mask = h->d_aux_mask;
  ga.wt = h->d_aux_wt;
  if (active > 0) {
    bool noise_static = !adaptive;
    if (noise_static && two_d && first_active < kDelayFrames) {
      for (int r = first_active; r < kDelayFrames; r++)
        if (h->noise_hist_version[r] != h->noise_version) noise_static = false;
    }
    const bool transparent = (p.reduction >= 0.999f && p.tonal >= 0.999f);
    if (!transparent) {
      // tonal mask: detection profile = MAX profile once a median exists
      // (tonal_detector.c:35-44), otherwise the (delayed) noise spectrum
      if (h->profile_learned || noise_static) {
        if (h->aux_mask_version != h->noise_version) {
          const float* det = h->profile_learned ? h->d_prof + 2 * Kp : h->d_noise_cur;
          if (!launch_tonal_mask(g, det, 0, 1, h->d_aux_mask, st)) return false;
          h->aux_mask_version = h->noise_version;
        }
      } else {
        if (!launch_tonal_mask(g, L.noise + (size_t)first_active * Kp, g.Kp, active,
                               L.mask + (size_t)first_active * Kp, st))
          return false;
        ga.mask = L.mask;
        ga.mask_stride = g.Kp;
      }
      if (noise_static) {
        if (h->aux_wt_version != h->noise_version || h->aux_wt_whitening != p.whitening) {
          if (!launch_whitening(g, h->d_noise_cur, 0, 1, p.whitening, h->d_aux_wt, st))
            return false;
          h->aux_wt_version = h->noise_version;
          h->aux_wt_whitening = p.whitening;
        }
      } else {
        if (!launch_whitening(g, L.noise + (size_t)first_active * Kp, g.Kp, active, p.whitening,
                              L.wt + (size_t)first_active * Kp, st))
          return false;
        ga.wt = L.wt;
        ga.wt_stride = g.Kp;
      }
    }
  }
  if (!launch_gain_chain(g, ga, L, *h, st)) return false;

  // -- carry the delay line and the NLM ring to the next call --
  if (two_d) {
    if (!d2d(h->d_X_hist_re, L.Xre + (size_t)M * Kp, 4 * Kp, st) ||
        !d2d(h->d_X_hist_im, L.Xim + (size_t)M * Kp, 4 * Kp, st) ||
        !d2d(h->d_noise_hist, L.noise + (size_t)M * Kp, 4 * Kp, st) ||
        !d2d(h->d_S_hist, L.S + (size_t)M * Kp, kSnrHist * Kp, st))
      return false;
    uint64_t nv[4];
    for (int r = 0; r < 4; r++) {
      const int src = M + r;   // row in the 4+M buffer
      nv[r] = src < 4 ? h->noise_hist_version[src] : (adaptive ? 0 : h->noise_version);
    }
    memcpy(h->noise_hist_version, nv, sizeof(nv));
    h->dn_frames += (uint64_t)M;
  }
  return true;
}

bool sub_run(Handle* h, Lane& L, uint32_t m, float* d_out) {
  const GeoK& g = h->geo->k;
  cudaStream_t st = L.stream;
  const size_t Kp = g.Kp;
  const uint32_t hl = hist_len(h);
  const int M = (int)((h->carry + m) / (uint32_t)g.hop);
  if (M > L.cap_frames) {
    set_error_msg("specbleach-b200: internal error, sub-call exceeds lane capacity");
    return false;
  }
  const float* Yre = nullptr;
  const float* Yim = nullptr;
  if (M > 0) {
    ctx().frames += (uint64_t)M;
    if (h->prm.learn_noise > 0) {
      // learn mode: profiles updated, spectrum untouched and NOT delayed
      // (denoiser_profile_core.c:32-40, spectral_2d_denoiser.c:337-341)
      if (h->host_dirty && !upload_profiles(h, st)) return false;
      if (!launch_forward(g, L.xbuf, M, L.Xre, L.Xim, L.P, st)) return false;
      if (!launch_learn(g, L.P, M, h->block_count[0], h->d_prof, h->d_med_ring, h->med_write, st))
        return false;
      h->med_write = (h->med_write + M) % kMedianFrames;
      h->block_count[0] += (uint32_t)M;                                  // noise_profile.c:112-127
      if (h->block_count[0] > (uint32_t)kMinWindowsAveraged) h->available[0] = true;
      h->available[1] = h->available[2] = h->available[3] = true;         // noise_estimator.c:121-136
      h->was_learning = true;
      h->noise_version++;
      if (!download_profiles(h, st)) return false;
      h->profile_dl_pending = true;
      Yre = L.Xre;
      Yim = L.Xim;
    } else {
      if (!denoise_frames(h, L, M)) return false;
      Yre = L.Yre;
      Yim = L.Yim;
    }
    if (!launch_inverse(g, Yre, Yim, M, L.synth, st)) return false;
  }
  (void)Kp;
  const int cur = h->tail_cur;
  if (!launch_ola(g, L.synth, M, h->carry, m, h->d_tail[cur], h->d_tail[cur ^ 1], d_out, st))
    return false;
  h->tail_cur = cur ^ 1;
  // new FIFO tail = last frame-hop+carry' samples seen so far
  const uint32_t carry2 = (h->carry + m) % (uint32_t)g.hop;
  const uint32_t hl2 = (uint32_t)(g.frame - g.hop) + carry2;
  if (!d2d(h->d_hist, L.xbuf + (hl + m - hl2), hl2, st)) return false;
  h->carry = carry2;
  return true;
}

void after_sync(Handle* h) {
  if (h->profile_dl_pending) {
    h->profile_dl_pending = false;
    refresh_profile_learned(h);
  }
}

uint32_t max_sub_samples(const Handle* h) {
  const GeoK& g = h->geo->k;
  return (uint32_t)sub_call_frames(g) * (uint32_t)g.hop - h->carry;
}

// ---------------------------------------------------------------------------
// Whole-call drivers
// ---------------------------------------------------------------------------
// What kind of memory a caller's buffer is.  Device and page-locked host memory are DMA targets;
// anything else (malloc, numpy, the stack) is pageable: cudaMemcpyAsync would stage it through
// the driver's own bounce buffers synchronously, which serialises copies and kernels.
static bool is_device_or_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice ||
         a.type == cudaMemoryTypeManaged;
}

// One host sub-call on the lane's three-stream pipeline: H2D of `m` samples from the page-locked
// `src` on the copy-in stream into staging buffer b (it may run while the kernels of the
// previous sub-call are still busy), kernels on the lane stream, D2H to the page-locked `dst`
// on the copy-out stream.  ev_in[b] = H2D done, ev_done[b] = D2H done.
static bool enqueue_host_sub(Handle* h, Lane& L, uint32_t m, const float* src, float* dst, int* b_out,
                             uint32_t skip = 0) {
  if (!sub_begin(h, L)) return false;
  float* x = L.xbuf + hist_len(h);
  const int b = (int)(L.io_seq & 1u);
  const bool reused = L.io_seq >= 2;
  L.io_seq++;
  if (reused) SB_CUDA(cudaStreamWaitEvent(L.s_in, L.ev_used[b], 0));
  SB_CUDA(cudaMemcpyAsync(L.in_stage[b], src, m * sizeof(float), cudaMemcpyHostToDevice, L.s_in));
  SB_CUDA(cudaEventRecord(L.ev_in[b], L.s_in));
  SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_in[b], 0));
  if (!d2d(x, L.in_stage[b], m, L.stream)) return false;
  SB_CUDA(cudaEventRecord(L.ev_used[b], L.stream));
  if (reused) SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_done[b], 0));
  if (!sub_run(h, L, m, L.out_stage[b])) return false;
  SB_CUDA(cudaEventRecord(L.ev_out[b], L.stream));
  SB_CUDA(cudaStreamWaitEvent(L.s_out, L.ev_out[b], 0));
  // `skip`: leading output samples nobody wants (warm-up of a time tile)
  if (skip < m)
    SB_CUDA(cudaMemcpyAsync(dst + skip, L.out_stage[b] + skip, (m - skip) * sizeof(float),
                            cudaMemcpyDeviceToHost, L.s_out));
  SB_CUDA(cudaEventRecord(L.ev_done[b], L.s_out));
  *b_out = b;
  return true;
}

// Pageable caller buffers (what a program written against the reference passes to
// specbleach[_2d]_process: examples/denoiser_demo.c:217-300): the call is cut into sub-calls of a
// quarter of the lane's capacity, each copied into one of two page-locked bounce buffers of the
// lane and handed to the same pipeline; its output comes back through two more.  The host
// copies of sub-call k run while the GPU works on k-1, so the call costs max(host copies,
// kernels) instead of their sum.  in == out is fine: output k-2 is written behind the input
// position already consumed.
static bool lane_ensure_bounce(Lane& L) {
  if (L.pin_cap >= L.cap_samples && L.pin_in[0]) return true;
  for (int b = 0; b < 2; b++) {
    if (L.pin_in[b]) cudaFreeHost(L.pin_in[b]);
    if (L.pin_out[b]) cudaFreeHost(L.pin_out[b]);
    L.pin_in[b] = L.pin_out[b] = nullptr;
  }
  L.pin_cap = 0;
  for (int b = 0; b < 2; b++) {
    SB_CUDA(cudaMallocHost((void**)&L.pin_in[b], L.cap_samples * sizeof(float)));
    SB_CUDA(cudaMallocHost((void**)&L.pin_out[b], L.cap_samples * sizeof(float)));
  }
  L.pin_cap = L.cap_samples;
  return true;
}

// memcpy of a large piece split over a few threads: one host thread moves ~5 GB/s between
// pageable and page-locked memory on the bench box, a 10-min channel is 115 MB each way, and the
// kernels for it take 14 ms.  (The reference spends 4 OpenMP threads per handle by default.)
static void copy_piece(void* dst, const void* src, size_t bytes) {
  constexpr size_t kMinSlice = 1u << 20;
  constexpr int kThreads = 4;
  int nt = (int)(bytes / kMinSlice);
  if (nt > kThreads) nt = kThreads;
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t slice = ((bytes / (size_t)nt) + 63) & ~(size_t)63;
  std::thread th[kThreads - 1];
  for (int t = 1; t < nt; t++) {
    const size_t off = (size_t)t * slice;
    const size_t len = off >= bytes ? 0 : (t == nt - 1 ? bytes - off : slice);
    auto work = [=] {
      if (len) memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len);
    };
    try {
      th[t - 1] = std::thread(work);
    } catch (const std::system_error&) {   // no thread to be had: this slice on the calling thread
      work();
    }
  }
  memcpy(dst, src, slice < bytes ? slice : bytes);
  for (int t = 1; t < nt; t++)
    if (th[t - 1].joinable()) th[t - 1].join();
}

static bool enqueue_call_pageable(Handle* h, Lane& L, uint32_t n, const float* in, float* out,
                                  bool in_pageable, bool out_pageable) {
  if (!lane_ensure_bounce(L)) return false;
  const GeoK& gk = h->geo->k;
  // pieces: a short first and last one (their copies are the only exposed ones), half a lane's
  // capacity in between (full-size NLM chunks, few launch ramps)
  const uint32_t ramp = (uint32_t)(sub_call_frames(gk) / 8 > 64 ? sub_call_frames(gk) / 8 : 64) * (uint32_t)gk.hop;
  const uint32_t piece = (uint32_t)(sub_call_frames(gk) / 2 > 64 ? sub_call_frames(gk) / 2 : 64) * (uint32_t)gk.hop;
  // The calling thread copies the input pieces in and enqueues; a helper thread of the call waits
  // for each piece's D2H and copies it out, so input and output copies (each ~10 GB/s per thread)
  // run side by side.  A bounce buffer pair b is handed back by the helper before the caller
  // reuses it; the helper touches no engine state, only the lane's events and bounce buffers.
  struct Piece { uint32_t pos, m; int b; };
  struct Shared {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Piece> q;
    bool busy[2] = {false, false};      // bounce pair b holds a piece that has not been copied out
    bool done = false, failed = false;
  } sh;
  const int device = h->device;
  auto drain = [&] {
    cudaSetDevice(device);
    for (;;) {
      Piece p;
      {
        std::unique_lock<std::mutex> lk(sh.mu);
        sh.cv.wait(lk, [&] { return !sh.q.empty() || sh.done; });
        if (sh.q.empty()) return;
        p = sh.q.front();
        sh.q.pop_front();
      }
      const bool ok = cudaEventSynchronize(L.ev_done[p.b]) == cudaSuccess;
      if (ok && out_pageable) copy_piece(out + p.pos, L.pin_out[p.b], (size_t)p.m * sizeof(float));
      {
        std::lock_guard<std::mutex> lk(sh.mu);
        sh.busy[p.b] = false;
        if (!ok) sh.failed = true;
      }
      sh.cv.notify_all();
    }
  };
  std::thread helper;
  try {
    helper = std::thread(drain);
  } catch (const std::system_error&) {
    set_error_msg("specbleach-b200: cannot start the copy-out thread of a pageable call");
    return false;
  }
  bool ok = true;
  uint32_t pos = 0;
  while (pos < n && ok) {
    uint32_t m = n - pos;
    const uint32_t cap = max_sub_samples(h);
    if (pos == 0 && m > ramp) m = ramp;
    else if (m > piece + ramp) m = piece;
    else if (m > 2 * ramp) m -= ramp;
    if (m > cap) m = cap;
    const int b = (int)(L.io_seq & 1u);
    {
      // bounce pair b: the piece it held (sub-call k-2) must have been copied out
      std::unique_lock<std::mutex> lk(sh.mu);
      sh.cv.wait(lk, [&] { return !sh.busy[b]; });
      if (sh.failed) { ok = false; break; }
    }
    // ... and its H2D must be over before the input half is overwritten
    if (L.io_seq >= 2 && cudaEventSynchronize(L.ev_in[b]) != cudaSuccess) { ok = false; break; }
    const float* src = in + pos;
    if (in_pageable) {
      copy_piece(L.pin_in[b], in + pos, (size_t)m * sizeof(float));
      src = L.pin_in[b];
    }
    float* dst = out_pageable ? L.pin_out[b] : out + pos;
    int used = 0;
    if (!enqueue_host_sub(h, L, m, src, dst, &used)) { ok = false; break; }
    {
      std::lock_guard<std::mutex> lk(sh.mu);
      sh.busy[used] = true;
      sh.q.push_back(Piece{pos, m, used});
    }
    sh.cv.notify_all();
    pos += m;
  }
  {
    std::lock_guard<std::mutex> lk(sh.mu);
    sh.done = true;
  }
  sh.cv.notify_all();
  helper.join();
  if (sh.failed) {
    set_error("pageable output copy", cudaGetLastError(), __FILE__, __LINE__);
    ok = false;
  }
  return ok;
}

// Enqueues the whole call on the lane; does not synchronise (device / page-locked buffers).
bool enqueue_call(Handle* h, Lane& L, uint32_t n, const float* in, float* out, bool device_io,
                  uint32_t skip_out = 0) {
  SB_CUDA(cudaSetDevice(h->device));
  uint32_t pos = 0;
  bool host_pending = false;
  int last_b = 0;
  // Host buffers: the first H2D and the last D2H of a call are not hidden behind kernels, so a
  // call that spans several sub-calls starts and ends with a short one (1/8 of a sub-call).
  // Sub-call boundaries never change the output (streaming equivalence).
  const GeoK& gk = h->geo->k;
  const uint32_t ramp = (uint32_t)(sub_call_frames(gk) / 8 > 64 ? sub_call_frames(gk) / 8 : 64) *
                        (uint32_t)gk.hop;
  if (!device_io && n >= 2 * ramp && skip_out == 0) {
    // large host call: pageable buffers go through the bounce pipeline (which synchronises
    // itself); small calls are not worth two cudaPointerGetAttributes
    const bool in_pg = !is_device_or_pinned(in), out_pg = !is_device_or_pinned(out);
    if (in_pg || out_pg) return enqueue_call_pageable(h, L, n, in, out, in_pg, out_pg);
  }
  while (pos < n) {
    uint32_t m = n - pos;
    const uint32_t cap = max_sub_samples(h);
    if (!device_io && n > cap) {
      if (pos == 0 && ramp < cap) m = ramp;
      else if (m <= cap && m > 2 * ramp) m -= ramp;
    }
    if (m > cap) m = cap;
    if (device_io) {
      if (!sub_begin(h, L)) return false;
      if (!d2d(L.xbuf + hist_len(h), in + pos, m, L.stream)) return false;
      if (!sub_run(h, L, m, out + pos)) return false;
    } else {
      const uint32_t sk = skip_out > pos ? (skip_out - pos < m ? skip_out - pos : m) : 0;
      if (!enqueue_host_sub(h, L, m, in + pos, out + pos, &last_b, sk)) return false;
      // whoever synchronises the lane stream also waits for this copy
      host_pending = true;
    }
    pos += m;
  }
  if (host_pending) {
    // the D2H copies are ordered on s_out: waiting for the last one waits for all
    SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_done[last_b], 0));
  }
  return true;
}

bool process_call(Handle* h, uint32_t n, const float* in, float* out, bool device_io) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  if (!h->params_loaded) {
    // the reference processes with zero-initialised parameters; we do the same
    Params p;          // calloc'd Denoiser2DParameters: every field zero
    p.reduction = 0.0f;
    p.tonal = 0.0f;    // spectral_2d_denoiser.c:174
    h->prm = p;
  }
  DeviceGuard restore_device;
  SB_CUDA(cudaSetDevice(h->device));
  Lane* L = get_lane(0, h->geo->k);
  if (!L) return false;
  if (!enqueue_call(h, *L, n, in, out, device_io)) {
    cudaStreamSynchronize(L->stream);
    return false;
  }
  SB_CUDA(cudaStreamSynchronize(L->stream));
  after_sync(h);
  return true;
}

bool process_batch(Handle** hs, uint32_t count, const uint32_t* n, const float* const* in,
                   float* const* out, bool device_io) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  if (count == 0) return true;
  DeviceGuard restore_device;
  Context& c = ctx();
  bool ok = true;
  // handles are dealt round-robin to lanes; each lane serialises its own
  // handles, different lanes overlap copies and kernels.
  for (uint32_t i = 0; i < count && ok; i++) {
    Handle* h = hs[i];
    if (!h || !in[i] || !out[i] || n[i] == 0) { ok = false; break; }
    if (cudaSetDevice(h->device) != cudaSuccess) { ok = false; break; }
    Lane* L = get_lane((int)(i % (uint32_t)c.num_lanes), h->geo->k);
    if (!L) { ok = false; break; }
    ok = enqueue_call(h, *L, n[i], in[i], out[i], device_io);
  }
  for (auto& dl : c.lanes)
    for (Lane& L : dl.second)
      if (L.stream) cudaStreamSynchronize(L.stream);
  for (uint32_t i = 0; i < count; i++)
    if (hs[i]) after_sync(hs[i]);
  if (ok) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("batch", e, __FILE__, __LINE__);
      ok = false;
    }
  }
  return ok;
}

// ---------------------------------------------------------------------------
// N3: ONE long stream cut into time tiles that run on other lanes / GPUs.
// ---------------------------------------------------------------------------
// What a handle carries between frames forgets its start geometrically: the
// 21-frame NLM ring and the 4-frame delay line are exact after 25 frames, the veto EMAs (0.5 per
// frame) after ~150, the forward-masking recursion (<= 0.91 per frame at 48 kHz / 50 ms) after
// ~400.  A tile therefore starts `warmup_frames` hops early on a FRESH handle that was given the
// finalised noise profile and the parameters, and its warm-up output is dropped.  Tile 0
// continues the caller's handle (exact); at the end the caller's handle takes over the state of
// the last tile, so the stream can go on with ordinary process() calls.  Manual-profile mode:
// with >= 512 warm-up frames the result equals the monolithic call bit for bit on every input
// tried (tests/test_gpu_io.py).  Adaptive estimators have long memories (Martin: 96-frame minimum
// history, Brandt: 150-frame window + hold) and are only approximated: opt-in by flag.
// Tile handles are pooled per (device, geometry, denoiser kind): creating one costs 17 cudaMalloc
// and destroying one as many device-synchronising cudaFree, far more than a tile's kernels.
struct TilePoolKey {
  int device;
  const Geometry* geo;
  bool two_d;
  bool operator<(const TilePoolKey& o) const {
    if (device != o.device) return device < o.device;
    if (geo != o.geo) return geo < o.geo;
    return two_d < o.two_d;
  }
};
static std::map<TilePoolKey, std::vector<Handle*>> g_tile_pool;

// back to the state of a freshly created handle (stream-ordered on `st`)
static bool reset_stream_state(Handle* t, cudaStream_t st) {
  const GeoK& g = t->geo->k;
  const size_t Kp = g.Kp;
  struct Z { float* p; size_t n; };
  const Z z[] = {{t->d_med_ring, (size_t)kMedianFrames * Kp}, {t->d_hist, (size_t)g.frame + g.hop},
                 {t->d_tail[0], (size_t)g.frame}, {t->d_tail[1], (size_t)g.frame}, {t->d_stable, Kp},
                 {t->d_sp_prev, Kp}, {t->d_band_state, (size_t)kBandStateRows * kBandPitch}, {t->d_X_hist_re, 4 * Kp},
                 {t->d_X_hist_im, 4 * Kp}, {t->d_noise_hist, 4 * Kp}, {t->d_S_hist, (size_t)kSnrHist * Kp}};
  for (const Z& e : z)
    if (e.p) SB_CUDA(cudaMemsetAsync(e.p, 0, e.n * sizeof(float), st));
  t->med_write = 0;
  t->tail_cur = 0;
  t->carry = 0;
  t->was_learning = false;
  t->profile_dl_pending = false;
  t->dn_frames = 0;
  for (int r = 0; r < 4; r++) t->noise_hist_version[r] = 0;
  t->aux_mask_version = t->aux_wt_version = 0;
  t->aux_wt_whitening = -1.0f;
  t->last_adaptive_state = 0;
  t->params_loaded = false;
  t->prm = Params();
  if (t->d_est) {               // a fresh estimator of the next load_parameters
    SB_CUDA(cudaStreamSynchronize(st));
    dev_free(t->d_est);
    t->d_est = nullptr;
    t->est_floats = 0;
  }
  t->est_method = -1;
  return true;
}

static Handle* clone_for_tile(const Handle* h, int device, cudaStream_t st_hint) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  const Geometry* geo = get_geometry(h->geo->sample_rate, h->geo->frame_ms);
  Handle* t = nullptr;
  if (geo) {
    std::vector<Handle*>& pool = g_tile_pool[TilePoolKey{device, geo, h->two_d}];
    if (!pool.empty()) {
      t = pool.back();
      pool.pop_back();
      if (!reset_stream_state(t, st_hint)) {
        handle_destroy(t);
        t = nullptr;
      }
    }
  }
  if (!t) t = handle_create(h->two_d, h->geo->sample_rate, h->geo->frame_ms);
  if (!t) return nullptr;
  for (int m = 0; m < 4; m++) {
    t->prof_host[m] = h->prof_host[m];
    t->block_count[m] = h->block_count[m];
    t->available[m] = h->available[m];
  }
  t->profile_learned = h->profile_learned;
  t->host_dirty = true;
  t->noise_version++;
  t->nlm_h = h->nlm_h;
  t->gain_type = h->gain_type;
  t->suppression_type = h->suppression_type;
  t->smoothing_type = h->smoothing_type;
  if (h->params_loaded && !handle_load_parameters(t, h->prm)) {   // creates the estimator if adaptive
    handle_destroy(t);
    return nullptr;
  }
  t->nlm_h = h->nlm_h;
  return t;
}

static void release_tile_handle(Handle* t) {
  g_tile_pool[TilePoolKey{t->device, t->geo, t->two_d}].push_back(t);
}

// One call for several handles (the channels of a file): every handle's samples are cut into
// `tiles` tiles, ALL tiles are enqueued before anything is waited for, so the tiles of different
// channels overlap on the lanes of each GPU and the single host thread pays one synchronisation.
bool process_tiled_batch(Handle** hs, uint32_t count, const uint32_t* ns, const float* const* ins,
                         float* const* outs, uint32_t tiles, uint32_t warmup_frames, const int* devices,
                         uint32_t device_count, uint32_t flags) {
  std::lock_guard<std::recursive_mutex> engine_lock(g_engine_mu);
  DeviceGuard restore_device;
  struct Tile { uint32_t in_lo, out_lo, out_hi; Handle* h; int device; Lane* lane; };
  struct Job {
    Handle* h; uint32_t n; const float* in; float* out;
    std::vector<Tile> plan;
    bool reg_in = false, reg_out = false, plain = false;
  };
  std::vector<Job> jobs(count);
  for (uint32_t c = 0; c < count; c++) {
    Handle* h = hs[c];
    if (!h || !ins[c] || !outs[c] || ns[c] == 0) {
      set_error_msg("specbleach_b200_process_tiled: NULL handle / buffer or zero samples");
      return false;
    }
    if (h->prm.learn_noise > 0 || !h->params_loaded) {
      set_error_msg("specbleach_b200_process_tiled: the handle must be in denoise mode (load_parameters with learn_noise = 0)");
      return false;
    }
    if (h->prm.adaptive && !(flags & 1u)) {
      set_error_msg("specbleach_b200_process_tiled: adaptive noise estimation is only approximated by time "
                    "tiles; pass SPECBLEACH_B200_TILED_ALLOW_ADAPTIVE to accept that");
      return false;
    }
    for (uint32_t d = 0; d < c; d++)
      if (hs[d] == h) {
        set_error_msg("specbleach_b200_process_tiled_batch: a handle appears twice");
        return false;
      }
    Job& J = jobs[c];
    J.h = h; J.n = ns[c]; J.in = ins[c]; J.out = outs[c];
    const uint32_t n = J.n;
    const GeoK& g = h->geo->k;
    const uint32_t hop = (uint32_t)g.hop;
    const uint32_t first = (hop - h->carry) % hop;        // first frame boundary of the stream inside this call
    const uint32_t frames = n > first ? (n - first) / hop : 0;
    uint32_t nt = tiles;
    if (nt > frames / 64) nt = frames / 64;               // tiles shorter than 64 frames are pointless
    if (nt <= 1) { J.plain = true; continue; }
    if ((J.in <= J.out && J.out < J.in + n) || (J.out <= J.in && J.in < J.out + n)) {
      // a tile's warm-up input is the previous tile's output region: not in place
      set_error_msg("specbleach_b200_process_tiled: input and output must not overlap");
      return false;
    }
    for (uint32_t i = 0; i < nt; i++) {
      const uint32_t base = frames / nt, extra = frames % nt;
      const uint32_t lo_f = i * base + (i < extra ? i : extra);
      const uint32_t hi_f = lo_f + base + (i < extra ? 1u : 0u);
      Tile t;
      t.out_lo = i == 0 ? 0u : first + lo_f * hop;
      t.out_hi = i == nt - 1 ? n : first + hi_f * hop;
      t.in_lo = t.out_lo;
      t.h = nullptr;
      t.device = h->device;
      t.lane = nullptr;
      if (i > 0) {
        const uint64_t back = (uint64_t)warmup_frames * hop;
        if ((uint64_t)(t.out_lo - first) < back) {
          // its warm-up would reach before the call: tile 0 (the exact continuation) absorbs it
          J.plan[0].out_hi = t.out_hi;
          continue;
        }
        t.in_lo = t.out_lo - (uint32_t)back;
      }
      J.plan.push_back(t);
    }
  }
  bool ok = true;
  // the clones take the refined profile; page-locked buffers are DMA'd by every device, pageable
  // ones are registered for the call
  for (Job& J : jobs) {
    if (!ok) break;
    if (J.plain) continue;
    if (cudaSetDevice(J.h->device) != cudaSuccess) { ok = false; break; }
    Lane* L0 = get_lane(0, J.h->geo->k);
    if (!L0 || !finalize_profiles(J.h, L0->stream)) { ok = false; break; }
    if (!is_device_or_pinned(J.in)) {
      if (cudaHostRegister(const_cast<float*>(J.in), (size_t)J.n * sizeof(float), cudaHostRegisterPortable) == cudaSuccess)
        J.reg_in = true;
      else cudaGetLastError();
    }
    if (!is_device_or_pinned(J.out)) {
      if (cudaHostRegister(J.out, (size_t)J.n * sizeof(float), cudaHostRegisterPortable) == cudaSuccess) J.reg_out = true;
      else cudaGetLastError();
    }
  }
  const int nlanes = ctx().num_lanes;
  std::map<int, size_t> next_lane;                        // per device: tiles go round its lanes
  for (Job& J : jobs) {
    if (!ok) break;
    if (J.plain) continue;
    const size_t nt = J.plan.size();
    for (size_t i = 0; i < nt && ok; i++) {
      Tile& t = J.plan[i];
      // tile 0 continues the handle on its own device; the others go round the listed devices
      if (i > 0 && devices && device_count) t.device = devices[i % device_count];
      if (cudaSetDevice(t.device) != cudaSuccess) {
        set_error("specbleach_b200_process_tiled: cudaSetDevice on a listed device", cudaGetLastError(),
                  __FILE__, __LINE__);
        ok = false;
        break;
      }
      t.lane = get_lane((int)(next_lane[t.device]++ % (size_t)nlanes), J.h->geo->k);
      if (!t.lane) { ok = false; break; }
      if (i == 0) {
        t.h = J.h;
      } else {
        t.h = clone_for_tile(J.h, t.device, t.lane->stream);
        if (!t.h) { ok = false; break; }
      }
      ok = enqueue_call(t.h, *t.lane, t.out_hi - t.in_lo, J.in + t.in_lo, J.out + t.in_lo, false,
                        t.out_lo - t.in_lo);
    }
  }
  for (auto& dl : ctx().lanes) {
    if (cudaSetDevice(dl.first) != cudaSuccess) continue;
    for (Lane& L : dl.second)
      if (L.stream) cudaStreamSynchronize(L.stream);
  }
  if (ok) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("process_tiled", e, __FILE__, __LINE__);
      ok = false;
    }
  }
  for (Job& J : jobs) {
    if (J.plain) continue;
    const size_t nt = J.plan.size();
    after_sync(J.h);
    // the stream goes on from the last tile's state: same device -> the two handles trade places;
    // another device -> its state travels as a snapshot blob (~0.5 MB)
    if (ok && nt > 1 && J.plan[nt - 1].h) {
      Handle* last = J.plan[nt - 1].h;
      if (last->device == J.h->device) {
        std::swap(*J.h, *last);
      } else {
        std::vector<unsigned char> blob(state_size(last));
        ok = state_save(last, blob.data(), blob.size()) && state_load(J.h, blob.data(), blob.size());
      }
    }
    for (size_t i = 1; i < nt; i++)
      if (J.plan[i].h) release_tile_handle(J.plan[i].h);
    if (J.reg_in) cudaHostUnregister(const_cast<float*>(J.in));
    if (J.reg_out) cudaHostUnregister(J.out);
  }
  // handles too short to tile: the ordinary call
  for (Job& J : jobs)
    if (ok && J.plain) ok = process_call(J.h, J.n, J.in, J.out, false);
  return ok;
}

bool process_tiled(Handle* h, uint32_t n, const float* in, float* out, uint32_t tiles,
                   uint32_t warmup_frames, const int* devices, uint32_t device_count, uint32_t flags) {
  return process_tiled_batch(&h, 1, &n, &in, &out, tiles, warmup_frames, devices, device_count, flags);
}

bool stats_collect(Context& c) {
  for (auto& dl : c.lanes)
    for (Lane& L : dl.second)
      if (L.stream) SB_CUDA(cudaStreamSynchronize(L.stream));
  SB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_kev_mu);
  // measurement knob: SB_TIMELINE=<file> appends "class,stream,start_ms,end_ms" of every timed
  // launch relative to the first one (event times of different streams share one clock): a
  // poor man's timeline of how the lanes interleave on the device
  if (const char* path = getenv("SB_TIMELINE")) {
    if (!g_kevents.empty()) {
      if (FILE* f = fopen(path, "a")) {
        for (KEvent& k : g_kevents) {
          float a = 0.f, b = 0.f;
          if (cudaEventElapsedTime(&a, g_kevents[0].e0, k.e0) == cudaSuccess &&
              cudaEventElapsedTime(&b, g_kevents[0].e0, k.e1) == cudaSuccess)
            fprintf(f, "%s,%p,%.4f,%.4f\n", kclass_name(k.cls), (void*)k.st, a, b);
        }
        fprintf(f, "#\n");
        fclose(f);
      }
    }
    cudaGetLastError();
  }
  for (KEvent& k : g_kevents) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, k.e0, k.e1) == cudaSuccess) {
      c.class_ms[k.cls] += ms;
      c.class_launches[k.cls]++;
      if (k.cls == kcNlm) { c.nlm_ms += ms; c.nlm_launches++; }
    }
    cudaEventDestroy(k.e0);
    cudaEventDestroy(k.e1);
  }
  g_kevents.clear();
  return true;
}

}  // namespace sb
