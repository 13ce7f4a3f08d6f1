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
#include <stdlib.h>
#include <string.h>

#include <mutex>

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
struct KEvent { int cls; cudaEvent_t e0, e1; };
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

static bool dev_alloc(float** p, size_t floats) {
  SB_CUDA(cudaMalloc((void**)p, floats * sizeof(float)));
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
    if (*b) cudaFree(*b);
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
            dev_zalloc(&h->d_band_state, 2 * kBandPitch) && dev_zalloc(&h->d_noise_cur, Kp) &&
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
    if (b) cudaFree(b);
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
      if (h->d_est) cudaFree(h->d_est);
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

static bool denoise_frames(Handle* h, Lane& L, int M) {
  const GeoK& g = h->geo->k;
  cudaStream_t st = L.stream;
  const size_t Kp = g.Kp;
  const bool two_d = h->two_d;
  const Params& p = h->prm;
  const int lead = two_d ? kDelayFrames : 0;   // history rows in front of X / noise

  // -- learn -> denoise transition: refine every available profile once --
  if (h->was_learning) {   // denoiser_profile_core.c:42-48
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
  }
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
static bool is_device_or_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice ||
         a.type == cudaMemoryTypeManaged;
}

// Enqueues the whole call on the lane; does not synchronise.
bool enqueue_call(Handle* h, Lane& L, uint32_t n, const float* in, float* out, bool device_io) {
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
  while (pos < n) {
    uint32_t m = n - pos;
    const uint32_t cap = max_sub_samples(h);
    if (!device_io && n > cap) {
      if (pos == 0 && ramp < cap) m = ramp;
      else if (m <= cap && m > 2 * ramp) m -= ramp;
    }
    if (m > cap) m = cap;
    if (!sub_begin(h, L)) return false;
    float* dst = L.xbuf + hist_len(h);
    if (device_io) {
      if (!d2d(dst, in + pos, m, L.stream)) return false;
      if (!sub_run(h, L, m, out + pos)) return false;
    } else {
      // H2D on its own stream into a staging buffer (it may run while the kernels of the
      // previous sub-call are still busy), kernels on the lane stream, D2H on a third stream
      const int b = (int)(L.io_seq & 1u);
      const bool reused = L.io_seq >= 2;
      L.io_seq++;
      if (reused) SB_CUDA(cudaStreamWaitEvent(L.s_in, L.ev_used[b], 0));
      SB_CUDA(cudaMemcpyAsync(L.in_stage[b], in + pos, m * sizeof(float), cudaMemcpyHostToDevice,
                              L.s_in));
      SB_CUDA(cudaEventRecord(L.ev_in[b], L.s_in));
      SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_in[b], 0));
      if (!d2d(dst, L.in_stage[b], m, L.stream)) return false;
      SB_CUDA(cudaEventRecord(L.ev_used[b], L.stream));
      if (reused) SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_done[b], 0));
      if (!sub_run(h, L, m, L.out_stage[b])) return false;
      SB_CUDA(cudaEventRecord(L.ev_out[b], L.stream));
      SB_CUDA(cudaStreamWaitEvent(L.s_out, L.ev_out[b], 0));
      SB_CUDA(cudaMemcpyAsync(out + pos, L.out_stage[b], m * sizeof(float), cudaMemcpyDeviceToHost,
                              L.s_out));
      SB_CUDA(cudaEventRecord(L.ev_done[b], L.s_out));
      // whoever synchronises the lane stream also waits for this copy
      host_pending = true;
      last_b = b;
    }
    pos += m;
  }
  if (host_pending) {
    // the D2H copies are ordered on s_out: waiting for the last one waits for all
    SB_CUDA(cudaStreamWaitEvent(L.stream, L.ev_done[last_b], 0));
  }
  (void)is_device_or_pinned;
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

bool stats_collect(Context& c) {
  for (auto& dl : c.lanes)
    for (Lane& L : dl.second)
      if (L.stream) SB_CUDA(cudaStreamSynchronize(L.stream));
  SB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_kev_mu);
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
