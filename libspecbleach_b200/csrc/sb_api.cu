// C ABI: the 22 drop-in symbols of the reference plus the specbleach_b200_*
// extensions.  Argument checking and return values follow the reference's
// src/processors/specbleach_denoiser.c and specbleach_2d_denoiser.c.
#include "../../include/specbleach_2d_denoiser.h"
#include "../../include/specbleach_b200.h"
#include "../../include/specbleach_denoiser.h"
#include "sb_engine.h"

#include <math.h>
#include <string.h>

using sb::Handle;

namespace {

// general_utils.c:33-39
float db_to_coefficient(float db) { return expf(db / 20.f * logf(10.f)); }
float remap_percentage_log_like_unity(float v) { return 1.f - expf(-3.f * v); }

Handle* H(SpectralBleachHandle p) { return reinterpret_cast<Handle*>(p); }

bool valid_mode(int mode) { return mode >= 1 && mode <= 4; }

uint32_t block_count(SpectralBleachHandle inst, int mode) {
  if (!inst || !valid_mode(mode)) return 0;
  return H(inst)->block_count[mode - 1];
}
float* profile_ptr(SpectralBleachHandle inst, int mode) {
  if (!inst || !valid_mode(mode)) return nullptr;
  return H(inst)->prof_host[mode - 1].data();
}
bool profile_available(SpectralBleachHandle inst, int mode) {
  if (!inst || !valid_mode(mode)) return false;
  return H(inst)->available[mode - 1];
}
bool load_profile(SpectralBleachHandle inst, const float* prof, uint32_t size, uint32_t blocks,
                  int mode) {
  if (!inst || !prof || !valid_mode(mode)) return false;
  return sb::guarded(false, [&] { return sb::handle_load_profile(H(inst), prof, size, blocks, mode); });
}

}  // namespace

extern "C" {

// ------------------------------- 1D ---------------------------------------
SpectralBleachHandle specbleach_initialize(uint32_t sample_rate, float frame_size) {
  if (sample_rate < 4000 || sample_rate > 192000 || !(frame_size > 0.0f)) return nullptr;
  return sb::guarded((Handle*)nullptr, [&] { return sb::handle_create(false, sample_rate, frame_size); });
}
void specbleach_free(SpectralBleachHandle instance) { sb::handle_destroy(H(instance)); }

bool specbleach_load_parameters(SpectralBleachHandle instance,
                                SpectralBleachDenoiserParameters q) {
  if (!instance) return false;
  sb::Params p;   // specbleach_denoiser.c:208-223
  p.learn_noise = q.learn_noise;
  p.residual_listen = q.residual_listen;
  p.reduction = db_to_coefficient(q.reduction_amount * -1.f);
  p.smoothing = remap_percentage_log_like_unity(q.smoothing_factor / 100.f);
  p.whitening = q.whitening_factor / 100.f;
  p.adaptive = q.adaptive_noise;
  p.method = q.noise_estimation_method;
  p.depth = q.masking_depth;
  p.strength = q.suppression_strength / 100.f;
  p.aggressiveness = q.aggressiveness;
  p.tonal = db_to_coefficient(q.tonal_reduction * -1.f);
  return sb::guarded(false, [&] { return sb::handle_load_parameters(H(instance), p); });
}
bool specbleach_process(SpectralBleachHandle instance, uint32_t n, const float* input,
                        float* output) {
  if (!instance || n == 0 || !input || !output) return false;
  return sb::guarded(false, [&] { return sb::process_call(H(instance), n, input, output, false); });
}
uint32_t specbleach_get_latency(SpectralBleachHandle instance) {
  if (!instance) return 0;
  return (uint32_t)H(instance)->geo->k.frame;   // stft_processor.c:69
}
uint32_t specbleach_get_noise_profile_size(SpectralBleachHandle instance) {
  if (!instance) return 0;
  return H(instance)->profile_size;
}
bool specbleach_load_noise_profile_for_mode(SpectralBleachHandle instance, const float* prof,
                                            uint32_t size, uint32_t blocks, int mode) {
  return load_profile(instance, prof, size, blocks, mode);
}
bool specbleach_reset_noise_profile(SpectralBleachHandle instance) {
  if (!instance) return false;
  sb::handle_reset_profile(H(instance));
  return true;
}
uint32_t specbleach_get_noise_profile_block_count_for_mode(SpectralBleachHandle instance,
                                                           int mode) {
  return block_count(instance, mode);
}
float* specbleach_get_noise_profile_for_mode(SpectralBleachHandle instance, int mode) {
  return profile_ptr(instance, mode);
}
bool specbleach_noise_profile_available_for_mode(SpectralBleachHandle instance, int mode) {
  return profile_available(instance, mode);
}

// ------------------------------- 2D ---------------------------------------
SpectralBleachHandle specbleach_2d_initialize(uint32_t sample_rate, float frame_size) {
  if (sample_rate == 0 || !(frame_size > 0.0f)) return nullptr;
  return sb::guarded((Handle*)nullptr, [&] { return sb::handle_create(true, sample_rate, frame_size); });
}
void specbleach_2d_free(SpectralBleachHandle instance) { sb::handle_destroy(H(instance)); }

bool specbleach_2d_load_parameters(SpectralBleachHandle instance,
                                   SpectralBleach2DDenoiserParameters q) {
  if (!instance) return false;
  sb::Params p;   // specbleach_2d_denoiser.c:234-248
  p.learn_noise = q.learn_noise;
  p.residual_listen = q.residual_listen;
  p.reduction = db_to_coefficient(q.reduction_amount * -1.f);
  p.smoothing = q.smoothing_factor;
  p.whitening = q.whitening_factor / 100.f;
  p.adaptive = q.adaptive_noise;
  p.method = q.noise_estimation_method;
  p.depth = q.nlm_masking_protection;
  p.strength = q.suppression_strength / 100.f;
  p.aggressiveness = q.aggressiveness;
  p.tonal = db_to_coefficient(q.tonal_reduction * -1.f);
  return sb::guarded(false, [&] { return sb::handle_load_parameters(H(instance), p); });
}
bool specbleach_2d_process(SpectralBleachHandle instance, uint32_t n, const float* input,
                           float* output) {
  if (!instance || n == 0 || !input || !output) return false;
  return sb::guarded(false, [&] { return sb::process_call(H(instance), n, input, output, false); });
}
uint32_t specbleach_2d_get_latency(SpectralBleachHandle instance) {
  if (!instance) return 0;
  const sb::GeoK& g = H(instance)->geo->k;
  // reported value: frame + 4 * (fft_size / 4)  (specbleach_2d_denoiser.c:63,102-118)
  return (uint32_t)g.frame + (uint32_t)sb::kNlmFuture * (uint32_t)(g.N / sb::kOverlap);
}
uint32_t specbleach_2d_get_noise_profile_size(SpectralBleachHandle instance) {
  if (!instance) return 0;
  return H(instance)->profile_size;
}
bool specbleach_2d_load_noise_profile_for_mode(SpectralBleachHandle instance, const float* prof,
                                               uint32_t size, uint32_t blocks, int mode) {
  return load_profile(instance, prof, size, blocks, mode);
}
bool specbleach_2d_reset_noise_profile(SpectralBleachHandle instance) {
  if (!instance) return false;
  sb::handle_reset_profile(H(instance));
  return true;
}
uint32_t specbleach_2d_get_noise_profile_block_count_for_mode(SpectralBleachHandle instance,
                                                              int mode) {
  return block_count(instance, mode);
}
float* specbleach_2d_get_noise_profile_for_mode(SpectralBleachHandle instance, int mode) {
  return profile_ptr(instance, mode);
}
bool specbleach_2d_noise_profile_available_for_mode(SpectralBleachHandle instance, int mode) {
  return profile_available(instance, mode);
}

// ----------------------------- extensions ----------------------------------
int specbleach_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
bool specbleach_b200_set_device(int device) {
  if (cudaSetDevice(device) != cudaSuccess) {
    sb::set_error("cudaSetDevice", cudaGetLastError(), __FILE__, __LINE__);
    return false;
  }
  return true;
}
const char* specbleach_b200_last_error(void) { return sb::last_error(); }

void* specbleach_b200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    sb::set_error("cudaMallocHost", cudaGetLastError(), __FILE__, __LINE__);
    return nullptr;
  }
  return p;
}
void specbleach_b200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
bool specbleach_b200_configure(uint32_t chunk_frames, uint32_t lanes) {
  std::lock_guard<std::recursive_mutex> engine_lock(sb::engine_mutex());
  sb::Context& c = sb::ctx();
  if (!c.lanes.empty()) {
    sb::set_error_msg("specbleach_b200_configure must be called before the first process()");
    return false;
  }
  if (chunk_frames != 0 && (chunk_frames < 32 || chunk_frames > 32768)) {
    sb::set_error_msg("specbleach_b200_configure: chunk_frames must be 0 (automatic) or in [32, 32768]");
    return false;
  }
  if (lanes < 1 || lanes > 32) {
    sb::set_error_msg("specbleach_b200_configure: lanes must be in [1, 32]");
    return false;
  }
  c.chunk_frames = (int)chunk_frames;
  c.num_lanes = (int)lanes;
  return true;
}
bool specbleach_b200_process_device(SpectralBleachHandle instance, uint32_t n,
                                    const float* d_input, float* d_output) {
  if (!instance || n == 0 || !d_input || !d_output) return false;
  return sb::guarded(false, [&] { return sb::process_call(H(instance), n, d_input, d_output, true); });
}
bool specbleach_b200_process_batch(SpectralBleachHandle* instances, uint32_t count,
                                   const uint32_t* n, const float* const* inputs,
                                   float* const* outputs) {
  if (!instances || !n || !inputs || !outputs) return false;
  return sb::guarded(false, [&] {
    return sb::process_batch(reinterpret_cast<Handle**>(instances), count, n, inputs, outputs, false);
  });
}
bool specbleach_b200_process_batch_device(SpectralBleachHandle* instances, uint32_t count,
                                          const uint32_t* n, const float* const* d_inputs,
                                          float* const* d_outputs) {
  if (!instances || !n || !d_inputs || !d_outputs) return false;
  return sb::guarded(false, [&] {
    return sb::process_batch(reinterpret_cast<Handle**>(instances), count, n, d_inputs, d_outputs, true);
  });
}

bool specbleach_b200_process_tiled(SpectralBleachHandle instance, uint32_t n, const float* input,
                                   float* output, uint32_t tiles, uint32_t warmup_frames,
                                   const int* devices, uint32_t device_count, uint32_t flags) {
  if (!instance || n == 0 || !input || !output) return false;
  return sb::guarded(false, [&] {
    return sb::process_tiled(H(instance), n, input, output, tiles, warmup_frames, devices, device_count, flags);
  });
}

bool specbleach_b200_process_tiled_batch(SpectralBleachHandle* instances, uint32_t count,
                                         const uint32_t* number_of_samples, const float* const* inputs,
                                         float* const* outputs, uint32_t tiles, uint32_t warmup_frames,
                                         const int* devices, uint32_t device_count, uint32_t flags) {
  if (!instances || !number_of_samples || !inputs || !outputs || count == 0) return false;
  return sb::guarded(false, [&] {
    return sb::process_tiled_batch(reinterpret_cast<Handle**>(instances), count, number_of_samples, inputs,
                                   outputs, tiles, warmup_frames, devices, device_count, flags);
  });
}

bool specbleach_b200_set_strategy(SpectralBleachHandle instance, int gain_type, int suppression_type,
                                  int smoothing_type) {
  if (!instance) return false;
  if (gain_type < 0 || gain_type > 2 || suppression_type < 0 || suppression_type > 4 ||
      smoothing_type < 1 || smoothing_type > 2) {
    sb::set_error_msg("specbleach_b200_set_strategy: gain 0..2, suppression 0..4, smoothing 1..2");
    return false;
  }
  std::lock_guard<std::recursive_mutex> engine_lock(sb::engine_mutex());
  Handle* h = H(instance);
  h->gain_type = gain_type;
  h->suppression_type = suppression_type;
  h->smoothing_type = smoothing_type;
  return true;
}

void specbleach_b200_profile(int enable) { sb::g_profile = enable != 0; }
void specbleach_b200_stats_reset(void) {
  sb::Context& c = sb::ctx();
  sb::stats_collect(c);
  c.nlm_ms = 0.0;
  c.nlm_launches = 0;
  c.nlm_frame_blocks = 0;
  c.frames = 0;
  for (int i = 0; i < sb::kcCount; i++) { c.class_ms[i] = 0.0; c.class_launches[i] = 0; }
  sb::g_launches = 0;
}
int specbleach_b200_stats_classes(void) { return sb::kcCount; }
const char* specbleach_b200_stats_class_name(int cls) { return sb::kclass_name(cls); }
bool specbleach_b200_stats_class(int cls, double* ms, uint64_t* launches) {
  if (cls < 0 || cls >= sb::kcCount || !ms || !launches) return false;
  sb::Context& c = sb::ctx();
  if (!sb::stats_collect(c)) return false;
  *ms = c.class_ms[cls];
  *launches = c.class_launches[cls];
  return true;
}
bool specbleach_b200_stats_get(SpecbleachB200Stats* out) {
  if (!out) return false;
  sb::Context& c = sb::ctx();
  if (!sb::stats_collect(c)) return false;
  out->kernel_launches = sb::g_launches;
  out->nlm_launches = c.nlm_launches;
  out->nlm_ms = c.nlm_ms;
  out->nlm_frame_blocks = c.nlm_frame_blocks;
  out->frames = c.frames;
  return true;
}
bool specbleach_b200_get_geometry(SpectralBleachHandle instance, uint32_t* frame,
                                  uint32_t* fft_size, uint32_t* hop, uint32_t* bins) {
  if (!instance) return false;
  const sb::GeoK& g = H(instance)->geo->k;
  if (frame) *frame = (uint32_t)g.frame;
  if (fft_size) *fft_size = (uint32_t)g.N;
  if (hop) *hop = (uint32_t)g.hop;
  if (bins) *bins = (uint32_t)g.K;
  return true;
}

// ---- FP32-pipe peak: FFMA micro-benchmark (denominator of the NLM roofline) ----
}  // extern "C" (kernel below has C++ linkage)

namespace {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fmaf(v[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the loop alive
}
}  // namespace

extern "C" {

int specbleach_b200_debug_check_guards(void) {
  std::lock_guard<std::recursive_mutex> engine_lock(sb::engine_mutex());
  return sb::check_guards();
}

double specbleach_b200_debug_fp32_peak_tflops(void) {
  float* d = nullptr;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  if (sb::dev_malloc(&d, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return 0.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, 0);
    ffma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  sb::dev_free(d);
  return best;
}

// ---- single-kernel entry points for the parity tests ----
bool specbleach_b200_debug_nlm(const float* snr, uint32_t frames, uint32_t bins, float h,
                               float* out) {
  if (!snr || !out || frames == 0 || bins < 8) return false;
  sb::GeoK g;
  memset(&g, 0, sizeof(g));
  g.K = (int)bins;
  g.Kp = ((int)bins + 3) & ~3;
  const size_t Kp = g.Kp;
  // rows: 20 zero history frames (a freshly calloc'd ring) followed by the map
  const size_t rows = sb::kSnrHist + frames;
  float *dS = nullptr, *dO = nullptr;
  bool ok = true;
  if (sb::dev_malloc(&dS, rows * Kp * sizeof(float)) != cudaSuccess ||
      sb::dev_malloc(&dO, (size_t)frames * Kp * sizeof(float)) != cudaSuccess) {
    sb::set_error("cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
    ok = false;
  }
  if (ok) {
    cudaMemset(dS, 0, rows * Kp * sizeof(float));
    cudaMemset(dO, 0, (size_t)frames * Kp * sizeof(float));
    ok = cudaMemcpy2D(dS + sb::kSnrHist * Kp, Kp * sizeof(float), snr, bins * sizeof(float),
                      bins * sizeof(float), frames, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok && frames > (uint32_t)sb::kSnrHist) {
    // frame j (j >= 20) -> target row (20 + j) - 4 of dS, written to out row j
    const int first = sb::kSnrHist;
    const int targets = (int)frames - first;
    for (int t0 = 0; t0 < targets && ok; t0 += 4096) {
      const int cnt = targets - t0 < 4096 ? targets - t0 : 4096;
      ok = sb::launch_nlm(g, dS, sb::kNlmPast + first + t0, cnt, h,
                          dO + (size_t)(first + t0) * Kp, 0);
    }
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    if (ok)
      ok = cudaMemcpy2D(out + (size_t)first * bins, bins * sizeof(float), dO + (size_t)first * Kp,
                        Kp * sizeof(float), bins * sizeof(float), targets,
                        cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  if (!ok && cudaPeekAtLastError() != cudaSuccess)
    sb::set_error("debug_nlm", cudaGetLastError(), __FILE__, __LINE__);
  sb::dev_free(dS);
  sb::dev_free(dO);
  return ok;
}

bool specbleach_b200_debug_rfft(uint32_t fft_size, uint32_t frames, const float* in,
                                float* halfcomplex, float* roundtrip) {
  if (!in || !halfcomplex || !roundtrip || frames == 0) return false;
  sb::Geometry* G = sb::make_fft_geometry((int)fft_size);
  if (!G) return false;
  const sb::GeoK& g = G->k;
  const size_t Kp = g.Kp, N = g.N;
  float *dx = nullptr, *dre = nullptr, *dim = nullptr, *dp = nullptr, *dsyn = nullptr;
  bool ok = sb::dev_malloc(&dx, frames * N * sizeof(float)) == cudaSuccess &&
            sb::dev_malloc(&dre, frames * Kp * sizeof(float)) == cudaSuccess &&
            sb::dev_malloc(&dim, frames * Kp * sizeof(float)) == cudaSuccess &&
            sb::dev_malloc(&dp, frames * Kp * sizeof(float)) == cudaSuccess &&
            sb::dev_malloc(&dsyn, frames * (size_t)g.frame_p * sizeof(float)) == cudaSuccess;
  if (ok) ok = cudaMemcpy(dx, in, frames * N * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  // hop == N: frame f = dx[f*N, (f+1)*N)
  ok = ok && sb::launch_forward(g, dx, (int)frames, dre, dim, dp, 0);
  ok = ok && sb::launch_inverse(g, dre, dim, (int)frames, dsyn, 0);
  ok = ok && cudaDeviceSynchronize() == cudaSuccess;
  if (ok) {
    std::vector<float> re(frames * Kp), im(frames * Kp);
    ok = cudaMemcpy(re.data(), dre, re.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(im.data(), dim, im.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy2D(roundtrip, N * sizeof(float), dsyn, (size_t)g.frame_p * sizeof(float),
                      N * sizeof(float), frames, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok) {
      for (size_t f = 0; f < frames; f++) {
        float* hc = halfcomplex + f * N;
        for (size_t k = 0; k <= N / 2; k++) hc[k] = re[f * Kp + k];
        for (size_t k = 1; k < N / 2; k++) hc[N - k] = im[f * Kp + k];
      }
    }
  }
  if (!ok && cudaPeekAtLastError() != cudaSuccess)
    sb::set_error("debug_rfft", cudaGetLastError(), __FILE__, __LINE__);
  sb::dev_free(dx); sb::dev_free(dre); sb::dev_free(dim); sb::dev_free(dp); sb::dev_free(dsyn);
  sb::free_geometry(G);
  return ok;
}

}  // extern "C"
