// The steps on either side of the hot path:
//   * noise-profile persistence (the reference only has get/load pointers,
//     specbleach_2d_denoiser.c:145-206) and a snapshot of the complete streaming state
//     of a handle, so that batch jobs can share one learned profile across thousands of
//     streams and resume a stream on another process / GPU;
//   * interleaved multichannel PCM in and out of the per-channel handles (what
//     examples/denoiser_demo.c does around libsndfile on the host, one sample at a time):
//     de-interleave + convert to float and convert + interleave back are device kernels,
//     the channels run as one batch.
#include "sb_engine.h"

#include <stdio.h>
#include <string.h>

#include <vector>

namespace sb {

// ---------------------------------------------------------------------------
// CRC-32 (IEEE 802.3) for the profile file and the state blob
// ---------------------------------------------------------------------------
static uint32_t crc32(const void* data, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  crc = ~crc;
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
  return ~crc;
}

// ---------------------------------------------------------------------------
// Noise profile file: "SBNP" v1, little endian
//   char[4] magic, u32 version, u32 sample_rate, f32 frame_ms, u32 two_d, u32 profile_size,
//   4 x { u32 block_count, u32 available, f32[profile_size] }   (modes 1..4), u32 crc32
// ---------------------------------------------------------------------------
struct ProfileHeader {
  char magic[4];
  uint32_t version, sample_rate;
  float frame_ms;
  uint32_t two_d, profile_size;
};

bool profile_serialize(Handle* h, std::vector<unsigned char>& out) {
  ProfileHeader hd;
  memcpy(hd.magic, "SBNP", 4);
  hd.version = 1;
  hd.sample_rate = h->geo->sample_rate;
  hd.frame_ms = h->geo->frame_ms;
  hd.two_d = h->two_d ? 1u : 0u;
  hd.profile_size = h->profile_size;
  out.clear();
  auto put = [&](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    out.insert(out.end(), b, b + n);
  };
  put(&hd, sizeof(hd));
  for (int m = 0; m < 4; m++) {
    const uint32_t bc = h->block_count[m], av = h->available[m] ? 1u : 0u;
    put(&bc, 4);
    put(&av, 4);
    put(h->prof_host[m].data(), (size_t)h->profile_size * sizeof(float));
  }
  const uint32_t c = crc32(out.data(), out.size());
  put(&c, 4);
  return true;
}

bool profile_deserialize(Handle* h, const unsigned char* data, size_t n) {
  ProfileHeader hd;
  if (n < sizeof(hd) + 4) {
    set_error_msg("specbleach-b200: noise profile file truncated");
    return false;
  }
  memcpy(&hd, data, sizeof(hd));
  uint32_t c;
  memcpy(&c, data + n - 4, 4);
  if (memcmp(hd.magic, "SBNP", 4) != 0 || hd.version != 1 || c != crc32(data, n - 4)) {
    set_error_msg("specbleach-b200: not a noise profile file (magic / version / checksum)");
    return false;
  }
  if (hd.sample_rate != h->geo->sample_rate || hd.frame_ms != h->geo->frame_ms ||
      (hd.two_d != 0) != h->two_d || hd.profile_size != h->profile_size ||
      n != sizeof(hd) + 4 * (8 + (size_t)hd.profile_size * sizeof(float)) + 4) {
    set_error_msg("specbleach-b200: noise profile was saved for another denoiser geometry");
    return false;
  }
  const unsigned char* p = data + sizeof(hd);
  for (int m = 0; m < 4; m++) {
    uint32_t bc, av;
    memcpy(&bc, p, 4);
    memcpy(&av, p + 4, 4);
    p += 8;
    // same path as specbleach[_2d]_load_noise_profile_for_mode (copy + block count)
    std::vector<float> tmp(hd.profile_size);
    memcpy(tmp.data(), p, (size_t)hd.profile_size * sizeof(float));
    p += (size_t)hd.profile_size * sizeof(float);
    if (!handle_load_profile(h, tmp.data(), hd.profile_size, bc, m + 1)) return false;
    h->available[m] = av != 0;
  }
  return true;
}

// ---------------------------------------------------------------------------
// Streaming-state snapshot: everything SURVEY appendix B lists, as one host blob
//   "SBST" v1 | geometry | host scalars | device arrays in a fixed order | crc32
// ---------------------------------------------------------------------------
struct StateHeader {
  char magic[4];
  uint32_t version, sample_rate;
  float frame_ms;
  uint32_t two_d, Kp, frame, hop, est_method;
  uint64_t est_floats;
};

constexpr uint32_t kStateVersion = 2;   // 2: strategy selectors, five band-state rows

struct StateScalars {
  Params prm;
  uint32_t params_loaded, block_count[4], available[4], profile_learned;
  int32_t med_write, tail_cur;
  uint32_t carry, was_learning;
  uint64_t dn_frames;
  float nlm_h;
  int32_t last_adaptive_state;
  float aggressiveness_state;
  int32_t gain_type, suppression_type, smoothing_type;
};

struct DevArray {
  float** ptr;
  size_t floats;
};

static std::vector<DevArray> state_arrays(Handle* h) {
  const GeoK& g = h->geo->k;
  const size_t Kp = g.Kp;
  std::vector<DevArray> a = {
      {&h->d_prof, 4 * Kp}, {&h->d_med_ring, (size_t)kMedianFrames * Kp},
      {&h->d_hist, (size_t)g.frame + g.hop}, {&h->d_tail[0], (size_t)g.frame},
      {&h->d_tail[1], (size_t)g.frame}, {&h->d_stable, Kp}, {&h->d_sp_prev, Kp},
      {&h->d_band_state, (size_t)kBandStateRows * kBandPitch}, {&h->d_floor, Kp}};
  if (h->two_d) {
    a.push_back({&h->d_X_hist_re, 4 * Kp});
    a.push_back({&h->d_X_hist_im, 4 * Kp});
    a.push_back({&h->d_noise_hist, 4 * Kp});
    a.push_back({&h->d_S_hist, (size_t)kSnrHist * Kp});
  }
  if (h->d_est) a.push_back({&h->d_est, h->est_floats});
  return a;
}

size_t state_size(Handle* h) {
  size_t n = sizeof(StateHeader) + sizeof(StateScalars) + 4;
  n += 4 * (size_t)h->profile_size * sizeof(float);
  for (const DevArray& d : state_arrays(h)) n += d.floats * sizeof(float);
  return n;
}

bool state_save(Handle* h, void* blob, size_t bytes) {
  std::lock_guard<std::recursive_mutex> engine_lock(engine_mutex());
  if (bytes < state_size(h)) {
    set_error_msg("specbleach-b200: state buffer too small (see specbleach_b200_state_size)");
    return false;
  }
  DeviceGuard restore_device;
  SB_CUDA(cudaSetDevice(h->device));
  SB_CUDA(cudaDeviceSynchronize());
  // load_noise_profile_for_mode / reset_noise_profile / profile_load since the last process():
  // the host mirror is newer than d_prof, and the blob must hold ONE profile, not two
  if (h->host_dirty) {
    const GeoK& g = h->geo->k;
    for (int m = 0; m < 4; m++)
      SB_CUDA(cudaMemcpy(h->d_prof + (size_t)m * g.Kp, h->prof_host[m].data(),
                         g.K * sizeof(float), cudaMemcpyHostToDevice));
    h->host_dirty = false;
  }
  unsigned char* p = static_cast<unsigned char*>(blob);
  StateHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, "SBST", 4);
  hd.version = kStateVersion;
  hd.sample_rate = h->geo->sample_rate;
  hd.frame_ms = h->geo->frame_ms;
  hd.two_d = h->two_d ? 1u : 0u;
  hd.Kp = (uint32_t)h->geo->k.Kp;
  hd.frame = (uint32_t)h->geo->k.frame;
  hd.hop = (uint32_t)h->geo->k.hop;
  hd.est_method = (uint32_t)h->est_method;
  hd.est_floats = h->d_est ? h->est_floats : 0;
  memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  StateScalars sc;
  memset(static_cast<void*>(&sc), 0, sizeof(sc));
  sc.prm = h->prm;
  sc.params_loaded = h->params_loaded;
  for (int m = 0; m < 4; m++) {
    sc.block_count[m] = h->block_count[m];
    sc.available[m] = h->available[m];
  }
  sc.profile_learned = h->profile_learned;
  sc.med_write = h->med_write;
  sc.tail_cur = h->tail_cur;
  sc.carry = h->carry;
  sc.was_learning = h->was_learning;
  sc.dn_frames = h->dn_frames;
  sc.nlm_h = h->nlm_h;
  sc.last_adaptive_state = h->last_adaptive_state;
  sc.aggressiveness_state = h->aggressiveness_state;
  sc.gain_type = h->gain_type;
  sc.suppression_type = h->suppression_type;
  sc.smoothing_type = h->smoothing_type;
  memcpy(p, &sc, sizeof(sc));
  p += sizeof(sc);
  for (int m = 0; m < 4; m++) {
    memcpy(p, h->prof_host[m].data(), (size_t)h->profile_size * sizeof(float));
    p += (size_t)h->profile_size * sizeof(float);
  }
  for (const DevArray& d : state_arrays(h)) {
    SB_CUDA(cudaMemcpy(p, *d.ptr, d.floats * sizeof(float), cudaMemcpyDeviceToHost));
    p += d.floats * sizeof(float);
  }
  const uint32_t c = crc32(blob, (size_t)(p - static_cast<unsigned char*>(blob)));
  memcpy(p, &c, 4);
  return true;
}

bool state_load(Handle* h, const void* blob, size_t bytes) {
  std::lock_guard<std::recursive_mutex> engine_lock(engine_mutex());
  const unsigned char* p = static_cast<const unsigned char*>(blob);
  StateHeader hd;
  if (bytes < sizeof(hd) + sizeof(StateScalars) + 4) {
    set_error_msg("specbleach-b200: state blob truncated");
    return false;
  }
  memcpy(&hd, p, sizeof(hd));
  if (memcmp(hd.magic, "SBST", 4) != 0 || hd.version != kStateVersion) {
    set_error_msg("specbleach-b200: not a state blob (magic / version)");
    return false;
  }
  if (hd.sample_rate != h->geo->sample_rate || hd.frame_ms != h->geo->frame_ms ||
      (hd.two_d != 0) != h->two_d || hd.Kp != (uint32_t)h->geo->k.Kp) {
    set_error_msg("specbleach-b200: state blob was saved for another denoiser geometry");
    return false;
  }
  // Validate the whole blob before the live handle is touched: a truncated or corrupt blob must
  // leave it exactly as it was.
  const GeoK& gk = h->geo->k;
  StateScalars sc;
  memcpy(&sc, p + sizeof(hd), sizeof(sc));
  const int new_method = (int)hd.est_method;     // -1 (0xffffffff): no estimator created yet
  if (new_method != -1 && new_method != kSpp && new_method != kBrandt && new_method != kMartin) {
    set_error_msg("specbleach-b200: state blob names an unknown noise estimator");
    return false;
  }
  const uint64_t want_est = (new_method == -1) ? 0 : (uint64_t)estimator_floats(gk, new_method);
  if (hd.est_floats != want_est || hd.frame != (uint32_t)gk.frame || hd.hop != (uint32_t)gk.hop) {
    set_error_msg("specbleach-b200: state blob does not fit this geometry (estimator size / frame / hop)");
    return false;
  }
  if (new_method == kBrandt && !brandt_supported(gk)) {
    set_error_msg("specbleach-b200: Brandt history too long for this frame size");
    return false;
  }
  size_t need = sizeof(hd) + sizeof(sc) + 4 + 4 * (size_t)h->profile_size * sizeof(float);
  {
    // the array list of the handle AS IT WILL BE: same as now except for the estimator
    float* keep = h->d_est;
    const size_t keep_n = h->est_floats;
    h->d_est = nullptr;
    for (const DevArray& d : state_arrays(h)) need += d.floats * sizeof(float);
    h->d_est = keep;
    h->est_floats = keep_n;
    need += (size_t)want_est * sizeof(float);
  }
  if (bytes < need) {
    set_error_msg("specbleach-b200: state blob truncated");
    return false;
  }
  uint32_t c;
  memcpy(&c, p + need - 4, 4);
  if (c != crc32(blob, need - 4)) {
    set_error_msg("specbleach-b200: state blob checksum mismatch");
    return false;
  }
  if (sc.med_write < 0 || sc.med_write >= kMedianFrames || sc.tail_cur < 0 || sc.tail_cur > 1 ||
      sc.carry >= (uint32_t)gk.hop || sc.gain_type < 0 || sc.gain_type > 2 || sc.suppression_type < 0 ||
      sc.suppression_type > 4 || sc.smoothing_type < 1 || sc.smoothing_type > 2) {
    set_error_msg("specbleach-b200: state blob holds out-of-range ring positions");
    return false;
  }
  DeviceGuard restore_device;
  SB_CUDA(cudaSetDevice(h->device));
  SB_CUDA(cudaDeviceSynchronize());
  // commit: the estimator of the snapshot decides the layout of d_est
  if (new_method != h->est_method || (want_est != 0) != (h->d_est != nullptr) ||
      (size_t)want_est != h->est_floats) {
    float* fresh = nullptr;
    if (want_est) SB_CUDA(dev_malloc(&fresh, (size_t)want_est * sizeof(float)));
    if (h->d_est) dev_free(h->d_est);
    h->d_est = fresh;
    h->est_floats = (size_t)want_est;
    h->est_method = new_method;
  }
  p += sizeof(hd) + sizeof(sc);
  h->prm = sc.prm;
  h->params_loaded = sc.params_loaded != 0;
  for (int m = 0; m < 4; m++) {
    h->block_count[m] = sc.block_count[m];
    h->available[m] = sc.available[m] != 0;
  }
  h->profile_learned = sc.profile_learned != 0;
  h->med_write = sc.med_write;
  h->tail_cur = sc.tail_cur;
  h->carry = sc.carry;
  h->was_learning = sc.was_learning != 0;
  h->dn_frames = sc.dn_frames;
  h->nlm_h = sc.nlm_h;
  h->last_adaptive_state = sc.last_adaptive_state;
  h->aggressiveness_state = sc.aggressiveness_state;
  h->gain_type = sc.gain_type;
  h->suppression_type = sc.suppression_type;
  h->smoothing_type = sc.smoothing_type;
  for (int m = 0; m < 4; m++) {
    memcpy(h->prof_host[m].data(), p, (size_t)h->profile_size * sizeof(float));
    p += (size_t)h->profile_size * sizeof(float);
  }
  for (const DevArray& d : state_arrays(h)) {
    SB_CUDA(cudaMemcpy(*d.ptr, p, d.floats * sizeof(float), cudaMemcpyHostToDevice));
    p += d.floats * sizeof(float);
  }
  h->host_dirty = false;          // d_prof came with the snapshot (state_save syncs it first)
  h->profile_dl_pending = false;
  h->noise_version++;             // cached tonal mask / whitening rows are stale
  for (int r = 0; r < 4; r++) h->noise_hist_version[r] = 0;
  return true;
}

// ---------------------------------------------------------------------------
// Interleaved PCM <-> planar float (libsndfile's conversions: int -> float divides by
// 2^(bits-1); float -> int multiplies by 2^(bits-1) - 1, rounds to nearest, saturates)
// ---------------------------------------------------------------------------
enum PcmFormat { kPcmF32 = 0, kPcmS16 = 1, kPcmS24 = 2, kPcmS32 = 3 };

static size_t pcm_bytes(int format) {
  return format == kPcmS16 ? 2 : format == kPcmS24 ? 3 : 4;
}

__global__ void __launch_bounds__(256)
deinterleave_kernel(const unsigned char* __restrict__ in, int format, uint32_t channels,
                    uint32_t frames, float* __restrict__ planar, size_t plane_pitch) {
  const size_t total = (size_t)frames * channels;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i % channels);
    const size_t f = i / channels;
    float v;
    if (format == kPcmF32) {
      v = reinterpret_cast<const float*>(in)[i];
    } else if (format == kPcmS16) {
      v = (float)reinterpret_cast<const short*>(in)[i] * (1.0f / 32768.0f);
    } else if (format == kPcmS24) {
      const unsigned char* p = in + 3 * i;
      int s = (int)p[0] | ((int)p[1] << 8) | ((int)(signed char)p[2] << 16);
      v = (float)s * (1.0f / 8388608.0f);
    } else {
      v = (float)reinterpret_cast<const int*>(in)[i] * (1.0f / 2147483648.0f);
    }
    planar[(size_t)c * plane_pitch + f] = v;
  }
}

__global__ void __launch_bounds__(256)
interleave_kernel(const float* __restrict__ planar, size_t plane_pitch, int format,
                  uint32_t channels, uint32_t frames, unsigned char* __restrict__ out) {
  const size_t total = (size_t)frames * channels;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i % channels);
    const size_t f = i / channels;
    const float v = planar[(size_t)c * plane_pitch + f];
    if (format == kPcmF32) {
      reinterpret_cast<float*>(out)[i] = v;
    } else if (format == kPcmS16) {
      float s = rintf(v * 32767.0f);
      s = fminf(fmaxf(s, -32768.0f), 32767.0f);
      reinterpret_cast<short*>(out)[i] = (short)s;
    } else if (format == kPcmS24) {
      float s = rintf(v * 8388607.0f);
      s = fminf(fmaxf(s, -8388608.0f), 8388607.0f);
      const int q = (int)s;
      unsigned char* p = out + 3 * i;
      p[0] = (unsigned char)(q & 0xff);
      p[1] = (unsigned char)((q >> 8) & 0xff);
      p[2] = (unsigned char)((q >> 16) & 0xff);
    } else {
      const double s = rint((double)v * 2147483647.0);
      reinterpret_cast<int*>(out)[i] =
          (int)(s < -2147483648.0 ? -2147483648.0 : (s > 2147483647.0 ? 2147483647.0 : s));
    }
  }
}

bool process_interleaved(Handle** hs, uint32_t channels, uint32_t frames, int format,
                         const void* in, void* out) {
  std::lock_guard<std::recursive_mutex> engine_lock(engine_mutex());
  if (!hs || channels == 0 || frames == 0 || !in || !out || format < 0 || format > 3) return false;
  for (uint32_t c = 0; c < channels; c++)
    if (!hs[c]) return false;
  DeviceGuard restore_device;
  SB_CUDA(cudaSetDevice(hs[0]->device));
  const size_t bytes = (size_t)frames * channels * pcm_bytes(format);
  const size_t pitch = ((size_t)frames + 31) & ~(size_t)31;
  unsigned char* d_pcm = nullptr;
  float* d_planar = nullptr;
  bool ok = true;
  if (dev_malloc(&d_pcm, bytes) != cudaSuccess ||
      dev_malloc(&d_planar, pitch * channels * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc", cudaGetLastError(), __FILE__, __LINE__);
    ok = false;
  }
  if (ok) {
    ok = cudaMemcpy(d_pcm, in, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    const size_t total = (size_t)frames * channels;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (ok) {
      deinterleave_kernel<<<blocks, 256>>>(d_pcm, format, channels, frames, d_planar, pitch);
      g_launches++;
      std::vector<uint32_t> n(channels, frames);
      std::vector<const float*> ins(channels);
      std::vector<float*> outs(channels);
      for (uint32_t c = 0; c < channels; c++) ins[c] = outs[c] = d_planar + (size_t)c * pitch;
      ok = cudaDeviceSynchronize() == cudaSuccess &&
           process_batch(hs, channels, n.data(), ins.data(), outs.data(), true);
    }
    if (ok) {
      interleave_kernel<<<blocks, 256>>>(d_planar, pitch, format, channels, frames, d_pcm);
      g_launches++;
      ok = cudaMemcpy(out, d_pcm, bytes, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (!ok && cudaPeekAtLastError() != cudaSuccess)
      set_error("process_interleaved", cudaGetLastError(), __FILE__, __LINE__);
  }
  dev_free(d_pcm);
  dev_free(d_planar);
  return ok;
}

}  // namespace sb

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

bool specbleach_b200_profile_save(void* instance, const char* path) {
  if (!instance || !path) return false;
  return sb::guarded(false, [&] {
    std::vector<unsigned char> buf;
    if (!sb::profile_serialize(static_cast<sb::Handle*>(instance), buf)) return false;
    FILE* f = fopen(path, "wb");
    if (!f) {
      sb::set_error_msg("specbleach-b200: cannot open noise profile file for writing");
      return false;
    }
    const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    return (fclose(f) == 0) && ok;
  });
}

bool specbleach_b200_profile_load(void* instance, const char* path) {
  if (!instance || !path) return false;
  return sb::guarded(false, [&] {
    FILE* f = fopen(path, "rb");
    if (!f) {
      sb::set_error_msg("specbleach-b200: cannot open noise profile file");
      return false;
    }
    // a profile file is a header plus four rows of at most fft_size floats: anything beyond a few
    // MB is not one, and is not read into memory
    constexpr size_t kMaxProfileFile = 64u << 20;
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0 && buf.size() <= kMaxProfileFile)
      buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    if (buf.size() > kMaxProfileFile) {
      sb::set_error_msg("specbleach-b200: noise profile file too large");
      return false;
    }
    return sb::profile_deserialize(static_cast<sb::Handle*>(instance), buf.data(), buf.size());
  });
}

size_t specbleach_b200_state_size(void* instance) {
  return instance ? sb::state_size(static_cast<sb::Handle*>(instance)) : 0;
}
bool specbleach_b200_state_save(void* instance, void* blob, size_t bytes) {
  if (!instance || !blob) return false;
  return sb::guarded(false, [&] { return sb::state_save(static_cast<sb::Handle*>(instance), blob, bytes); });
}
bool specbleach_b200_state_load(void* instance, const void* blob, size_t bytes) {
  if (!instance || !blob) return false;
  return sb::guarded(false, [&] { return sb::state_load(static_cast<sb::Handle*>(instance), blob, bytes); });
}

bool specbleach_b200_process_interleaved(void** instances, uint32_t channels, uint32_t frames,
                                         int format, const void* input, void* output) {
  return sb::guarded(false, [&] {
    return sb::process_interleaved(reinterpret_cast<sb::Handle**>(instances), channels, frames, format, input,
                                   output);
  });
}

}  // extern "C"
