// Engine-level declarations shared by sb_engine.cu and sb_api.cu.
#pragma once

#include "sb_common.h"

#include <exception>
#include <map>
#include <mutex>
#include <vector>

namespace sb {

struct Context {
  int chunk_frames = 0;      // frames per internal sub-call; 0 = automatic (sub_call_frames)
  int num_lanes = 4;
  std::map<int, std::vector<Lane>> lanes;   // per device: streams and scratch belong to one GPU
  // instrumentation
  double nlm_ms = 0.0;
  uint64_t nlm_launches = 0;
  uint64_t nlm_frame_blocks = 0;
  uint64_t frames = 0;
  double class_ms[kcCount] = {0};
  uint64_t class_launches[kcCount] = {0};
};
Context& ctx();
std::recursive_mutex& engine_mutex();   // serialises device work of concurrent host threads
int sub_call_frames(const GeoK& g);   // chunk_frames, or its per-geometry default

void set_error_msg(const char* msg);

// Entry points select the handle's device; the caller's current device is put back on return
// (a library call must not change it behind the host program's back).
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

Handle* handle_create(bool two_d, uint32_t sample_rate, float frame_ms);
void handle_destroy(Handle* h);
bool handle_load_parameters(Handle* h, const Params& p);
bool handle_load_profile(Handle* h, const float* prof, uint32_t size, uint32_t blocks, int mode);
void handle_reset_profile(Handle* h);

Lane* get_lane(int idx, const GeoK& g);
bool process_call(Handle* h, uint32_t n, const float* in, float* out, bool device_io);
bool process_batch(Handle** hs, uint32_t count, const uint32_t* n, const float* const* in,
                   float* const* out, bool device_io);
bool process_tiled(Handle* h, uint32_t n, const float* in, float* out, uint32_t tiles,
                   uint32_t warmup_frames, const int* devices, uint32_t device_count, uint32_t flags);
bool process_tiled_batch(Handle** hs, uint32_t count, const uint32_t* ns, const float* const* ins,
                         float* const* outs, uint32_t tiles, uint32_t warmup_frames, const int* devices,
                         uint32_t device_count, uint32_t flags);
bool stats_collect(Context& c);

// sb_io.cu: the streaming state of a handle as one host blob
size_t state_size(Handle* h);
bool state_save(Handle* h, void* blob, size_t bytes);
bool state_load(Handle* h, const void* blob, size_t bytes);

void set_error_msg(const char* msg);

// No C++ exception may cross the C boundary (std::bad_alloc from a host-side vector, a
// std::system_error from a mutex): the call fails the way the reference's do -- NULL / false --
// and specbleach_b200_last_error() says why.
template <typename R, typename F>
R guarded(R on_error, F&& f) noexcept {
  try {
    return f();
  } catch (const std::exception& e) {
    set_error_msg(e.what());
  } catch (...) {
    set_error_msg("specbleach-b200: unknown C++ exception");
  }
  return on_error;
}

// sb_geometry.cu: private geometry for the FFT-only debug entry point
Geometry* make_fft_geometry(int fft_size);
void free_geometry(Geometry* G);

}  // namespace sb
