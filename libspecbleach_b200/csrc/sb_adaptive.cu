// Time-recursive adaptive noise estimators as one-thread-per-bin scans over the
// frames of a call.  The frame-level control flow of the reference (silence
// gate, first-frame snap, Martin's sub-window counters) depends only on the
// per-frame mean power, which is reduced first; every thread then replays the
// same scalar state machine, so bins stay independent.
//
// Reference functions restated here:
//   spp_mmse_noise_estimator_run / _update_seed / _apply_floor
//        spp_mmse_noise_estimator.c:28-62,98-194
//   martin_noise_estimator_run / _set_state / _apply_floor
//        martin_noise_estimator.c:82-209
//   brandt_noise_estimator_*: sb_brandt.cu
//   adaptive_estimator_run tail (gap interpolation + smoothing of the OUTPUT)
//        adaptive_noise_estimator.c:156-187   (done by interp_smooth_kernel)
//   denoiser_profile_core_update adaptive branch  denoiser_profile_core.c:55-88
#include "sb_common.h"

#include <float.h>

namespace sb {

// estimator state layout in Handle::d_est (floats):
//   [0] first-frame flag (1 = still waiting for the first non-silent frame)
//   [1] Martin frame_count   [2] Martin subwin_index
//   [8..10] the same three values as they stand AFTER the scan in flight: every CTA of a scan
//        reads [0..2] when it starts, so the scan must not overwrite them (a CTA scheduled after
//        CTA 0 has retired would see the next call's flag and counters); thread 0 leaves the new
//        values here and header_commit_kernel moves them to [0..2] once the scan has finished
//   [32 ...] per-bin arrays, pitch Kp:
//        SPP:    psd, smoothed_spp
//        Martin: psd, cur_min, hist[8]
constexpr int kEstHeader = 32;
constexpr int kEstNext = 8;
constexpr int kBatch = 16;   // frames whose loads are in flight per thread in the scans

namespace {

__global__ void __launch_bounds__(256)
frame_energy_kernel(const GeoK g, const float* __restrict__ P, float* __restrict__ energy) {
  __shared__ float red[256];
  const int f = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < g.K; k += blockDim.x) s += P[(size_t)f * g.Kp + k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) energy[f] = red[0] / (float)g.K;
}

__global__ void seed_kernel(const GeoK g, int method, const float* __restrict__ floorv,
                            float* __restrict__ est) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0 && method == kMartin) {
    est[0] = 0.f;   // set_state clears is_first_frame and frame_count (martin:184-187)
    est[1] = 0.f;
  }
  if (k >= g.K) return;
  float* a = est + kEstHeader;
  const float val = fmaxf(floorv[k], FLT_MIN);
  if (method == kSpp) {
    a[k] = val;            // spp:169-180 (is_first_frame NOT cleared)
    a[g.Kp + k] = 0.f;
  } else {
    const float v = val / kMartinBias;
    a[k] = v;
    a[g.Kp + k] = v;
    for (int d = 0; d < kMartinSubwinCount; d++) a[(size_t)(2 + d) * g.Kp + k] = v;
  }
}

__device__ __forceinline__ float spp_probability(float obs, float prev) {
  // spp_mmse_noise_estimator.c:36-56
  if (prev < kSpectralEps) prev = kSpectralEps;
  const float ratio = obs / prev;
  const float exponent = -ratio * (kSppXi / (1.f + kSppXi));
  float e = expf(exponent);
  if (!isfinite(e)) e = (exponent > 0.f) ? FLT_MAX : 0.f;
  const float den = (1.f + kSppXi) * e;
  float spp = 1.f / (1.f + den);
  return fmaxf(0.f, fminf(1.f, spp));
}

__global__ void spp_scan_kernel(const GeoK g, const float* __restrict__ P,
                                const float* __restrict__ energy, int frames,
                                const float* __restrict__ floorv, float* __restrict__ est,
                                float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.K) return;
  float* a = est + kEstHeader;
  float psd = a[k], sspp = a[g.Kp + k];
  bool first = est[0] != 0.f;
  const float fl = floorv[k];
  // the loads do not depend on the recurrence: fetch kBatch frames, then run the chain
  for (int f0 = 0; f0 < frames; f0 += kBatch) {
    float pb[kBatch], eb[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; i++) {
      const bool in = f0 + i < frames;
      pb[i] = in ? P[(size_t)(f0 + i) * g.Kp + k] : 0.f;
      eb[i] = in ? energy[f0 + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kBatch; i++) {
    const int f = f0 + i;
    if (f >= frames) break;
    const size_t o = (size_t)f * g.Kp + k;
    const float p = pb[i];
    const bool silent = eb[i] < kSilenceThreshold;
    float res;
    if (first) {
      if (silent) {
        res = 0.f;
      } else {
        psd = p; sspp = 0.f; res = p; first = false;
      }
    } else if (silent) {
      res = psd;
    } else {
      float s1 = spp_probability(p, psd);
      if (sspp > kSppCap) s1 = fminf(s1, kSppCap);
      const float s0 = 1.f - s1;
      const float mmse = (s0 * p) + (s1 * psd);
      res = (kSppAlphaPow * psd) + ((1.f - kSppAlphaPow) * mmse);
      sspp = (kSppSmooth * sspp) + (kSppCurrent * s1);
      psd = res;
    }
    out[o] = res;
    if (psd < fl) psd = fl;   // apply_floor, every frame (spp:182-194)
    }
  }
  a[k] = psd;
  a[g.Kp + k] = sspp;
  if (k == 0) {
    est[kEstNext + 0] = first ? 1.f : 0.f;
    est[kEstNext + 1] = est[1];
    est[kEstNext + 2] = est[2];
  }
}

__global__ void header_commit_kernel(float* __restrict__ est) {
  if (threadIdx.x < 3) est[threadIdx.x] = est[kEstNext + threadIdx.x];
}

__global__ void martin_scan_kernel(const GeoK g, const float* __restrict__ P,
                                   const float* __restrict__ energy, int frames,
                                   const float* __restrict__ floorv, float* __restrict__ est,
                                   float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.K) return;
  float* a = est + kEstHeader;
  float psd = a[k], cmin = a[g.Kp + k];
  float hist[kMartinSubwinCount];
#pragma unroll
  for (int d = 0; d < kMartinSubwinCount; d++) hist[d] = a[(size_t)(2 + d) * g.Kp + k];
  bool first = est[0] != 0.f;
  unsigned fc = (unsigned)est[1];
  unsigned idx = (unsigned)est[2];
  const float fl = floorv[k];
  for (int f0 = 0; f0 < frames; f0 += kBatch) {
    float pb[kBatch], eb[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; i++) {
      const bool in = f0 + i < frames;
      pb[i] = in ? P[(size_t)(f0 + i) * g.Kp + k] : 0.f;
      eb[i] = in ? energy[f0 + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kBatch; i++) {
    const int f = f0 + i;
    if (f >= frames) break;
    const size_t o = (size_t)f * g.Kp + k;
    const float p = pb[i];
    const bool silent = eb[i] < kSilenceThreshold;
    float res;
    bool emitted = false;
    if (first) {
      if (silent) {
        res = 0.f;
      } else {
        const float v = p / kMartinBias;
        psd = v; cmin = v;
#pragma unroll
        for (int d = 0; d < kMartinSubwinCount; d++) hist[d] = v;
        res = p;
        first = false;
        fc = 1;
      }
      emitted = true;   // both branches return before calculate_output (martin:101-122)
    } else if (!silent) {
      psd = (kMartinAlpha * psd) + ((1.0f - kMartinAlpha) * p);
      if (psd < cmin) cmin = psd;
      if (fc >= (unsigned)kMartinSubwinLen) {
#pragma unroll
        for (int d = 0; d < kMartinSubwinCount; d++)
          if ((unsigned)d == idx) hist[d] = cmin;
        cmin = psd;
        idx = (idx + 1) % kMartinSubwinCount;
        fc = 0;
      }
    }
    if (!emitted) {
      float mv = cmin;
#pragma unroll
      for (int d = 0; d < kMartinSubwinCount; d++)
        if (hist[d] < mv) mv = hist[d];
      res = mv * kMartinBias;
      fc++;
    }
    out[o] = res;
    // apply_floor on the whole state, every frame (martin:194-209)
    if (psd < fl) psd = fl;
    if (cmin < fl) cmin = fl;
#pragma unroll
    for (int d = 0; d < kMartinSubwinCount; d++)
      if (hist[d] < fl) hist[d] = fl;
    }
  }
  a[k] = psd;
  a[g.Kp + k] = cmin;
#pragma unroll
  for (int d = 0; d < kMartinSubwinCount; d++) a[(size_t)(2 + d) * g.Kp + k] = hist[d];
  if (k == 0) {
    est[kEstNext + 0] = first ? 1.f : 0.f;
    est[kEstNext + 1] = (float)fc;
    est[kEstNext + 2] = (float)idx;
  }
}

}  // namespace

size_t estimator_floats(const GeoK& g, int method) {
  if (method == kBrandt) return brandt_estimator_floats(g);
  const size_t per_bin = (method == kSpp) ? 2 : (2 + kMartinSubwinCount);
  return kEstHeader + per_bin * (size_t)g.Kp;
}

bool launch_estimator_seed(const GeoK& g, int method, const float* floorv, float* est,
                           cudaStream_t st) {
  if (method == kBrandt) return launch_brandt_seed(g, floorv, est, st);
  seed_kernel<<<(g.K + 255) / 256, 256, 0, st>>>(g, method, floorv, est);
  SB_LAUNCH_CHECK();
  return true;
}

// Produces the per-frame noise rows of a call: estimator scan -> gap
// interpolation + smoothing of the output -> clamp to the manual floor.
bool launch_adaptive(const GeoK& g, int method, const float* P, int frames, const float* floorv,
                     float* est, float* energy, int* idx_scratch, bool fresh_seed,
                     float* noise_rows, cudaStream_t st) {
  KTimer kt_(kcAdaptive, st);
  if (frames <= 0) return true;
  frame_energy_kernel<<<frames, 256, 0, st>>>(g, P, energy);
  SB_LAUNCH_CHECK();
  if (method == kBrandt) {
    if (!launch_brandt(g, P, frames, floorv, est, energy, idx_scratch, fresh_seed, noise_rows, st))
      return false;
  } else if (method == kSpp) {
    spp_scan_kernel<<<(g.K + 63) / 64, 64, 0, st>>>(g, P, energy, frames, floorv, est, noise_rows);
  } else {
    martin_scan_kernel<<<(g.K + 63) / 64, 64, 0, st>>>(g, P, energy, frames, floorv, est,
                                                       noise_rows);
  }
  if (method != kBrandt) {
    SB_LAUNCH_CHECK();
    header_commit_kernel<<<1, 32, 0, st>>>(est);
    SB_LAUNCH_CHECK();
  }
  return launch_interp_smooth(g, noise_rows, g.Kp, frames, 0xffffffffu, kInterpThreshold,
                              kNoiseSmoothing, floorv, st);
}

}  // namespace sb
