// Manual noise-profile learning and the profile clean-up pass.
//
// Reference functions restated here (whole-batch form):
//   noise_estimation_run x4        noise_estimator.c:83-141
//   get_rolling_mean_spectrum      spectral_utils.c:126-144
//   get_rolling_median_spectrum    spectral_utils.c:167-196  (25-frame window)
//   max_spectrum / min_spectrum    spectral_utils.c:79-103
//   noise_estimation_finalize      noise_estimator.c:143-162
//   interpolate_spectrum_gaps      spectral_utils.c:224-251
//   smooth_spectrum                spectral_utils.c:198-222
#include "sb_common.h"

namespace sb {
namespace {

// One thread per bin walks the learn frames of this call in order.
// prof = [mean | median | max | min] rows of pitch Kp; only the state after the
// last frame is observable (the profile is read between process() calls).
__global__ void learn_kernel(const GeoK g, const float* __restrict__ P, int frames,
                             unsigned count0, float* __restrict__ prof,
                             float* __restrict__ ring, int wr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.K) return;
  float mean = prof[k];
  float mx = prof[2 * g.Kp + k];
  float mn = prof[3 * g.Kp + k];
  unsigned c = count0;
  for (int f = 0; f < frames; f++) {
    const float p = P[(size_t)f * g.Kp + k];
    // the block count is read BEFORE it is incremented (noise_estimator.c:94-100):
    // the first two frames both copy
    if (c <= 1u) mean = p;
    else mean += (p - mean) / (float)c;
    c++;
    mx = fmaxf(mx, p);
    mn = fminf(mn, p);
  }
  prof[k] = mean;
  prof[2 * g.Kp + k] = mx;
  prof[3 * g.Kp + k] = mn;
  // 25-frame median history (zero-filled until 25 frames were learned)
  const int first = frames > kMedianFrames ? frames - kMedianFrames : 0;
  for (int f = first; f < frames; f++)
    ring[(size_t)((wr + f) % kMedianFrames) * g.Kp + k] = P[(size_t)f * g.Kp + k];
  float v[kMedianFrames];
#pragma unroll
  for (int i = 0; i < kMedianFrames; i++) v[i] = ring[(size_t)i * g.Kp + k];
  // insertion sort, then the middle element (odd count: spectral_utils.c:153-165)
  for (int i = 1; i < kMedianFrames; i++) {
    const float key = v[i];
    int j = i - 1;
    while (j >= 0 && v[j] > key) { v[j + 1] = v[j]; j--; }
    v[j + 1] = key;
  }
  prof[g.Kp + k] = v[kMedianFrames / 2];
}

// Gap interpolation followed by the 3-tap smoothing, one CTA per row.
// A bin is a gap when value < threshold; gaps strictly inside (0, K-1) that are
// closed by a valid bin are linearly interpolated between the bin before the
// run and the closing bin.  smooth3 then reads the interpolated values.
constexpr int kISThreads = 256;

__global__ void __launch_bounds__(kISThreads)
interp_smooth_kernel(const GeoK g, float* __restrict__ rows, int row_stride, unsigned enable,
                     int nrows, float threshold, float factor, const float* __restrict__ floorv) {
  extern __shared__ float sm[];
  const int K = g.K;
  float* v = sm;                                  // [K] original
  float* t = sm + K;                              // [K] interpolated
  int* Lidx = reinterpret_cast<int*>(sm + 2 * K); // [K] anchor at or before k
  int* Ridx = Lidx + K;                           // [K] valid bin at or after k (K if none)
  __shared__ int cl[kISThreads], cr[kISThreads];
  const int row = blockIdx.x;
  if (nrows <= 32 && !((enable >> row) & 1u)) return;
  float* __restrict__ r = rows + (size_t)row * row_stride;
  for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = r[k];
  __syncthreads();
  const int CH = (K + kISThreads - 1) / kISThreads;
  const int k0 = threadIdx.x * CH;
  const int k1 = (k0 + CH < K) ? k0 + CH : K;
  {
    int last = -1, first = K;
    for (int k = k0; k < k1; k++) {
      const bool valid = !(v[k] < threshold) || k == 0;
      if (valid) last = k;
    }
    for (int k = k1 - 1; k >= k0; k--) {
      const bool valid = !(v[k] < threshold);
      if (valid) first = k;
    }
    cl[threadIdx.x] = last;
    cr[threadIdx.x] = first;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;   // bin 0 is always an anchor
    for (int c = 0; c < kISThreads; c++) {
      const int own = cl[c];
      cl[c] = run;            // anchor carried INTO chunk c
      if (own >= 0) run = own;
    }
  } else if (threadIdx.x == 32) {
    int run = K;
    for (int c = kISThreads - 1; c >= 0; c--) {
      const int own = cr[c];
      cr[c] = run;            // first valid bin AFTER chunk c
      if (own < K) run = own;
    }
  }
  __syncthreads();
  {
    int run = cl[threadIdx.x];
    for (int k = k0; k < k1; k++) {
      if (!(v[k] < threshold) || k == 0) run = k;
      Lidx[k] = run;
    }
    run = cr[threadIdx.x];
    for (int k = k1 - 1; k >= k0; k--) {
      if (!(v[k] < threshold)) run = k;
      Ridx[k] = run;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float x = v[k];
    if (K >= 3 && k >= 1 && k <= K - 2 && x < threshold) {
      const int p = Lidx[k - 1];
      const int j = Ridx[k];
      if (j < K) {
        const float sv = v[p], ev = v[j];
        const float step = (ev - sv) / (float)(j - p);
        x = sv + (step * (float)(k - p));
      }
    }
    t[k] = x;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float x = t[k];
    if (K >= 2 && factor > 0.0f) {
      if (k == 0) x = ((t[0] + t[1]) / 2.0f * factor) + (t[0] * (1.0f - factor));
      else if (k == K - 1) x = (((t[K - 2] + t[K - 1]) / 2.0f) * factor) + (t[K - 1] * (1.0f - factor));
      else x = (((t[k - 1] + t[k] + t[k + 1]) / 3.0f) * factor) + (t[k] * (1.0f - factor));
    }
    if (floorv) {   // denoiser_profile_core.c:83-87
      const float fl = floorv[k];
      if (x < fl) x = fl;
    }
    r[k] = x;
  }
}

}  // namespace

bool launch_learn(const GeoK& g, const float* P, int frames, uint32_t count0, float* prof,
                  float* med_ring, int med_write, cudaStream_t st) {
  KTimer kt_(kcLearn, st);
  if (frames <= 0) return true;
  learn_kernel<<<(g.K + 127) / 128, 128, 0, st>>>(g, P, frames, count0, prof, med_ring, med_write);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_interp_smooth(const GeoK& g, float* rows, int row_stride, int nrows,
                          unsigned enable_mask, float threshold, float factor,
                          const float* floorv, cudaStream_t st) {
  KTimer kt_(kcInterp, st);
  if (nrows <= 0) return true;
  const size_t smem = sizeof(float) * 4 * (size_t)g.K;
  if (smem > 48 * 1024)
    SB_CUDA(cudaFuncSetAttribute((const void*)interp_smooth_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  interp_smooth_kernel<<<nrows, kISThreads, smem, st>>>(g, rows, row_stride, enable_mask, nrows,
                                                        threshold, factor, floorv);
  SB_LAUNCH_CHECK();
  return true;
}

}  // namespace sb
