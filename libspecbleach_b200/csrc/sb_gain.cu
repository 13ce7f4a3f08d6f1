// Gain chain: everything between the (NLM-smoothed) spectrum and the gains that
// multiply the delayed FFT frame.  Frame-parallel parts are (frame, bin) kernels
// with one CTA per frame row and a thread per 4 bins, and (frame, band) kernels
// with one CTA per frame; the three time recursions of the masking veto are
// sequential scans over frames staged through shared memory (rowscan_kernel).
//
// Reference functions restated here (whole-batch form):
//   get_morphed_profile            spectral_utils.c:253-279
//   nlm_filter_calculate_snr       nlm_filter.c:453-479
//   nlm_filter_reconstruct_magnitude nlm_filter.c:481-508
//   calculate_berouti_per_bin      suppression_engine.c:108-133
//   tonal_reducer_run / detect_tonal_components  tonal_reducer.c:54-114, tonal_detector.c:24-162
//   spectral_smoothing_run (FIXED) spectral_smoother.c:111-139,183-192   (1D)
//   masking_veto_apply             masking_veto.c:127-267
//   compute_masking_thresholds     masking_estimator.c:188-349
//   apply_thresholds_as_floor      absolute_hearing_thresholds.c:141-159
//   wiener_subtraction             gain_calculator.c:27-69
//   spectral_whitening_get_weights spectral_whitening.c:59-111
//   noise_floor_manager_apply      noise_floor_manager.c:77-163
//   denoiser_post_process_apply    denoiser_post_process.c:26-51
#include "sb_common.h"

#include <float.h>

namespace sb {
namespace {

constexpr float kInvPower = 1.0f / kPowerLawExp;
constexpr float kInvAdd = 1.0f / kAdditivityExp;

__global__ void morph_kernel(const GeoK g, const float* __restrict__ prof, float a,
                             float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.Kp) return;
  float v = 0.f;
  if (k < g.K) {
    const float mean = prof[k], med = prof[g.Kp + k], mx = prof[2 * g.Kp + k],
                mn = prof[3 * g.Kp + k];
    if (a < 0.0f) {
      const float t = -a;
      v = (mean * (1.0f - t)) + (med * t);
    } else {
      v = (mean * (1.0f - a)) + (mx * a);
    }
    v = fmaxf(v, mn);
  }
  out[k] = v;
}

__global__ void fill_rows_kernel(const GeoK g, const float* __restrict__ row,
                                 float* __restrict__ dst, int rows) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.Kp) return;
  const float v = row[k];
  for (int r = blockIdx.y; r < rows; r += gridDim.y) dst[(size_t)r * g.Kp + k] = v;
}

// The (frame, bin) elementwise kernels below run one CTA per frame row, a thread per group
// of 4 bins (128-bit loads and stores; rows are Kp = 4-aligned floats from cudaMalloc), CTA
// size = the row's quads rounded up to a warp (row_threads): a thread per bin on a 2D grid
// reached ~45 % of the HBM rate, with a quarter of the lanes past the end of the row.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float snr_one(float p, float n, bool in) {
  const float den = n > kNlmNoiseFloorMin ? n : kNlmNoiseFloorMin;
  return in ? p / den : 0.f;
}

__global__ void __launch_bounds__(512)
snr_kernel(const GeoK g, const float* __restrict__ P, const float* __restrict__ noise,
           float* __restrict__ S) {
  const size_t ro = (size_t)blockIdx.x * g.Kp;
  for (int k = 4 * threadIdx.x; k < g.Kp; k += 4 * blockDim.x) {
    const float4 p = ld4(P + ro + k), n = ld4(noise + ro + k);
    float4 v;
    v.x = snr_one(p.x, n.x, k < g.K);
    v.y = snr_one(p.y, n.y, k + 1 < g.K);
    v.z = snr_one(p.z, n.z, k + 2 < g.K);
    v.w = snr_one(p.w, n.w, k + 3 < g.K);
    st4(S + ro + k, v);
  }
}

// 51-bin edge-clamped median over frequency, then the tonal mask
// (tonal_detector.c:24-162).  A thread walks a segment of 32 consecutive bins of one
// row and keeps the 51-value window SORTED in shared memory (column layout, conflict
// free): the first window is insertion-sorted, every next one replaces the value that
// leaves by the value that enters with one shifting pass, so a bin costs O(51) instead
// of the O(51^2) of a selection per bin (3.4 ms -> 0.1 ms per 4800 frames when the
// noise estimate changes every frame).  median = sorted[25].
constexpr int kTmSeg = 32;       // bins per thread
constexpr int kTmThreads = 128;

__global__ void __launch_bounds__(kTmThreads)
tonal_mask_kernel(const GeoK g, const float* __restrict__ det, int det_stride,
                  float* __restrict__ mask, int mask_stride, int rows, int segs) {
  __shared__ float win[kTonalWindow][kTmThreads];
  const int id = blockIdx.x * kTmThreads + threadIdx.x;
  if (id >= rows * segs) return;
  const int row = id / segs;
  const int k0 = (id - row * segs) * kTmSeg;
  const float* __restrict__ d = det + (size_t)row * det_stride;
  float* __restrict__ mo = mask + (size_t)row * mask_stride;
  const int half = kTonalWindow / 2;
  const int K = g.K;
  const int t = threadIdx.x;
  auto at = [&](int idx) { return __ldg(&d[idx < 0 ? 0 : (idx > K - 1 ? K - 1 : idx)]); };
  // insertion sort of the first window: bins k0-25 .. k0+25 (clamped)
  for (int i = 0; i < kTonalWindow; i++) {
    const float v = at(k0 - half + i);
    int p = i;
    while (p > 0 && win[p - 1][t] > v) {
      win[p][t] = win[p - 1][t];
      p--;
    }
    win[p][t] = v;
  }
  const int kend = (k0 + kTmSeg < g.Kp) ? k0 + kTmSeg : g.Kp;
  for (int k = k0; k < kend; k++) {
    float m = 0.f;
    if (k < K) {
      const float med = win[half][t];
      const float peak = at(k);
      if (!(peak - med < kMinPeakProminence)) {
        const float ratio = peak / (med + 1e-20f);
        if (ratio > kPeakThreshold) m = (peak - med) / (peak + 1e-20f);
      }
    }
    mo[k] = m;
    if (k + 1 < kend && k + 1 < K) {
      // slide: bin k-25 leaves, bin k+26 enters
      const float old = at(k - half), nw = at(k + 1 + half);
      if (old != nw) {
        int lo = 0, hi = kTonalWindow - 1;           // first position holding `old`
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (win[mid][t] < old) lo = mid + 1;
          else hi = mid;
        }
        int p = lo;
        if (nw > old) {
          while (p < kTonalWindow - 1 && win[p + 1][t] < nw) {
            win[p][t] = win[p + 1][t];
            p++;
          }
        } else {
          while (p > 0 && win[p - 1][t] > nw) {
            win[p][t] = win[p - 1][t];
            p--;
          }
        }
        win[p][t] = nw;
      }
    }
  }
}

// Few rows (the static manual profile: one row per parameter change): selection per bin,
// one thread per bin -- more parallel than the sliding window when there is little to do.
// grid = (bin tiles of 256, rows); each CTA stages its bins plus a 25-bin halo.
__global__ void __launch_bounds__(256)
tonal_mask_direct_kernel(const GeoK g, const float* __restrict__ det, int det_stride,
                  float* __restrict__ mask, int mask_stride) {
  __shared__ float sm[256 + kTonalWindow];
  const int row = blockIdx.y;
  const float* __restrict__ d = det + (size_t)row * det_stride;
  const int half = kTonalWindow / 2;
  const int K = g.K;
  const int k0 = blockIdx.x * 256;
  for (int e = threadIdx.x; e < 256 + 2 * half; e += blockDim.x) {
    int idx = k0 + e - half;
    idx = idx < 0 ? 0 : (idx > K - 1 ? K - 1 : idx);
    sm[e] = d[idx];
  }
  __syncthreads();
  const int k = k0 + threadIdx.x;
  if (k >= g.Kp) return;
  float m = 0.f;
  if (k < K) {
    const float* w = &sm[threadIdx.x];   // window = w[0..50]
    float med = w[half];
    for (int i = 0; i < kTonalWindow; i++) {
      const float v = w[i];
      int less = 0, leq = 0;
#pragma unroll
      for (int j = 0; j < kTonalWindow; j++) {
        const float u = w[j];
        less += (u < v);
        leq += (u <= v);
      }
      if (less <= half && half < leq) { med = v; break; }
    }
    const float peak = w[half];
    if (!(peak - med < kMinPeakProminence)) {
      const float ratio = peak / (med + 1e-20f);
      if (ratio > kPeakThreshold) m = (peak - med) / (peak + 1e-20f);
    }
  }
  mask[(size_t)row * mask_stride + k] = m;
}

// Whitening weights (spectral_whitening.c:59-111): smooth3(noise) -> median = sorted[K/2] ->
// power-law weights.  The reference qsorts the K values for the one order statistic it needs;
// here it is a 4-pass radix select on the float bit patterns (order-preserving for the
// non-negative smoothed noise; negative / NaN inputs would order as their raw bits), exact
// and O(K) instead of the 66 barrier-separated stages of a bitonic sort.
__global__ void __launch_bounds__(512)
whitening_kernel(const GeoK g, const float* __restrict__ noise, int noise_stride, float whit,
                 float* __restrict__ wt, int wt_stride) {
  extern __shared__ float sm[];
  float* smo = sm;                                   // [K]
  __shared__ unsigned hist[256];
  __shared__ unsigned sel_prefix, sel_rank;
  const int row = blockIdx.x;
  const float* __restrict__ n = noise + (size_t)row * noise_stride;
  const int K = g.K;
  const float f = 0.5f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float c = n[k];
    float v;
    if (K < 2) v = c;
    else if (k == 0) v = ((c + n[1]) / 2.0f * f) + (c * (1.0f - f));
    else if (k == K - 1) v = (((n[K - 2] + c) / 2.0f) * f) + (c * (1.0f - f));
    else v = (((n[k - 1] + c + n[k + 1]) / 3.0f) * f) + (c * (1.0f - f));
    smo[k] = v;
  }
  if (threadIdx.x == 0) {
    sel_prefix = 0u;
    sel_rank = (unsigned)(K / 2);
  }
  __syncthreads();
  for (int pass = 3; pass >= 0; pass--) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned prefix = sel_prefix;
    const int shift = 8 * pass;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const unsigned bits = __float_as_uint(smo[k]);
      // keys agreeing with the digits already fixed (the higher passes)
      if (pass == 3 || (bits >> (shift + 8)) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // bucket whose cumulative count crosses the wanted rank: 8 buckets per lane + warp scan
      const unsigned lane = threadIdx.x;
      unsigned local[8], tot = 0u;
#pragma unroll
      for (int i = 0; i < 8; i++) { local[i] = hist[8 * lane + i]; tot += local[i]; }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
      }
      const unsigned before = incl - tot, r = sel_rank;
      if (r >= before && r < incl) {                 // exactly one lane
        unsigned acc = before;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (r >= acc && r < acc + local[i]) {
            sel_prefix = (prefix << 8) | (8u * lane + (unsigned)i);
            sel_rank = r - acc;
          }
          acc += local[i];
        }
      }
    }
    __syncthreads();
  }
  float anchor = __uint_as_float(sel_prefix);
  if (anchor < kSpectralEps) anchor = kSpectralEps;
  for (int k = threadIdx.x; k < g.Kp; k += blockDim.x) {
    float w = 1.0f;
    if (k < K && whit > 0.0f && smo[k] > kSpectralEps) w = powf(anchor / smo[k], whit);
    wt[(size_t)row * wt_stride + k] = w;
  }
}

// ---------------------------------------------------------------------------
// (frame, bin): magnitude, Berouti alpha, tonal boost, clean-signal estimate
// ---------------------------------------------------------------------------
struct PreArgs {
  int first_active, two_d, mask_stride;
  float alpha_max_user, tonal_gain;
  int suppression;      // SuppressionType of the reference (suppression_engine.h:39-46), 0 = Berouti per bin
};

// one bin; *mg and *cl are only meaningful in 2D.  al_band: the alpha of the bin's band / frame
// for the strategies that do not work per bin (N4: suppression_engine.c:135-241).
__device__ __forceinline__ float pre_one(const PreArgs& a, float refv, float n, float m, float* mg,
                                         float* cl, float al_band) {
  float spec;
  if (a.two_d) {
    const float den = n > kNlmNoiseFloorMin ? n : kNlmNoiseFloorMin;
    spec = refv * den;               // nlm_filter.c:481-508 (ref = smoothed SNR)
    *mg = spec;
    *cl = fmaxf(spec - n, 0.0f);
  } else {
    spec = refv;                     // 1D: alpha from the UNsmoothed power (spectral_denoiser.c:298-300)
  }
  float al;
  if (a.suppression == 0) {
    // suppression_engine.c:108-133
    const float snr_db = 10.0f * log10f(spec / (n + kSpectralEps));
    if (snr_db <= kSnrLowDb) al = a.alpha_max_user;
    else if (snr_db >= kSnrHighDb) al = kAlphaMin;
    else {
      const float ns = (snr_db - kSnrLowDb) / (kSnrHighDb - kSnrLowDb);
      al = a.alpha_max_user - (ns * (a.alpha_max_user - kAlphaMin));
    }
  } else if (a.suppression == 4) {
    al = kAlphaMin;        // SUPPRESSION_NONE (suppression_engine.c:269-274)
  } else {
    al = al_band;          // global SNR / critical-band SNR / masking thresholds: computed upstream
  }
  // tonal_reducer.c:92-113
  if (!(a.tonal_gain >= 0.999f)) {
    if (m > 0.0f) {
      const float strength = 1.0f - a.tonal_gain;
      const float alpha_needed = kAlphaMin + (strength * (kAlphaMaxTonal - kAlphaMin));
      const float target_alpha = kAlphaMin + m * (alpha_needed - kAlphaMin);
      al = fmaxf(al, target_alpha);
    }
  }
  return al;
}

// SUP: where alpha comes from -- 0: per bin in pre_one (Berouti, or none), 1: one per band
// (sup_band_kernel), 2: one per bin from sup_alpha_kernel.  A template so that the shipped
// strategy keeps its instruction count.
template <int SUP>
__global__ void __launch_bounds__(512)
gain_pre_kernel(const GeoK g, PreArgs a, const float* __restrict__ ref,
                const float* __restrict__ noise_t, const float* __restrict__ mask,
                float* __restrict__ Mg, float* __restrict__ alpha, float* __restrict__ clean,
                const float* __restrict__ sup_alpha) {
  const int i = a.first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  const bool tonal = !(a.tonal_gain >= 0.999f);
  const float* __restrict__ mrow = mask + (size_t)i * a.mask_stride;
  for (int k = 4 * threadIdx.x; k < g.Kp; k += 4 * blockDim.x) {
    const float4 r4 = ld4(ref + ro + k), n4 = ld4(noise_t + ro + k);
    const float rv[4] = {r4.x, r4.y, r4.z, r4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
    float al[4], mg[4], cl[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k + u < g.K) {
        const float m = tonal ? mrow[k + u] : 0.f;
        mg[u] = 0.f; cl[u] = 0.f;
        float alb = 1.f;
        if (SUP == 1) alb = sup_alpha[(size_t)i * kBandPitch + __ldg(&g.band_of_bin[k + u])];
        else if (SUP == 2) alb = sup_alpha[ro + k + u];
        al[u] = pre_one(a, rv[u], nv[u], m, &mg[u], &cl[u], alb);
      } else {
        al[u] = 1.f; mg[u] = 0.f; cl[u] = 0.f;
      }
    }
    st4(alpha + ro + k, make_float4(al[0], al[1], al[2], al[3]));
    if (a.two_d) {
      st4(Mg + ro + k, make_float4(mg[0], mg[1], mg[2], mg[3]));
      st4(clean + ro + k, make_float4(cl[0], cl[1], cl[2], cl[3]));
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// N4: the oversubtraction strategies the shipped reference build never selects.
// ---------------------------------------------------------------------------
struct SupArgs {
  int first_active, two_d, mode;   // mode 1: global SNR, 2: critical-band SNR
  float strength;
};

__device__ __forceinline__ float sup_alpha_from_snr(float spec_sum, float noise_sum, float strength) {
  // suppression_engine.c:147-162 / :177-194
  const float snr_db = 10.0f * log10f(spec_sum / (noise_sum + kSpectralEps));
  if (snr_db <= 0.0f) return strength;
  if (snr_db >= 20.0f) return kAlphaMin;
  const float ns = snr_db / 20.0f;
  return ((1.0f - ns) * strength) + (ns * kAlphaMin);
}

// (frame, band): band sums of the reference spectrum and of the noise -> alpha per band
// (calculate_critical_bands_snr), or their totals -> one alpha for the frame
// (calculate_global_snr), written to all bands.  ref = smoothed SNR (2D) or power (1D).
__global__ void __launch_bounds__(256)
sup_band_kernel(const GeoK g, SupArgs a, const float* __restrict__ ref,
                const float* __restrict__ noise_t, float* __restrict__ band_alpha) {
  __shared__ float PS[kMaxBandTasks], PN[kMaxBandTasks];
  __shared__ float BS[kBandPitch], BN[kBandPitch];
  const int i = a.first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int task = warp; task < g.ntask; task += nw) {
    const int b = g.task_band[task];
    const int k0 = g.task_k[task];
    const int kend = g.task_band[task + 1 < g.ntask ? task + 1 : task] == b && task + 1 < g.ntask
                         ? g.task_k[task + 1] : g.band_start[b + 1];
    float ss = 0.f, sn = 0.f;
    for (int k = k0 + lane; k < kend; k += 32) {
      const float n = noise_t[ro + k];
      float sp = ref[ro + k];
      if (a.two_d) sp *= (n > kNlmNoiseFloorMin ? n : kNlmNoiseFloorMin);   // nlm_filter.c:481-508
      ss += sp;
      sn += n;
    }
    ss = warp_sum(ss);
    sn = warp_sum(sn);
    if (lane == 0) { PS[task] = ss; PN[task] = sn; }
  }
  __syncthreads();
  if (threadIdx.x < kBandPitch) {
    const int b = threadIdx.x;
    float ss = 0.f, sn = 0.f;
    if (b < g.nb)
      for (int t = g.band_task0[b]; t < g.band_task0[b + 1]; t++) { ss += PS[t]; sn += PN[t]; }
    BS[b] = ss;
    BN[b] = sn;
  }
  __syncthreads();
  if (threadIdx.x < kBandPitch) {
    const int b = threadIdx.x;
    float al = kAlphaMin;
    if (a.mode == 1) {
      float ss = 0.f, sn = 0.f;
      for (int q = 0; q < g.nb; q++) { ss += BS[q]; sn += BN[q]; }
      al = sup_alpha_from_snr(ss, sn, a.strength);
    } else if (b < g.nb) {
      al = sup_alpha_from_snr(BS[b], BN[b], a.strength);
    }
    band_alpha[(size_t)i * kBandPitch + b] = al;
  }
}

// clean-signal estimate of calculate_masking_thresholds (suppression_engine.c:202-206): 1D only
// (in 2D gain_pre_kernel has already written max(Mg - n, 0))
__global__ void __launch_bounds__(512)
sup_clean_kernel(const GeoK g, int first_active, const float* __restrict__ P,
                 const float* __restrict__ noise_t, float* __restrict__ clean) {
  const size_t ro = (size_t)(first_active + blockIdx.x) * g.Kp;
  for (int k = 4 * threadIdx.x; k < g.Kp; k += 4 * blockDim.x) {
    const float4 p = ld4(P + ro + k), n = ld4(noise_t + ro + k);
    st4(clean + ro + k, make_float4(fmaxf(p.x - n.x, 0.f), fmaxf(p.y - n.y, 0.f), fmaxf(p.z - n.z, 0.f),
                                    fmaxf(p.w - n.w, 0.f)));
  }
}

// (frame, bin): masking threshold of the bin's band, floored by the absolute threshold of
// hearing, -> noise-to-mask ratio -> alpha (suppression_engine.c:208-231; beta is 0 in both
// processors: undersubtraction = 0)
__global__ void __launch_bounds__(512)
sup_alpha_kernel(const GeoK g, int first_active, float strength, const float* __restrict__ band_T,
                 const float* __restrict__ noise_t, float* __restrict__ alpha_bin) {
  __shared__ float SPLs[kBandPitch], VT[kBandPitch];
  const int i = first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  if (threadIdx.x < g.nb) {
    const int b = threadIdx.x;
    const float T = powf(band_T[(size_t)i * kBandPitch + b], kInvPower);   // band_T holds u = T^0.6
    const float spl = (10.0f * log10f(T + kSpectralEps)) + g.spl_ref;      // absolute_hearing_thresholds.c:141-159
    float vT = powf(10.0f, (spl - g.spl_ref) / 10.0f) - kSpectralEps;
    if (vT < 0.0f) vT = 0.0f;
    SPLs[b] = spl;
    VT[b] = vT;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < g.Kp; k += blockDim.x) {
    float al = kAlphaMin;
    if (k < g.K) {
      const int b = __ldg(&g.band_of_bin[k]);
      const float thr = (SPLs[b] >= __ldg(&g.abs_thr_spl[k])) ? VT[b] : __ldg(&g.abs_thr_lin[k]);
      const float nmr_db = 10.0f * log10f(noise_t[ro + k] / (thr + kSpectralEps));
      if (nmr_db <= 0.0f) al = kAlphaMin + ((strength - kAlphaMin) * 0.1f);
      else if (nmr_db >= 20.0f) al = strength;
      else {
        const float nn = nmr_db / 20.0f;
        al = ((1.0f - nn) * kAlphaMin) + (nn * strength);
      }
    }
    alpha_bin[ro + k] = al;
  }
}

// TRANSIENT_AWARE smoothing of the 1D processor (spectral_smoother.c:141-181): band energies of
// the unsmoothed power -> transient_detector_process per band (transient_detector.c:66-111),
// a recursion over frames with state (smoothed energies + initialised flag) -> onset weights.
__global__ void __launch_bounds__(256)
band_energy_kernel(const GeoK g, int first_active, const float* __restrict__ P, float* __restrict__ band_e) {
  __shared__ float PS[kMaxBandTasks];
  const int i = first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int task = warp; task < g.ntask; task += nw) {
    const int b = g.task_band[task];
    const int k0 = g.task_k[task];
    const int kend = g.task_band[task + 1 < g.ntask ? task + 1 : task] == b && task + 1 < g.ntask
                         ? g.task_k[task + 1] : g.band_start[b + 1];
    float ss = 0.f;
    for (int k = k0 + lane; k < kend; k += 32) ss += P[ro + k];
    ss = warp_sum(ss);
    if (lane == 0) PS[task] = ss;
  }
  __syncthreads();
  if (threadIdx.x < kBandPitch) {
    const int b = threadIdx.x;
    float ss = 0.f;
    if (b < g.nb)
      for (int t = g.band_task0[b]; t < g.band_task0[b + 1]; t++) ss += PS[t];
    band_e[(size_t)i * kBandPitch + b] = ss;
  }
}

constexpr float kTransientAlpha = 0.8f;          // configurations.h:152
constexpr float kOnsetSensitivity = 0.25f;       // :151
constexpr float kMinInnovationEnergy = 1e-10f;   // :150

// one warp, lane = band: weights[f][b] in place of the energies; state[0..31] smoothed energies,
// state[32] = 1 once initialised
__global__ void onset_scan_kernel(int nb, int rows, float* __restrict__ band_e, float* __restrict__ state) {
  const int b = threadIdx.x;
  if (b >= kBandPitch) return;
  float sm = state[b];
  bool init = state[kBandPitch] != 0.f;
  for (int f = 0; f < rows; f++) {
    const float cur = band_e[(size_t)f * kBandPitch + b];
    if (!init) { sm = cur; init = true; }
    const float ratio = cur / (sm + 1e-9f);
    float w = (ratio - 1.0f) / kOnsetSensitivity;
    w = (cur < kMinInnovationEnergy) ? 0.0f : w;
    w = fminf(fmaxf(w, 0.0f), 1.0f);
    const float aa = (w > 0.0f) ? (kTransientAlpha * 0.5f) : kTransientAlpha;
    sm = (cur * (1.0f - aa)) + (sm * aa);
    band_e[(size_t)f * kBandPitch + b] = (b < nb) ? w : 0.f;
  }
  state[b] = sm;
  if (b == 0 && rows > 0) state[kBandPitch] = 1.f;
}

// ---------------------------------------------------------------------------
// (frame, band): band energies, spreading, tonality -> a_t[b]
// ---------------------------------------------------------------------------
__device__ __forceinline__ float spreading_exponent(float dz, float level_db) {
  // masking_estimator.c:305-318: the spreading gain is 10^(this)
  const float s_up = fminf(fmaxf(15.0f - ((level_db - 40.0f) * 0.2f), 5.0f), 15.0f);
  const float s_total = (25.0f + s_up) / 2.0f;
  const float s_offset = (25.0f - s_up) / 2.0f;
  const float y = dz + 0.474f;
  const float sf_db = 15.81f + (s_offset * y) - (s_total * sqrtf(1.0f + (y * y)));
  return sf_db / 10.0f;
}

// (E * 10^t)^0.4 as 2^(0.4 (log2 E + t log2 10)): the exponent in double (exact to float
// precision, log2 E computed once per band), one exp2f on the fraction.  Within 2 ulp of the exact
// value, where powf(E * exp10f(t), 0.4f) stacks three roundings; 625 x 2 of these per frame.
__device__ __forceinline__ float spread_term(double lg2_energy, float t) {
  const double a = (double)kAdditivityExp * fma((double)t, 3.321928094887362, lg2_energy);
  if (!(a > -125.0)) return 0.0f;          // E == 0 (log2 = -inf) and products below float range
  const double r = rint(a);
  return exp2f((float)(a - r)) * __int_as_float(((int)r + 127) << 23);
}

__global__ void __launch_bounds__(256)
band_kernel(const GeoK g, int first_active, int two_d, const float* __restrict__ stable,
            const float* __restrict__ Xre_t, const float* __restrict__ noise_t,
            float* __restrict__ band_a) {
  __shared__ float E[2][kBandPitch], SL[2][kBandPitch], SV[2][kBandPitch];
  __shared__ float LV[2][kBandPitch], SPR[2][kBandPitch];
  __shared__ float PART[6][kMaxBandTasks];
  __shared__ float TERM[2][kMaxBands][kMaxBands + 1];
  __shared__ double LG2E[2][kBandPitch];
  const int i = first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int task = warp; task < g.ntask; task += nw) {
    const int b = g.task_band[task];
    const int k0 = g.task_k[task];
    const int kend = g.task_band[task + 1 < g.ntask ? task + 1 : task] == b && task + 1 < g.ntask
                         ? g.task_k[task + 1] : g.band_start[b + 1];
    float e0 = 0.f, l0 = 0.f, v0 = 0.f, e1 = 0.f, l1 = 0.f, v1 = 0.f;
    // a task is at most 128 bins in all but extreme geometries: its (up to) 4 loads per lane
    // and row are requested together, then summed in the same k order as a plain loop
    for (int kb = k0 + lane; kb < kend; kb += 128) {
      float sl[4], xr[4], nz[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = kb + 32 * u;
        const bool in = k < kend;
        sl[u] = in ? stable[ro + k] : 0.f;
        xr[u] = (in && two_d) ? Xre_t[ro + k] : 0.f;
        nz[u] = (in && two_d) ? noise_t[ro + k] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (kb + 32 * u < kend) {
          const float s = sl[u];
          const float sv = fmaxf(s, kSpectralEps);
          e0 += s; v0 += sv; l0 += log10f(sv);
          if (two_d) {
            const float fu = fmaxf(xr[u] - nz[u], 0.0f);   // masking_veto.c:151-157
            const float fv = fmaxf(fu, kSpectralEps);
            e1 += fu; v1 += fv; l1 += log10f(fv);
          }
        }
      }
    }
    // the six sums in one transposing butterfly (9 shuffles instead of 30): at xor 16 / 8 / 4 a
    // lane keeps the half of its values that its own bit selects and hands over the other
    // half, then xor 2 / 1 finish the one value left.  Same pairing as warp_sum, same bits.
    float v8[8] = {e0, l0, v0, e1, l1, v1, 0.f, 0.f};
    {
      const bool h = lane & 16;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float keep = h ? v8[q + 4] : v8[q], give = h ? v8[q] : v8[q + 4];
        v8[q] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      }
    }
    {
      const bool h = lane & 8;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float keep = h ? v8[q + 2] : v8[q], give = h ? v8[q] : v8[q + 2];
        v8[q] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      }
    }
    {
      const bool h = lane & 4;
      const float keep = h ? v8[1] : v8[0], give = h ? v8[0] : v8[1];
      v8[0] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 2);
    v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 1);
    // lane 4 * idx holds sum idx, idx = 4 * bit4 + 2 * bit3 + bit2
    if ((lane & 3) == 0) {
      const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      if (idx < 6) PART[idx][task] = v8[0];
    }
  }
  __syncthreads();
  if (threadIdx.x < g.nb) {
    const int b = threadIdx.x;
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = g.band_task0[b]; t < g.band_task0[b + 1]; t++)
#pragma unroll
      for (int q = 0; q < 6; q++) a[q] += PART[q][t];
    E[0][b] = a[0]; SL[0][b] = a[1]; SV[0][b] = a[2];
    E[1][b] = a[3]; SL[1][b] = a[4]; SV[1][b] = a[5];
  }
  __syncthreads();
  const int nspec = two_d ? 2 : 1;
  if (threadIdx.x < nspec * kBandPitch) {
    const int s = threadIdx.x / kBandPitch, b = threadIdx.x % kBandPitch;
    if (b < g.nb) {
      LV[s][b] = (10.0f * log10f(E[s][b] + kSpectralEps)) + kDbFsToSpl;
      LG2E[s][b] = log2((double)E[s][b]);
    }
  }
  __syncthreads();
  // spreading terms (E_j g(i-j, L_j))^0.4 for every (spectrum, band i, band j): all threads
  const int nb2 = g.nb * g.nb;
  const unsigned nbmagic = (65536u + (unsigned)g.nb - 1u) / (unsigned)g.nb;   // r / nb for r < 1024, nb <= 32
  for (int item = threadIdx.x; item < nspec * nb2; item += blockDim.x) {
    const int s = item >= nb2 ? 1 : 0;
    const int r = item - s * nb2;
    const int b = (int)(((unsigned)r * nbmagic) >> 16), j = r - b * g.nb;
    const float dz = (float)b - (float)j;
    TERM[s][b][j] = spread_term(LG2E[s][j], spreading_exponent(dz, LV[s][j]));
  }
  __syncthreads();
  if (threadIdx.x < nspec * kBandPitch) {
    const int s = threadIdx.x / kBandPitch, b = threadIdx.x % kBandPitch;
    if (b < g.nb) {
      float acc = 0.f;
      for (int j = 0; j < g.nb; j++) acc += TERM[s][b][j];   // reference order (j ascending)
      const float spr = powf(acc, kInvAdd);
      // tonality (masking_estimator.c:320-349)
      const float bins = (float)g.band_start[b + 1] - (float)g.band_start[b];
      float tf = 1.0f;
      if (bins > 1.0f) {
        const float sfm = 10.0f * ((SL[s][b] / bins) - log10f(SV[s][b] / bins));
        tf = fminf(fmaxf((sfm - 0.0f) / (-60.0f - 0.0f), 0.0f), 1.0f);
      }
      const float bark_idx = fminf((float)(b + 1), 25.0f);
      const float offset = (tf * (14.5f + bark_idx)) + (6.0f * (1.0f - tf));
      float thr = powf(10.0f, (log10f(spr + kSpectralEps) - (offset / 10.0f)));
      if (s == 1) thr *= g.bwd;      // backward masking (masking_estimator.c:279-283)
      SPR[s][b] = powf(thr, kPowerLawExp);
    }
  }
  __syncthreads();
  if (threadIdx.x < kBandPitch) {
    const int b = threadIdx.x;
    float a = 0.f;
    if (b < g.nb) a = SPR[0][b] + (two_d ? SPR[1][b] : 0.f);
    band_a[(size_t)i * kBandPitch + b] = a;
  }
}

// (frame, band): absolute-threshold floor, band NMR -> protection
__global__ void __launch_bounds__(256)
nmr_kernel(const GeoK g, int first_active, const float* __restrict__ band_T,
           const float* __restrict__ noise_t, float* __restrict__ band_p) {
  const int i = first_active + blockIdx.x;
  const size_t ro = (size_t)i * g.Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __shared__ float SPLs[kBandPitch], VT[kBandPitch];
  __shared__ float PN[2][kMaxBandTasks];
  if (threadIdx.x < g.nb) {
    const int b = threadIdx.x;
    const float T = powf(band_T[(size_t)i * kBandPitch + b], kInvPower);   // band_T holds u = T^0.6
    // absolute_hearing_thresholds.c:141-159
    const float spl = (10.0f * log10f(T + kSpectralEps)) + g.spl_ref;
    float vT = powf(10.0f, (spl - g.spl_ref) / 10.0f) - kSpectralEps;
    if (vT < 0.0f) vT = 0.0f;
    SPLs[b] = spl;
    VT[b] = vT;
  }
  __syncthreads();
  for (int task = warp; task < g.ntask; task += nw) {
    const int b = g.task_band[task];
    const int k0 = g.task_k[task];
    const int kend = g.task_band[task + 1 < g.ntask ? task + 1 : task] == b && task + 1 < g.ntask
                         ? g.task_k[task + 1] : g.band_start[b + 1];
    const float spl = SPLs[b], vT = VT[b];
    float sn = 0.f, sth = 0.f;
    for (int kb = k0 + lane; kb < kend; kb += 128) {   // loads batched as in band_kernel
      float nz[4], ts[4], tl[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = kb + 32 * u;
        const bool in = k < kend;
        nz[u] = in ? noise_t[ro + k] : 0.f;
        ts[u] = in ? __ldg(&g.abs_thr_spl[k]) : 0.f;
        tl[u] = in ? __ldg(&g.abs_thr_lin[k]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (kb + 32 * u < kend) {
          sn += nz[u];
          sth += (spl >= ts[u]) ? vT : tl[u];
        }
      }
    }
    sn = warp_sum(sn);
    sth = warp_sum(sth);
    if (lane == 0) {
      PN[0][task] = sn;
      PN[1][task] = sth;
    }
  }
  __syncthreads();
  if (threadIdx.x < g.nb) {
    const int b = threadIdx.x;
    float sn = 0.f, sth = 0.f;
    for (int t = g.band_task0[b]; t < g.band_task0[b + 1]; t++) {
      sn += PN[0][t];
      sth += PN[1][t];
    }
    float p = 0.f;
    if (g.band_start[b + 1] > g.band_start[b]) {   // masking_veto.c:178-213
      const float nmr = 10.0f * log10f((sn + kSpectralEps) / (sth + kSpectralEps));
      if (nmr <= 0.0f) p = 1.0f;
      else if (nmr >= kVetoNmrRange) p = 0.0f;
      else p = 1.0f - (nmr / kVetoNmrRange);
    }
    band_p[(size_t)i * kBandPitch + b] = p;
  }
}

// ---------------------------------------------------------------------------
// Sequential-in-time scans, staged through shared memory.
//
// The recursions above are strictly sequential over frames (and must stay so:
// streaming invariance is bit-exact), independent over columns (bins / bands).
// Walking global memory directly leaves one dependent ~1 us round trip per 16
// frames on the critical path (305-360 us per 4096 frames, a third of a
// sub-call's kernel time).  Here a CTA owns 32 columns: 7 mover warps stream
// 32-frame tiles global -> shared with cp.async, four tiles ahead of the
// scanner warp (lane = column), which runs the recurrence on shared memory in
// place; the movers then write the finished tile back.  One barrier per tile.
//   MODE 0: m = 0.5 v + 0.5 m   (bin EMA masking_veto.c:137-147, band EMA
//           spectral_smoother.c:213-223), in place
//   MODE 1: u = c[col] u + a    (forward masking, linear in u = T^0.6:
//           T = (T_f^p + (T_prev*fwd)^p + (T_fut*bwd)^p)^(1/p), masking_estimator.c:270-288;
//           T = u^(1/0.6) is taken in nmr_kernel), A -> B
// ---------------------------------------------------------------------------
constexpr int kRsTile = 64;      // frames per tile
constexpr int kRsRing = 5;       // tiles resident in shared memory
constexpr int kRsAhead = 3;      // prefetch distance (tiles)

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct RsCoef { float c[32]; };   // MODE 1: per-column coefficient (one CTA, <= 32 bands)

template <int MODE>
__global__ void __launch_bounds__(256)
rowscan_kernel(int ncols, int pitch, int rows, const float* A, float* B,   // B may alias A
               const RsCoef coef, float* __restrict__ state) {
  __shared__ __align__(16) float ring[kRsRing][kRsTile][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < ncols;
  const int ntiles = (rows + kRsTile - 1) / kRsTile;
  constexpr int kMovers = 7;

  // movers (warps 1..7): 16-byte vectors, 8 lanes per 32-column row, 4 rows per warp
  // instruction.  Rows are Kp-pitched (Kp % 4 == 0), so whole vectors are always in bounds.
  const int mrow = lane >> 3, mvec = (lane & 7) * 4;
  const bool mv = blockIdx.x * 32 + mvec < pitch && blockIdx.x * 32 + mvec < ((ncols + 3) & ~3);
  const float* Ab = A + blockIdx.x * 32 + mvec;
  float* Bb = B + blockIdx.x * 32 + mvec;
  auto load_tile = [&](int t) {
    if (warp > 0 && t < ntiles) {
      for (int r = 4 * (warp - 1) + mrow; r < kRsTile; r += 4 * kMovers) {
        const int f = t * kRsTile + r;
        if (f < rows && mv) cp_async16(&ring[t % kRsRing][r][mvec], Ab + (size_t)f * pitch);
      }
    }
    cp_async_commit();
  };
  auto store_tile = [&](int tb) {
    for (int r = 4 * (warp - 1) + mrow; r < kRsTile; r += 4 * kMovers) {
      const int f = tb * kRsTile + r;
      if (f < rows && mv)
        *reinterpret_cast<float4*>(Bb + (size_t)f * pitch) =
            *reinterpret_cast<const float4*>(&ring[tb % kRsRing][r][mvec]);
    }
  };
  float m = 0.f, c = 0.f;
  if (warp == 0 && cv) {
    m = state[col];
    if (MODE == 1) c = coef.c[lane];
  }
  for (int t = 0; t < kRsAhead; t++) load_tile(t);
  for (int t = 0; t < ntiles; t++) {
    // tile t has landed once at most kRsAhead-1 younger groups are pending
    cp_async_wait<kRsAhead - 1>();
    __syncthreads();
    if (warp == 0) {
      const int nr = (rows - t * kRsTile < kRsTile) ? rows - t * kRsTile : kRsTile;
      float* tp = &ring[t % kRsRing][0][lane];
      // One dependent FFMA per frame is the floor of this loop.  MODE 0: both products of
      // 0.5 v + 0.5 m are exact, so fma(0.5, m, 0.5 v) rounds to the same value and keeps the
      // multiply of v off the chain.  Full tiles are unrolled in groups of 16 frames whose
      // loads are issued before the chain needs them.  (Measured: the scanner warp and the
      // movers each need ~10 cycles per frame, the barrier ~300 per tile; a deeper ring
      // (12 tiles, 10 ahead), 4 / 8 / 16 mover warps or unrolled movers changed nothing.)
      if (nr == kRsTile) {
#pragma unroll
        for (int r0 = 0; r0 < kRsTile; r0 += 16) {
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; u++) {
            v[u] = tp[(r0 + u) * 32];
            if (MODE == 0) v[u] *= kVetoSmoothing;
          }
#pragma unroll
          for (int u = 0; u < 16; u++) {
            if (MODE == 0) m = fmaf(1.0f - kVetoSmoothing, m, v[u]);
            else m = fmaf(c, m, v[u]);
            tp[(r0 + u) * 32] = m;
          }
        }
      } else {
        for (int r = 0; r < nr; r++) {
          const float v = tp[r * 32];
          if (MODE == 0) m = fmaf(1.0f - kVetoSmoothing, m, kVetoSmoothing * v);
          else m = fmaf(c, m, v);
          tp[r * 32] = m;
        }
      }
    } else if (t > 0) {
      store_tile(t - 1);                     // finished in the previous iteration
    }
    load_tile(t + kRsAhead);
  }
  __syncthreads();
  if (warp > 0 && ntiles > 0) store_tile(ntiles - 1);
  if (warp == 0 && cv) state[col] = m;
}

// 1D variant of the staged scan: two coupled recurrences per bin over (P, noise) rows --
// release-only smoothing of the power spectrum (spectral_smoother.c:183-192, state update
// :132-133) and the stabilised clean-signal EMA of the veto (masking_veto.c:137-147).
// P tiles become Mg tiles and noise tiles become `stable` tiles in place in shared memory.
constexpr int kRs1Tile = 32, kRs1Ring = 5, kRs1Ahead = 3;

__global__ void __launch_bounds__(256)
rowscan1d_kernel(int ncols, int pitch, int rows, float smoothing, const float* P,
                 const float* noise, float* Mg, float* stable, float* st_state, float* sp_state,
                 const float* __restrict__ onset, const int* __restrict__ band_of_bin) {
  __shared__ __align__(16) float ringP[kRs1Ring][kRs1Tile][32];
  __shared__ __align__(16) float ringN[kRs1Ring][kRs1Tile][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < ncols;
  const int ntiles = (rows + kRs1Tile - 1) / kRs1Tile;
  constexpr int kMovers = 7;
  const int mrow = lane >> 3, mvec = (lane & 7) * 4;
  const bool mv = blockIdx.x * 32 + mvec < pitch && blockIdx.x * 32 + mvec < ((ncols + 3) & ~3);
  const size_t cb = (size_t)blockIdx.x * 32 + mvec;
  auto load_tile = [&](int t) {
    if (warp > 0 && t < ntiles) {
      for (int r = 4 * (warp - 1) + mrow; r < kRs1Tile; r += 4 * kMovers) {
        const int f = t * kRs1Tile + r;
        if (f < rows && mv) {
          cp_async16(&ringP[t % kRs1Ring][r][mvec], P + (size_t)f * pitch + cb);
          cp_async16(&ringN[t % kRs1Ring][r][mvec], noise + (size_t)f * pitch + cb);
        }
      }
    }
    cp_async_commit();
  };
  auto store_tile = [&](int tb) {
    for (int r = 4 * (warp - 1) + mrow; r < kRs1Tile; r += 4 * kMovers) {
      const int f = tb * kRs1Tile + r;
      if (f < rows && mv) {
        *reinterpret_cast<float4*>(Mg + (size_t)f * pitch + cb) =
            *reinterpret_cast<const float4*>(&ringP[tb % kRs1Ring][r][mvec]);
        *reinterpret_cast<float4*>(stable + (size_t)f * pitch + cb) =
            *reinterpret_cast<const float4*>(&ringN[tb % kRs1Ring][r][mvec]);
      }
    }
  };
  float st = 0.f, sp = 0.f;
  // TRANSIENT_AWARE (spectral_smoother.c:141-181): the smoothing of a bin is scaled by
  // 1 - onset weight of its band in that frame; onset == nullptr is the FIXED smoother
  const float* ow = nullptr;
  if (warp == 0 && cv) {
    st = st_state[col];
    sp = sp_state[col];
    if (onset) ow = onset + __ldg(&band_of_bin[col]);
  }
  for (int t = 0; t < kRs1Ahead; t++) load_tile(t);
  for (int t = 0; t < ntiles; t++) {
    cp_async_wait<kRs1Ahead - 1>();
    __syncthreads();
    if (warp == 0) {
      const int nr = (rows - t * kRs1Tile < kRs1Tile) ? rows - t * kRs1Tile : kRs1Tile;
      float* pp = &ringP[t % kRs1Ring][0][lane];
      float* pn = &ringN[t % kRs1Ring][0][lane];
#pragma unroll 8
      for (int r = 0; r < nr; r++) {
        const float c = pp[r * 32], n = pn[r * 32];
        float sm = smoothing;
        if (ow) sm = smoothing * (1.0f - ow[(size_t)(t * kRs1Tile + r) * kBandPitch]);
        float v = c;
        if (c > sp) v = (sm * sp) + ((1.0f - sm) * c);
        sp = v;
        const float cl = fmaxf(v - n, 0.0f);
        st = (kVetoSmoothing * cl) + ((1.0f - kVetoSmoothing) * st);
        pp[r * 32] = v;
        pn[r * 32] = st;
      }
    } else if (t > 0) {
      store_tile(t - 1);
    }
    load_tile(t + kRs1Ahead);
  }
  __syncthreads();
  if (warp > 0 && ntiles > 0) store_tile(ntiles - 1);
  if (warp == 0 && cv) {
    st_state[col] = st;
    sp_state[col] = sp;
  }
}

// ---------------------------------------------------------------------------
// (frame, bin): veto interpolation, Wiener gain, whitened floor, apply
// ---------------------------------------------------------------------------
struct FinalArgs {
  int first_active, frames, mask_stride, wt_stride, use_veto, residual;
  float depth, reduction, tonal, whitening;
  int gain_type;        // GainCalculationType of the reference (gain_calculator.h:27-31), 0 = Wiener
};

__device__ __forceinline__ float final_gain(const GeoK& g, const FinalArgs& a, const float* prot, int k,
                                            float n, float spec, float al, float stab, float mroot,
                                            float wtv) {
  if (a.use_veto) {
    // masking_veto.c:230-266
    // bp = lerp(prot[cb], prot[cb + 1], t); t (sb_geometry.cu) is 0 where the reference takes
    // prot[cb] alone, and prot[] is 0 past the last band, so the one expression covers all cases
    const int cb = __ldg(&g.cband[k]);
    const float tc = __ldg(&g.veto_t[k]);
    const float y0 = prot[cb], y1 = prot[cb + 1];
    const float bp = (y0 * (1.0f - tc)) + (y1 * tc);
    const float bin_snr = stab / (n + kSpectralEps);
    const float sc = fminf(bin_snr / 1.0f, 1.0f);
    const float veto = bp * sc * a.depth;
    al = al + ((1.0f - al) * veto);
  }
  float gain;
  if (a.gain_type == 2) {
    // generalized_spectral_subtraction (gain_calculator.c:113-139) with beta = 0: both
    // processors pass undersubtraction 0 (spectral_denoiser.c:296, spectral_2d_denoiser.c:399)
    if (spec > FLT_MIN) {
      const float ratio = n / spec;
      const float ratio_sq = ratio * ratio;
      if (ratio_sq < (1.0f / (al + 0.0f))) gain = fmaxf(sqrtf(fmaxf(1.0f - (al * ratio_sq), 0.0f)), 0.0f);
      else gain = 0.0f;
    } else {
      gain = 1.0f;
    }
  } else {
    const float sn = n * al;
    if (a.gain_type == 1) {
      // spectral_gating (gain_calculator.c:71-111)
      if (sn > FLT_MIN) gain = (sn > spec) ? 0.0f : 1.0f;
      else gain = 1.0f;
    } else {
      // wiener_subtraction (gain_calculator.c:27-69)
      if (sn > FLT_MIN) gain = (spec > sn) ? (spec - sn) / spec : 0.0f;
      else gain = 1.0f;
    }
  }
  // noise_floor_manager.c:94-157
  if (a.reduction >= 0.999f && a.tonal >= 0.999f) {
    gain = 1.0f;
  } else {
    const float m = mroot;           // mask^(1/4), taken by the caller
    const float dual = (a.reduction * (1.0f - m)) + (a.tonal * m);
    const float target = dual + (a.whitening * (a.reduction - dual));
    const float rdepth = 1.0f - target;
    const float eff = 1.0f + (rdepth * (wtv - 1.0f));
    float fl = target * eff;
    if (fl > 1.0f) fl = 1.0f;
    gain = fmaxf(fl, gain);
  }
  return gain;
}

__global__ void __launch_bounds__(512)
final_kernel(const GeoK g, FinalArgs a, const float* __restrict__ Xre_t,
             const float* __restrict__ Xim_t, const float* __restrict__ noise_t,
             const float* __restrict__ Mg, const float* __restrict__ alpha,
             const float* __restrict__ stable, const float* __restrict__ band_p,
             const float* __restrict__ wt, const float* __restrict__ mask,
             float* __restrict__ Yre, float* __restrict__ Yim) {
  __shared__ float prot[kBandPitch];
  const int i = blockIdx.x;
  const bool active = i >= a.first_active;
  if (active && a.use_veto) {
    if (threadIdx.x < kBandPitch) {
      const int b = threadIdx.x;
      float v = 0.f;
      if (b < g.nb) {
        // spatial 3-tap from band 1 upwards (spectral_smoother.c:194-211)
        const float* p = band_p + (size_t)i * kBandPitch;
        const float cur = p[b];
        if (b == 0 || g.nb < 2) v = cur;
        else {
          const float next = (b < g.nb - 1) ? p[b + 1] : cur;
          v = (0.25f * p[b - 1]) + (0.5f * cur) + (0.25f * next);
        }
      }
      prot[b] = v;
    }
    __syncthreads();
  }
  const size_t ro = (size_t)i * g.Kp;
  const bool floor_on = !(a.reduction >= 0.999f && a.tonal >= 0.999f);
  const float* __restrict__ mrow = mask + (size_t)i * a.mask_stride;
  const float* __restrict__ wrow = wt + (size_t)i * a.wt_stride;
  for (int k = 4 * threadIdx.x; k < g.Kp; k += 4 * blockDim.x) {
    const float4 xr4 = ld4(Xre_t + ro + k), xi4 = ld4(Xim_t + ro + k);
    if (!active) {
      st4(Yre + ro + k, xr4);
      st4(Yim + ro + k, xi4);
      continue;
    }
    const float4 n4 = ld4(noise_t + ro + k), sp4 = ld4(Mg + ro + k), al4 = ld4(alpha + ro + k);
    float4 sb4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.use_veto) sb4 = ld4(stable + ro + k);
    const float xr[4] = {xr4.x, xr4.y, xr4.z, xr4.w}, xi[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
    const float nv[4] = {n4.x, n4.y, n4.z, n4.w}, sp[4] = {sp4.x, sp4.y, sp4.z, sp4.w};
    const float al[4] = {al4.x, al4.y, al4.z, al4.w}, sb[4] = {sb4.x, sb4.y, sb4.z, sb4.w};
    float yr[4], yi[4], mq[4];
#pragma unroll
    for (int u = 0; u < 4; u++) mq[u] = (floor_on && k + u < g.K) ? mrow[k + u] : 0.f;
    // noise_floor_manager.c:129-131: mask^(1/4); the tonal mask is zero almost everywhere, the
    // two square roots are skipped for a whole group of bins without a tonal component
    if (mq[0] > 0.0f || mq[1] > 0.0f || mq[2] > 0.0f || mq[3] > 0.0f) {
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (mq[u] > 0.0f) mq[u] = sqrtf(sqrtf(mq[u]));
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (k + u < g.K) {
        const float wtv = floor_on ? wrow[k + u] : 1.f;
        const float gain = final_gain(g, a, prot, k + u, nv[u], sp[u], al[u], sb[u], mq[u], wtv);
        // denoiser_post_process.c:36-50
        if (a.residual) {
          yr[u] = xr[u] - (xr[u] * gain);
          yi[u] = xi[u] - (xi[u] * gain);
        } else {
          yr[u] = xr[u] * gain;
          yi[u] = xi[u] * gain;
        }
      } else {
        yr[u] = xr[u];
        yi[u] = xi[u];
      }
    }
    st4(Yre + ro + k, make_float4(yr[0], yr[1], yr[2], yr[3]));
    st4(Yim + ro + k, make_float4(yi[0], yi[1], yi[2], yi[3]));
  }
}

}  // namespace

// CTA size of the one-row-per-CTA elementwise kernels: the row's quads over the fewest rounds
// of at most 512 threads, rounded up to a warp
static int row_threads(const GeoK& g) {
  const int quads = g.Kp / 4;
  const int rounds = (quads + 511) / 512;
  const int t = (quads + rounds - 1) / rounds;
  return ((t + 31) / 32) * 32;
}

bool launch_morph_profile(const GeoK& g, const float* prof, float aggressiveness, float* out,
                          cudaStream_t st) {
  KTimer kt_(kcMisc, st);
  morph_kernel<<<(g.Kp + 255) / 256, 256, 0, st>>>(g, prof, aggressiveness, out);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_fill_rows(const GeoK& g, const float* row, float* dst, int rows, cudaStream_t st) {
  KTimer kt_(kcMisc, st);
  if (rows <= 0) return true;
  dim3 grid((g.Kp + 255) / 256, rows < 1024 ? rows : 1024);
  fill_rows_kernel<<<grid, 256, 0, st>>>(g, row, dst, rows);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_snr(const GeoK& g, const float* P, const float* noise, float* S, int frames,
                cudaStream_t st) {
  KTimer kt_(kcSnr, st);
  if (frames <= 0) return true;
  snr_kernel<<<frames, row_threads(g), 0, st>>>(g, P, noise, S);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_tonal_mask(const GeoK& g, const float* det, int det_stride, int rows, float* mask,
                       cudaStream_t st) {
  KTimer kt_(kcTonal, st);
  if (rows <= 0) return true;
  if (rows < 32) {
    dim3 grid((g.Kp + 255) / 256, rows);
    tonal_mask_direct_kernel<<<grid, 256, 0, st>>>(g, det, det_stride, mask, g.Kp);
    SB_LAUNCH_CHECK();
    return true;
  }
  const int segs = (g.Kp + kTmSeg - 1) / kTmSeg;
  const int blocks = (rows * segs + kTmThreads - 1) / kTmThreads;
  tonal_mask_kernel<<<blocks, kTmThreads, 0, st>>>(g, det, det_stride, mask, g.Kp, rows, segs);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_whitening(const GeoK& g, const float* noise, int noise_stride, int rows,
                      float whitening, float* wt, cudaStream_t st) {
  KTimer kt_(kcWhiten, st);
  if (rows <= 0) return true;
  const size_t smem = sizeof(float) * (size_t)g.K;
  if (smem > 48 * 1024)
    SB_CUDA(cudaFuncSetAttribute((const void*)whitening_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  whitening_kernel<<<rows, 512, smem, st>>>(g, noise, noise_stride, whitening, wt, g.Kp);
  SB_LAUNCH_CHECK();
  return true;
}

// Runs the whole chain for the frames of one sub-call.  Rows of the "delayed"
// buffers (Xre/Xim/noise) are indexed by the frame index i of this call: in 2D
// they carry 4 history rows in front, so row i IS frame i-4; in 1D row i is
// frame i itself.
bool launch_gain_chain(const GeoK& g, const GainArgs& a, Lane& L, Handle& h, cudaStream_t st) {
  const int M = a.frames;
  if (M <= 0) return true;
  const int fa = a.first_active < M ? a.first_active : M;
  const int active = M - fa;
  const bool use_veto = !(a.depth < 0.0f);   // masking_veto.c:131-134
  if (active > 0) {
    PreArgs pa;
    pa.first_active = fa;
    pa.two_d = a.two_d ? 1 : 0;
    pa.mask_stride = a.mask_stride;
    pa.alpha_max_user = kAlphaMin + (a.strength * (kAlphaMax - kAlphaMin));
    pa.tonal_gain = a.tonal;
    pa.suppression = a.suppression_type;
    const float* refspec = a.two_d ? L.Shat : L.P;
    const float* sup_alpha = nullptr;
    if (a.suppression_type == 1 || a.suppression_type == 2) {
      // N4: global / critical-band SNR -> one alpha per band row (L.band_a is free until the veto)
      KTimer kt_(kcGainPre, st);
      SupArgs sa;
      sa.first_active = fa;
      sa.two_d = a.two_d ? 1 : 0;
      sa.mode = a.suppression_type;
      sa.strength = a.strength;
      sup_band_kernel<<<active, 256, 0, st>>>(g, sa, refspec, L.noise, L.band_a);
      SB_LAUNCH_CHECK();
      sup_alpha = L.band_a;
    } else if (a.suppression_type == 3) {
      // N4: masking thresholds of the clean-signal estimate max(spectrum - noise, 0) through the
      // suppression engine's own masking estimator (same Bark tables; its own forward-masking
      // state, h.d_band_state row 2) -> per-bin alpha in L.wt-free scratch: L.Yre (written last)
      KTimer kt_(kcGainPre, st);
      float* cleanp = L.stable;                       // 2D: filled by the plain pre pass below
      if (a.two_d) {
        PreArgs p0 = pa;
        p0.suppression = 4;                           // alpha is recomputed afterwards
        gain_pre_kernel<0><<<active, row_threads(g), 0, st>>>(g, p0, refspec, L.noise, a.mask, L.Mg, L.alpha,
                                                              L.stable, nullptr);
        SB_LAUNCH_CHECK();
      } else {
        cleanp = L.Mg;                                // 1D: Mg is only written by the smoother later
        sup_clean_kernel<<<active, row_threads(g), 0, st>>>(g, fa, L.P, L.noise, L.Mg);
        SB_LAUNCH_CHECK();
      }
      band_kernel<<<active, 256, 0, st>>>(g, fa, 0, cleanp, L.Xre, L.noise, L.band_a);
      SB_LAUNCH_CHECK();
      RsCoef fc;
      for (int b = 0; b < 32; b++) fc.c[b] = b < g.nb ? g.fwd_pow[b] : 0.f;
      rowscan_kernel<1><<<1, 256, 0, st>>>(g.nb, kBandPitch, active, L.band_a + (size_t)fa * kBandPitch,
                                           L.band_T + (size_t)fa * kBandPitch, fc,
                                           h.d_band_state + 2 * kBandPitch);
      SB_LAUNCH_CHECK();
      sup_alpha_kernel<<<active, 512, 0, st>>>(g, fa, a.strength, L.band_T, L.noise, L.Yre);
      SB_LAUNCH_CHECK();
      sup_alpha = L.Yre;
    }
    {
      KTimer kt_(kcGainPre, st);
    const int sup = a.suppression_type;
    if (sup == 1 || sup == 2)
      gain_pre_kernel<1><<<active, row_threads(g), 0, st>>>(g, pa, refspec, L.noise, a.mask, L.Mg, L.alpha,
                                                            L.stable, sup_alpha);
    else if (sup == 3)
      gain_pre_kernel<2><<<active, row_threads(g), 0, st>>>(g, pa, refspec, L.noise, a.mask, L.Mg, L.alpha,
                                                            L.stable, sup_alpha);
    else
      gain_pre_kernel<0><<<active, row_threads(g), 0, st>>>(g, pa, refspec, L.noise, a.mask, L.Mg, L.alpha,
                                                            L.stable, sup_alpha);
    SB_LAUNCH_CHECK();
    }
    if (use_veto || !a.two_d) {
      {
        KTimer kt_(kcBinScan, st);
      if (a.two_d) {
        float* sp = L.stable + (size_t)fa * g.Kp;
        rowscan_kernel<0><<<(g.K + 31) / 32, 256, 0, st>>>(g.K, g.Kp, active, sp, sp, RsCoef(),
                                                           h.d_stable);
      } else {
        const size_t o = (size_t)fa * g.Kp;
        const float* onset = nullptr;
        if (a.smoothing_type == 2) {
          // N4: TRANSIENT_AWARE -- band energies of the unsmoothed power, onset weights per band
          band_energy_kernel<<<active, 256, 0, st>>>(g, fa, L.P, L.band_T);
          SB_LAUNCH_CHECK();
          onset_scan_kernel<<<1, kBandPitch, 0, st>>>(g.nb, active, L.band_T + (size_t)fa * kBandPitch,
                                                      h.d_band_state + 3 * kBandPitch);
          SB_LAUNCH_CHECK();
          onset = L.band_T + (size_t)fa * kBandPitch;
        }
        rowscan1d_kernel<<<(g.K + 31) / 32, 256, 0, st>>>(g.K, g.Kp, active, a.smoothing, L.P + o,
                                                          L.noise + o, L.Mg + o, L.stable + o,
                                                          h.d_stable, h.d_sp_prev, onset, g.band_of_bin);
      }
      SB_LAUNCH_CHECK();
      }
    }
    if (use_veto) {
      {
        KTimer kt_(kcBand, st);
      band_kernel<<<active, 256, 0, st>>>(g, fa, a.two_d ? 1 : 0, L.stable, L.Xre, L.noise,
                                          L.band_a);
      SB_LAUNCH_CHECK();
      }
      {
        KTimer kt_(kcTScan, st);
      RsCoef fc;
      for (int b = 0; b < 32; b++) fc.c[b] = b < g.nb ? g.fwd_pow[b] : 0.f;
      rowscan_kernel<1><<<1, 256, 0, st>>>(g.nb, kBandPitch, active,
                                           L.band_a + (size_t)fa * kBandPitch,
                                           L.band_T + (size_t)fa * kBandPitch, fc, h.d_band_state);
      SB_LAUNCH_CHECK();
      }
      {
        KTimer kt_(kcNmr, st);
      nmr_kernel<<<active, 256, 0, st>>>(g, fa, L.band_T, L.noise, L.band_p);
      SB_LAUNCH_CHECK();
      }
      {
        KTimer kt_(kcPScan, st);
      rowscan_kernel<0><<<1, 256, 0, st>>>(g.nb, kBandPitch, active,
                                           L.band_p + (size_t)fa * kBandPitch,
                                           L.band_p + (size_t)fa * kBandPitch, RsCoef(),
                                           h.d_band_state + kBandPitch);
      SB_LAUNCH_CHECK();
      }
    }
  }
  FinalArgs fa2;
  fa2.first_active = fa;
  fa2.frames = M;
  fa2.mask_stride = a.mask_stride;
  fa2.wt_stride = a.wt_stride;
  fa2.use_veto = use_veto ? 1 : 0;
  fa2.residual = a.residual ? 1 : 0;
  fa2.depth = a.depth;
  fa2.reduction = a.reduction;
  fa2.tonal = a.tonal;
  fa2.whitening = a.whitening;
  fa2.gain_type = a.gain_type;
  {
    KTimer kt_(kcFinal, st);
  final_kernel<<<M, row_threads(g), 0, st>>>(g, fa2, L.Xre, L.Xim, L.noise, L.Mg, L.alpha, L.stable,
                                     L.band_p, a.wt, a.mask, L.Yre, L.Yim);
  SB_LAUNCH_CHECK();
  }
  return true;
}

}  // namespace sb
