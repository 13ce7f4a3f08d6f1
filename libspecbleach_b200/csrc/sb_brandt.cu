// Brandt trimmed-mean noise estimator (brandt_noise_estimator.c:134-299) for all
// frames of a call.
//
// The reference keeps, per bin, a ring of the last H non-silent power values
// (H = 5000 ms / (fft_size*1000/sr*0.5), quirk Q6), sorts it every frame and, for
// p in {.1,.25,.5,.75,1}, fits an exponential to the q = p*H smallest values
// (trimmed mean x bias factor) and scores the fit with an Anderson-Darling-like
// L1 distance between the empirical and the truncated-exponential CDF.  The best
// fit replaces the bin's noise estimate when its confidence 1 - ad >= 0.90,
// otherwise the previous estimate is held.
//
// Whole-call form: the sorted window of a frame depends only on the H most
// recent non-silent frames, so frames are evaluated independently:
//   1. compact the non-silent frames of the call (silence gate, :138-143,169-173)
//   2. brandt_eval_kernel: one warp per (bin, tile of 64 consecutive non-silent
//      frames).  The tile's first window is rank-sorted from scratch; every next
//      window is the previous one with the oldest value removed and the new one
//      inserted (one scatter pass through shared memory).  The five fits are
//      evaluated lane-parallel over the sorted values.
//   3. brandt_hold_kernel: thread per bin, "keep the last accepted estimate"
//      over the frames of the call (the only time recursion of the estimator).
//   4. brandt_state_kernel: roll the per-handle ring forward.
// apply_floor (:284-299) clamps the WHOLE history after every frame, so every
// value except the one inserted by the current frame is read as
// max(value, floor/factor); the frame that follows a re-seed still sees the
// un-clamped jittered seed, exactly as the reference does.
//
// Sums over the sorted values are warp-tree reductions (the reference's are a
// gcc-vectorised sum and a scalar chain); both differ from exact arithmetic by
// ~1e-7 relative, see DESIGN.md.
#include "sb_common.h"

#include <math.h>

namespace sb {

constexpr int kBrandtTile = 64;        // consecutive non-silent frames per warp
constexpr int kBrandtHdrIndex = 3;     // est[3] = history_index
constexpr int kEstHeaderB = 32;        // same header size as sb_adaptive.cu

struct BrandtCfg {
  int H;
  int q[5];
  float cf[5];        // correction factor per candidate (:100-103)
  float inv_cf;       // 1 / correction_factor(0.5) (:254, :289)
};

static float correction_factor(float p) {   // brandt_noise_estimator.c:44-56
  if (p <= 0.0f || p >= 1.0f) return 1.0f;
  const float term = (1.0f - p) / p * logf(1.0f - p);
  const float den = 1.0f + term;
  if (fabsf(den) < 1e-6f) return 1.0f;
  return 1.0f / den;
}

static BrandtCfg brandt_cfg(const GeoK& g) {
  BrandtCfg c;
  // brandt_noise_estimator.c:77-92
  const float ms_per_frame = (float)g.N * 1000.0f / (float)g.sr;
  float frame_duration = ms_per_frame * 0.5f;
  if (frame_duration < 0.1f) frame_duration = 0.1f;
  unsigned H = (unsigned)(5000.0f / frame_duration);
  if (H < 5u) H = 5u;
  c.H = (int)H;
  static const float pc[5] = {0.1f, 0.25f, 0.5f, 0.75f, 1.0f};
  for (int i = 0; i < 5; i++) {
    c.q[i] = (int)(unsigned)(pc[i] * (float)H);
    c.cf[i] = correction_factor(pc[i]);
  }
  c.inv_cf = 1.0f / correction_factor(0.5f);
  return c;
}

int brandt_history(const GeoK& g) { return brandt_cfg(g).H; }

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Compaction of the non-silent frames: idx[0] = count, idx[1 + n] = frame of ordinal n.
__global__ void __launch_bounds__(1024)
brandt_compact_kernel(const float* __restrict__ energy, int frames, int* __restrict__ idx) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int f0 = 0; f0 < frames; f0 += 1024) {
    const int f = f0 + threadIdx.x;
    const bool live = f < frames && !(energy[f] < kSilenceThreshold);
    const unsigned m = __ballot_sync(0xffffffffu, live);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int before = 0;
    for (int i = 0; i < w; i++) before += warp_tot[i];
    if (live) idx[1 + base + before + __popc(m & ((1u << lane) - 1u))] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < 32; i++) t += warp_tot[i];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) idx[0] = base;
}

__global__ void brandt_seed_kernel(const GeoK g, BrandtCfg c, const float* __restrict__ floorv,
                                   float* __restrict__ est) {
  // set_state / update_seed, :252-282 (history_index is NOT reset)
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (k >= g.K) return;
  float* last = est + kEstHeaderB;
  float* hist = last + g.Kp;
  const float p = floorv[k];
  const float val = __fmul_rn(p, c.inv_cf);
  const int j = ((t + k) % 11) - 5;
  const float jitter = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(0.01f, (float)j), 5.0f));
  hist[(size_t)t * g.Kp + k] = __fmul_rn(val, jitter);
  if (t == 0) last[k] = p;
  if (t == 0 && k == 0) est[0] = 0.f;   // is_first_frame = false
}

// value of extended row r for bin k: rows [0,H) = ring, oldest first; rows H+n = the
// n-th non-silent frame of this call.
__device__ __forceinline__ float ext_value(const float* __restrict__ hist,
                                           const float* __restrict__ P,
                                           const int* __restrict__ nz, int H, int idx0, int Kp,
                                           int k, int r) {
  if (r < H) {
    int slot = idx0 + r;
    if (slot >= H) slot -= H;
    return hist[(size_t)slot * Kp + k];
  }
  return P[(size_t)nz[r - H] * Kp + k];
}

// dst = src with one occurrence of x_old removed and x_new inserted (both sorted ascending)
__device__ __forceinline__ void sorted_replace(const float* src, float* dst, int H, float x_old,
                                               float x_new, int lane) {
  int c_lt_old = 0, c_le_new = 0;
  for (int i0 = 0; i0 < H; i0 += 32) {
    const int i = i0 + lane;
    const float v = i < H ? src[i] : 0.f;
    c_lt_old += __popc(__ballot_sync(0xffffffffu, i < H && v < x_old));
    c_le_new += __popc(__ballot_sync(0xffffffffu, i < H && v <= x_new));
  }
  const int po = c_lt_old;                                   // first occurrence of x_old
  const int pv = c_le_new - ((x_old <= x_new) ? 1 : 0);      // position of x_new
  for (int i = lane; i < H; i += 32) {
    if (i == po) continue;
    const float v = src[i];
    dst[i - (i > po ? 1 : 0) + (v > x_new ? 1 : 0)] = v;
  }
  if (lane == 0) dst[pv] = x_new;
  __syncwarp();
}

__global__ void __launch_bounds__(256)
brandt_eval_kernel(const GeoK g, BrandtCfg c, const float* __restrict__ P,
                   const float* __restrict__ floorv, const float* __restrict__ est,
                   const int* __restrict__ idx, int fresh_seed, int warps_per_block,
                   float* __restrict__ out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * warps_per_block + warp;
  const int nnz = idx[0];
  const int* nz = idx + 1;
  const int n0 = blockIdx.y * kBrandtTile;
  if (k >= g.K || n0 >= nnz) return;
  const int H = c.H;
  const int nT = (nnz - n0 < kBrandtTile) ? nnz - n0 : kBrandtTile;
  float* chron = sm + (size_t)warp * (3 * H + kBrandtTile);
  float* bufA = chron + H + kBrandtTile;
  float* bufB = bufA + H;
  const float* hist = est + kEstHeaderB + g.Kp;
  const int idx0 = (int)est[kBrandtHdrIndex];
  const float flv = __fmul_rn(floorv[k], c.inv_cf);           // apply_floor value (:289-291)
  // the frame right after a re-seed sees the jittered seed un-clamped
  const bool raw_first = fresh_seed && nz[0] == 0;

  // chronological values of the tile: extended rows n0+1 .. n0+nT-1+H (raw)
  const int nchron = H + nT - 1;
  for (int r = lane; r < nchron; r += 32)
    chron[r] = ext_value(hist, P, nz, H, idx0, g.Kp, k, n0 + 1 + r);
  __syncwarp();

  float* cur = bufA;
  float* oth = bufB;
  bool need_sort = true;
  for (int t = 0; t < nT; t++) {
    const int n = n0 + t;
    const bool raw_old = raw_first && n == 0;
    if (need_sort) {
      // rank sort of the window chron[t .. t+H): stable, ties broken by position
      for (int i = lane; i < H; i += 32) {
        float vi = chron[t + i];
        if (i < H - 1 && !raw_old) vi = fmaxf(vi, flv);
        int rank = 0;
        for (int j = 0; j < H; j++) {
          float vj = chron[t + j];
          if (j < H - 1 && !raw_old) vj = fmaxf(vj, flv);
          rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0;
        }
        cur[rank] = vi;
      }
      __syncwarp();
      need_sort = false;
    } else {
      // previous newest value becomes a clamped history value ...
      const float prev_raw = chron[t + H - 2];
      if (prev_raw < flv) {
        sorted_replace(cur, oth, H, prev_raw, flv, lane);
        float* tmp = cur; cur = oth; oth = tmp;
      }
      // ... the oldest one leaves, the new raw value enters
      const float x_old = fmaxf(chron[t - 1], flv);
      sorted_replace(cur, oth, H, x_old, chron[t + H - 1], lane);
      float* tmp = cur; cur = oth; oth = tmp;
    }
    if (raw_old) need_sort = true;

    // ---- the five trimmed-exponential fits (:198-241, calculate_ad_norm :113-132) ----
    // pass 1: trimmed sums.  (i0, ci) pairs with i0 >= q[ci] are skipped warp-uniformly
    float ps[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < H; i0 += 32) {
      const int i = i0 + lane;
      const float v = i < H ? cur[i] : 0.f;
#pragma unroll
      for (int ci = 0; ci < 5; ci++)
        if (i0 < c.q[ci] && i < c.q[ci]) ps[ci] += v;
    }
    float tot1[5];
#pragma unroll
    for (int ci = 0; ci < 5; ci++) tot1[ci] = warp_sum(ps[ci]);
    // per-fit scalars: lane ci computes fit ci (IEEE divisions, one expf), then broadcast
    float s_mu = 0.f, s_mu_inv = 0.f, s_den_inv = 0.f, s_q_inv = 0.f;
    int s_flags = 0;                                   // bit 0: ok, bit 1: live
    {
      float t = tot1[0];
      int q = c.q[0];
      float cf = c.cf[0];
#pragma unroll
      for (int ci = 1; ci < 5; ci++)
        if (lane == ci) { t = tot1[ci]; q = c.q[ci]; cf = c.cf[ci]; }
      const float mu_trunc = __fdiv_rn(t, (float)q);
      const bool okc = q >= 10 && mu_trunc > kSilenceThreshold;
      s_mu = __fmul_rn(mu_trunc, cf);
      bool livec = okc && !(s_mu < 1e-15f);
      s_mu_inv = __fdiv_rn(1.0f, s_mu);
      s_q_inv = __fdiv_rn(1.0f, (float)q);
      const float b = cur[q > 0 ? q - 1 : 0];
      const float den = 1.0f - expf(-__fmul_rn(b, s_mu_inv));
      if (fabsf(den) < kSpectralEps) livec = false;
      s_den_inv = __fdiv_rn(1.0f, den);
      s_flags = (okc ? 1 : 0) | (livec ? 2 : 0);
    }
    float mu[5], mu_inv[5], den_inv[5], q_inv[5];
    bool ok[5], live[5];
#pragma unroll
    for (int ci = 0; ci < 5; ci++) {
      mu[ci] = __shfl_sync(0xffffffffu, s_mu, ci);
      mu_inv[ci] = __shfl_sync(0xffffffffu, s_mu_inv, ci);
      den_inv[ci] = __shfl_sync(0xffffffffu, s_den_inv, ci);
      q_inv[ci] = __shfl_sync(0xffffffffu, s_q_inv, ci);
      const int fl = __shfl_sync(0xffffffffu, s_flags, ci);
      ok[ci] = (fl & 1) != 0;
      live[ci] = (fl & 2) != 0;
    }
    // pass 2: L1 distance between the empirical and the truncated-exponential CDF
    float ts[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i0 = 0; i0 < H; i0 += 32) {
      const int i = i0 + lane;
      const float v = i < H ? cur[i] : 0.f;
      const float fi = (float)(i + 1);
#pragma unroll
      for (int ci = 0; ci < 5; ci++) {
        if (i0 < c.q[ci] && live[ci]) {                // warp-uniform
          if (i < c.q[ci]) {
            const float e1 = 1.0f - expf(-__fmul_rn(v, mu_inv[ci]));
            const float fe = __fmul_rn(fi, q_inv[ci]);
            ts[ci] += fabsf(fmaf(-e1, den_inv[ci], fe));   // gcc emits one vfnmadd here
          }
        }
      }
    }
    float min_ad = 2.0f, best = 0.f;
#pragma unroll
    for (int ci = 0; ci < 5; ci++) {
      const float tot = warp_sum(ts[ci]);
      const float ad = live[ci] ? __fmul_rn(tot + tot, q_inv[ci]) : 1.0f;
      if (ok[ci] && ad < min_ad) {
        min_ad = ad;
        best = mu[ci];
      }
    }
    // 1 - min_ad >= 0.90 as compiled: !(1.0f - 0.9f < min_ad)
    const bool accept = !((1.0f - 0.9f) < min_ad);
    if (lane == 0) out[(size_t)nz[n] * g.Kp + k] = accept ? best : -1.0f;
    __syncwarp();
  }
}

// out[f][k]: accepted estimate or -1 (non-silent frames); silent frames hold as well.
__global__ void brandt_hold_kernel(const GeoK g, const float* __restrict__ energy, int frames,
                                   float* __restrict__ est, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.K) return;
  float* lastp = est + kEstHeaderB;
  float last = lastp[k];
  constexpr int U = 8;
  for (int f0 = 0; f0 < frames; f0 += U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int f = f0 + u;
      v[u] = (f < frames && !(energy[f] < kSilenceThreshold)) ? out[(size_t)f * g.Kp + k] : -1.0f;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int f = f0 + u;
      if (f < frames) {
        if (v[u] >= 0.f) last = v[u];
        out[(size_t)f * g.Kp + k] = last;
      }
    }
  }
  lastp[k] = last;
}

// ring slot j <- newest value written to it by this call (clamped), else clamp in place
__global__ void brandt_state_kernel(const GeoK g, BrandtCfg c, const float* __restrict__ P,
                                    const float* __restrict__ floorv,
                                    const int* __restrict__ idx, float* __restrict__ est) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k >= g.K) return;
  const int H = c.H;
  const int nnz = idx[0];
  const int idx0 = (int)est[kBrandtHdrIndex];
  float* hist = est + kEstHeaderB + g.Kp;
  const float flv = __fmul_rn(floorv[k], c.inv_cf);
  int m = j - idx0;
  if (m < 0) m += H;
  float v;
  if (m < nnz) {
    const int n = m + H * ((nnz - 1 - m) / H);
    v = P[(size_t)idx[1 + n] * g.Kp + k];
  } else {
    v = hist[(size_t)j * g.Kp + k];
  }
  hist[(size_t)j * g.Kp + k] = fmaxf(v, flv);
}

__global__ void brandt_commit_kernel(BrandtCfg c, const int* __restrict__ idx,
                                     float* __restrict__ est) {
  const int idx0 = (int)est[kBrandtHdrIndex];
  est[kBrandtHdrIndex] = (float)((idx0 + idx[0]) % c.H);
}

}  // namespace

size_t brandt_estimator_floats(const GeoK& g) {
  return kEstHeaderB + (size_t)(1 + brandt_cfg(g).H) * g.Kp;
}

bool brandt_supported(const GeoK& g) {
  const BrandtCfg c = brandt_cfg(g);
  return sizeof(float) * (size_t)(3 * c.H + kBrandtTile) <= 200 * 1024;
}

bool launch_brandt_seed(const GeoK& g, const float* floorv, float* est, cudaStream_t st) {
  const BrandtCfg c = brandt_cfg(g);
  dim3 grid((g.K + 255) / 256, c.H);
  brandt_seed_kernel<<<grid, 256, 0, st>>>(g, c, floorv, est);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_brandt(const GeoK& g, const float* P, int frames, const float* floorv, float* est,
                   const float* energy, int* idx, bool fresh_seed, float* noise_rows,
                   cudaStream_t st) {
  const BrandtCfg c = brandt_cfg(g);
  brandt_compact_kernel<<<1, 1024, 0, st>>>(energy, frames, idx);
  SB_LAUNCH_CHECK();
  const size_t per_warp = sizeof(float) * (size_t)(3 * c.H + kBrandtTile);
  int wpb = (int)((200 * 1024) / per_warp);
  if (wpb > 8) wpb = 8;
  if (wpb < 1) {
    set_error_msg("specbleach-b200: Brandt history too long for this frame size");
    return false;
  }
  const size_t smem = per_warp * (size_t)wpb;
  // the attribute belongs to the (device, context) pair: set it on every call, as launch_nlm3
  // does, so that handles on several GPUs of one process all get it
  if (smem > 48 * 1024)
    SB_CUDA(cudaFuncSetAttribute((const void*)brandt_eval_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.K + wpb - 1) / wpb, (frames + kBrandtTile - 1) / kBrandtTile);
  brandt_eval_kernel<<<grid, 32 * wpb, smem, st>>>(g, c, P, floorv, est, idx, fresh_seed ? 1 : 0,
                                                   wpb, noise_rows);
  SB_LAUNCH_CHECK();
  brandt_hold_kernel<<<(g.K + 63) / 64, 64, 0, st>>>(g, energy, frames, est, noise_rows);
  SB_LAUNCH_CHECK();
  dim3 sgrid((g.K + 255) / 256, c.H);
  brandt_state_kernel<<<sgrid, 256, 0, st>>>(g, c, P, floorv, idx, est);
  SB_LAUNCH_CHECK();
  brandt_commit_kernel<<<1, 1, 0, st>>>(c, idx, est);
  SB_LAUNCH_CHECK();
  return true;
}

}  // namespace sb
