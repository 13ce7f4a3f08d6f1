// 2D Non-Local-Means over the a-priori-SNR spectrogram S[frame][bin].
//
// Replaces nlm_filter_process_avx / nlm_filter_process_generic
// (nlm_filter_avx.c:87-248, nlm_filter.c:261-430) for ALL target frames of a
// call at once.  Per target frame t and 8-bin paste block:
//   for dt in [-16,4], df in [-8,8]:
//     d = SSD of the 8x8 patch centred on (t, block centre) against the patch
//         centred on (t+dt, clamp(centre+df)); patch rows go through the
//         reference's 21-slot ring, i.e. row offset x maps to
//         W(x) = ((x+16) mod 21) - 16         (nlm_filter_internal.h:74-86)
//     skip if d > 4*h_b^2; w = schraudolph_exp(-d/h_b^2); skip if w < 1e-10
//     acc[bin] += w * S[t+dt][clamp(bin+df)], wsum += w
//   out = acc/wsum (target value when wsum <= 1e-10 or block energy < 1e-6)
//
// Mapping (direct form, reference accumulation order for d):
//   CTA   = one target frame x 12 paste blocks, 252 threads
//   thread= (paste block, dt); the 8x8 target patch lives in 64 registers, the
//           17 df candidates are evaluated 4 at a time from three aligned
//           LDS.128 per patch row (12 floats cover df0..df0+3), 8 "lane"
//           accumulators per candidate FMA'd over the rows and reduced with
//           the AVX horizontal-sum tree ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7))
//           (simd_utils.h:217-276).
//   smem  = 21 ring rows x (12+3) column groups; every group of 8 bins sits in
//           a 12-float slot so that consecutive paste blocks are 48 B apart and
//           a quarter-warp's LDS.128 hit 32 distinct banks.  Halo columns are
//           edge-replicated, which reproduces clamp_index() for every element.
// The (rare) partial last block (K % 8 != 0) is handled by a small tail kernel.
//
// This translation unit is compiled with -ftz=true: the reference runs NLM
// under FTZ|DAZ (nlm_filter_avx.c:88, simd_utils.h:67-94).
#include "sb_common.h"

namespace sb {
namespace {

constexpr int BT = 12;                    // paste blocks per CTA
constexpr int NDT = kNlmRing;             // 21 candidate frames
constexpr int NTHREADS = BT * NDT;        // 252
constexpr int GROUPS = BT + 3;            // 8-bin column groups per tile row (8 left, 16 right halo)
constexpr int PITCH = 208;                // floats; >= 12*GROUPS and == 16 (mod 32)
static_assert(PITCH >= 12 * GROUPS, "tile pitch");

__device__ __forceinline__ int ring_row(int x) {   // x in [-20, 7] -> smem row of t+W(x)
  return (x + 16 + NDT) % NDT;
}

__device__ __forceinline__ float schraudolph_exp(float x) {   // simd_utils.h:670-684
  if (x < -20.0f) return 0.0f;
  return __uint_as_float(__float2uint_rz(fmaf(12102203.0f, x, 1064866805.0f)));
}

__device__ __forceinline__ float hsum8(const float* a) {
  return ((a[0] + a[4]) + (a[1] + a[5])) + ((a[2] + a[6]) + (a[3] + a[7]));
}

// tile-relative column -> float offset inside a tile row
__device__ __forceinline__ int tcol(int c_rel) { return 12 * (c_rel >> 3) + (c_rel & 7); }

__device__ __forceinline__ void block_params(int bc, int K, float h, float* inv, float* thr) {
  // nlm_filter_avx.c:131-147
  float h2 = h * h;
  if (K > 1) {
    const float nf = (float)bc / (float)(K - 1);
    const float fs = 1.0f + kNlmFreqScale * nf * nf;
    const float hv = h * fs;
    h2 = hv * hv;
  }
  *inv = 1.0f / h2;
  *thr = kNlmThrMul * h2;
}

__global__ void __launch_bounds__(NTHREADS, 2)
nlm_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
           float h, float* __restrict__ out, int pitchO) {
  __shared__ __align__(16) float tile[NDT * PITCH];
  __shared__ float part[BT][NDT][9];

  const int target = blockIdx.y;
  const int row_t = first_target_row + target;     // row of the target frame inside S
  const int blk0 = blockIdx.x * BT;
  const int c_origin = blk0 * 8 - 8;               // global bin of tile column 0

  // ---- stage the 21 ring rows (offsets -16..+4), edge-replicated columns ----
  for (int e = threadIdx.x; e < NDT * GROUPS * 8; e += NTHREADS) {
    const int r = e / (GROUPS * 8);
    const int c_rel = e - r * (GROUPS * 8);
    int c = c_origin + c_rel;
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    tile[r * PITCH + tcol(c_rel)] = S[(size_t)(row_t - kNlmPast + r) * pitchS + c];
  }
  __syncthreads();

  const int dti = threadIdx.x / BT;                // 0..20
  const int bl = threadIdx.x - dti * BT;           // local paste block
  const int b = blk0 + bl;
  const int dt = dti - kNlmPast;
  float pacc[8];
  float pw = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) pacc[i] = 0.0f;

  if (b < nfull) {
    const int bs = 8 * b;
    const int bc = bs + 4;
    const int cb = 8 + 8 * bl;                     // tile column of bin bs
    // target patch: ring rows of offsets -4..+3, bins bs..bs+7
    float tg[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float4* p = reinterpret_cast<const float4*>(&tile[(kNlmPast - 4 + r) * PITCH + tcol(cb)]);
      const float4 v0 = p[0], v1 = p[1];
      tg[r][0] = v0.x; tg[r][1] = v0.y; tg[r][2] = v0.z; tg[r][3] = v0.w;
      tg[r][4] = v1.x; tg[r][5] = v1.y; tg[r][6] = v1.z; tg[r][7] = v1.w;
    }
    // block energy gate (nlm_filter_avx.c:123-129); target row = ring offset 0
    float tsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) tsum += tg[4][i];
    if (tsum >= 1e-6f) {
      float inv, thr;
      block_params(bc, K, h, &inv, &thr);
      int rows[8];
#pragma unroll
      for (int r = 0; r < 8; r++) rows[r] = ring_row(dt + r - 4) * PITCH;
      const float* crow = &tile[ring_row(dt) * PITCH];   // pasted frame t+dt

      // distance of the candidate whose centre clamps to bin 0 (only block 0)
      float dsub_lo = 0.0f, dsub_hi = 0.0f;
      if (bc < kNlmRangeFreq) {
        float a[8];
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          // candidate centre 0 -> bins -4..3 -> tile columns cb-bs-4 .. (16-byte aligned: cb-bs = 8)
          const float4* p = reinterpret_cast<const float4*>(&tile[rows[r] + tcol(cb - bs - 4)]);
          const float4 v0 = p[0];
          const float4 v1 = *reinterpret_cast<const float4*>(&tile[rows[r] + tcol(cb - bs)]);
          const float c[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float d = tg[r][q] - c[q];
            a[q] = fmaf(d, d, a[q]);
          }
        }
        dsub_lo = hsum8(a);
      }

#pragma unroll 1
      for (int g = 0; g < 5; g++) {
        const int df0 = -kNlmRangeFreq + 4 * g;
        const int ncand = (g == 4) ? 1 : 4;
        float acc[4][8];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
          for (int q = 0; q < 8; q++) acc[c][q] = 0.0f;
        const int col0 = cb + df0;                 // tile column of bin bs+df0 (multiple of 4)
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const float* rp = &tile[rows[r]];
          const float4 v0 = *reinterpret_cast<const float4*>(rp + tcol(col0));
          const float4 v1 = *reinterpret_cast<const float4*>(rp + tcol(col0 + 4));
          float4 v2 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g < 4) v2 = *reinterpret_cast<const float4*>(rp + tcol(col0 + 8));
          const float seg[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w,
                                 v2.x, v2.y, v2.z, v2.w};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            if (c < ncand) {
#pragma unroll
              for (int q = 0; q < 8; q++) {
                const float d = tg[r][q] - seg[c + q];
                acc[c][q] = fmaf(d, d, acc[c][q]);
              }
            }
          }
        }
        // weights + paste, candidates in df order
#pragma unroll
        for (int c = 0; c < 4; c++) {
          if (c < ncand) {
            const int df = df0 + c;
            float d = hsum8(acc[c]);
            const int cc = bc + df;
            if (cc < 0) d = dsub_lo;               // centre clamps to 0: same patch as df = -bc
            else if (cc > K - 1) d = dsub_hi;      // centre clamps to K-1
            else if (cc == K - 1) dsub_hi = d;
            if (!(d > thr)) {
              const float w = schraudolph_exp(-d * inv);
              if (!(w < kNlmMinWeight)) {
                const float* cp = crow + 0;
#pragma unroll
                for (int i = 0; i < 8; i++) pacc[i] = fmaf(w, cp[tcol(cb + i + df)], pacc[i]);
                pw += w;
              }
            }
          }
        }
      }
    }
  }

  // ---- reduce over dt in reference order (-16 .. +4) and normalise ----
  if (b < nfull) {
#pragma unroll
    for (int i = 0; i < 8; i++) part[bl][dti][i] = pacc[i];
    part[bl][dti][8] = pw;
  }
  __syncthreads();
  if (threadIdx.x < BT * 8) {
    const int bl2 = threadIdx.x >> 3;
    const int i = threadIdx.x & 7;
    const int b2 = blk0 + bl2;
    if (b2 < nfull) {
      float acc = 0.0f, ws = 0.0f;
#pragma unroll
      for (int d = 0; d < NDT; d++) {
        acc += part[bl2][d][i];
        ws += part[bl2][d][8];
      }
      const int bin = 8 * b2 + i;
      const float tv = tile[kNlmPast * PITCH + tcol(8 + 8 * bl2 + i)];
      out[(size_t)target * pitchO + bin] = (ws > kNlmMinWeight) ? acc / ws : tv;   // nlm_filter_avx.c:237-243
    }
  }
}

// ---------------------------------------------------------------------------
// Tail kernel: the partial paste block [8*nfull, K) with clamped centre K-1.
// One CTA per target, one thread per (dt, df) candidate, then bins are pasted
// by `lim` threads walking the 357 weights in the reference's flat order.
// ---------------------------------------------------------------------------
constexpr int TAILW = 32;    // staged columns: bins bc-12 .. bc+19 (clamped)

__global__ void __launch_bounds__(384)
nlm_tail_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
                float h, float* __restrict__ out, int pitchO) {
  __shared__ float tile[NDT][TAILW + 1];
  __shared__ float wgt[NDT * 17];
  const int target = blockIdx.x;
  const int row_t = first_target_row + target;
  const int bs = 8 * nfull;
  const int lim = K - bs;
  int bc = bs + 4;
  if (bc >= K) bc = K - 1;
  const int c_origin = bc - 12;
  for (int e = threadIdx.x; e < NDT * TAILW; e += blockDim.x) {
    const int r = e / TAILW, c_rel = e - r * TAILW;
    int c = c_origin + c_rel;
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    tile[r][c_rel] = S[(size_t)(row_t - kNlmPast + r) * pitchS + c];
  }
  __syncthreads();
  float tsum = 0.0f;
  for (int i = 0; i < lim; i++) tsum += tile[kNlmPast][bs + i - c_origin];
  const bool live = tsum >= 1e-6f;
  float inv, thr;
  block_params(bc, K, h, &inv, &thr);
  if (threadIdx.x < NDT * 17) {
    float w = 0.0f;
    if (live) {
      const int dti = threadIdx.x / 17;
      const int df = threadIdx.x - dti * 17 - kNlmRangeFreq;
      const int dt = dti - kNlmPast;
      int cc = bc + df;
      cc = cc < 0 ? 0 : (cc > K - 1 ? K - 1 : cc);
      float a[8];
#pragma unroll
      for (int q = 0; q < 8; q++) a[q] = 0.0f;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float* tr = tile[kNlmPast - 4 + r];
        const float* cr = tile[ring_row(dt + r - 4)];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float d = tr[bc - 4 + q - c_origin] - cr[cc - 4 + q - c_origin];
          a[q] = fmaf(d, d, a[q]);
        }
      }
      const float d = hsum8(a);
      if (!(d > thr)) {
        w = schraudolph_exp(-d * inv);
        if (w < kNlmMinWeight) w = 0.0f;
      }
    }
    wgt[threadIdx.x] = w;
  }
  __syncthreads();
  if (threadIdx.x < lim) {
    const int i = threadIdx.x;
    float acc = 0.0f, ws = 0.0f;
    for (int dti = 0; dti < NDT; dti++) {
      const float* cr = tile[dti];   // frame t+dt, dt = dti-16, never wraps
      for (int dfi = 0; dfi < 17; dfi++) {
        const float w = wgt[dti * 17 + dfi];
        if (w > 0.0f) {
          int cb = bs + i + dfi - kNlmRangeFreq;
          cb = cb < 0 ? 0 : (cb > K - 1 ? K - 1 : cb);
          acc = fmaf(w, cr[cb - c_origin], acc);
          ws += w;
        }
      }
    }
    const float tv = tile[kNlmPast][bs + i - c_origin];
    out[(size_t)target * pitchO + bs + i] = (ws > kNlmMinWeight) ? acc / ws : tv;
  }
}

}  // namespace

bool launch_nlm(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                float* Shat, cudaStream_t st) {
  if (targets <= 0) return true;
  const int nfull = g.K / 8;
  if (nfull > 0) {
    dim3 grid((nfull + BT - 1) / BT, targets);
    nlm_kernel<<<grid, NTHREADS, 0, st>>>(S, g.Kp, g.K, nfull, first_target_row, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  }
  if (g.K % 8 != 0) {
    nlm_tail_kernel<<<targets, 384, 0, st>>>(S, g.Kp, g.K, nfull, first_target_row, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  }
  return true;
}

}  // namespace sb
