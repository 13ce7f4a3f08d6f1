// 2D Non-Local-Means over the a-priori-SNR spectrogram S[frame][bin]: launcher and the
// kernel for the partial last paste block.
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
// The full paste blocks are sb_nlm3.cu (row-SSD sharing, two passes).  The partial last block
// (K % 8 != 0, e.g. the single bin 1600 of K = 1601) has its centre clamped to K-1 and every
// patch element clamped individually, so its patches are not aligned to the 8-bin groups of
// the main kernel: nlm_tail_kernel below evaluates it in direct form with the reference's
// accumulation order.
// (The first two generations of the main kernel -- direct form, 2.87 ms per 1333 targets, and
// row-SSD sharing with mbarrier producer / reducer warps, 1.57 ms per 4096 -- were removed in
// round 2; they are in the history of this file.)
//
// This translation unit is compiled with -ftz=true: the reference runs NLM
// under FTZ|DAZ (nlm_filter_avx.c:88, simd_utils.h:67-94).
#include "sb_common.h"

namespace sb {
namespace {

constexpr int NDT = kNlmRing;             // 21 candidate frames

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

// ---------------------------------------------------------------------------
// Tail kernel: the partial paste block [8*nfull, K) with clamped centre K-1.
// A warp per target (8 consecutive targets per CTA over one staged window), lanes over
// the distinct (dt, clamped centre) candidates, then the bins are pasted in the
// reference's flat (dt, df) order with the running sums chained through shuffles.
// ---------------------------------------------------------------------------
constexpr int TAILT = 8;     // targets (warps) per CTA
constexpr int TAILW = 32;    // staged columns: bins bc-12 .. bc+19 (clamped)

__global__ void __launch_bounds__(TAILT * 32)
nlm_tail_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
                int targets, float h, float* __restrict__ out, int pitchO) {
  // A warp per target, TAILT consecutive targets per CTA over one staged window of
  // TAILT + 20 frames; the target's own 8x8 patch stays in registers, lanes take the
  // candidates.
  __shared__ float tile[TAILT + NDT - 1][TAILW + 1];
  __shared__ float wgt[TAILT][NDT * 17];
  __shared__ float live_w[TAILT][NDT * 17];
  __shared__ unsigned short live_idx[TAILT][NDT * 17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tau0 = blockIdx.x * TAILT;
  const int nt = targets - tau0 < TAILT ? targets - tau0 : TAILT;      // targets of this CTA
  const int row0 = first_target_row + tau0 - kNlmPast;                 // frame of tile row 0
  const int bs = 8 * nfull;
  const int lim = K - bs;
  int bc = bs + 4;
  if (bc >= K) bc = K - 1;
  const int c_origin = bc - 12;
  for (int e = threadIdx.x; e < (nt + NDT - 1) * TAILW; e += blockDim.x) {
    const int r = e / TAILW, c_rel = e - r * TAILW;
    int c = c_origin + c_rel;
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    tile[r][c_rel] = S[(size_t)(row0 + r) * pitchS + c];
  }
  __syncthreads();
  if (warp >= nt) return;
  const int target = tau0 + warp;
  const float (*tw)[TAILW + 1] = &tile[warp];     // tw[r]: frame t - 16 + r of this target
  float tsum = 0.0f;
  for (int i = 0; i < lim; i++) tsum += tw[kNlmPast][bs + i - c_origin];
  const bool live = tsum >= 1e-6f;
  float inv, thr;
  block_params(bc, K, h, &inv, &thr);
  // Candidates whose clamped centre cc coincides have the same patch, hence the same distance
  // and weight: with the block centre on (or next to) the last bin, df >= K-1-bc all clamp to
  // K-1 (9 of the 17 df when K = 8 nfull + 1).  One evaluation per DISTINCT (dt, cc).
  const int cc_min = bc - kNlmRangeFreq < 0 ? 0 : bc - kNlmRangeFreq;
  const int cc_max = bc + kNlmRangeFreq > K - 1 ? K - 1 : bc + kNlmRangeFreq;
  const int ncc = cc_max - cc_min + 1;
  float* wg = wgt[warp];
  if (live) {
    float tp[8][8];                                // the target patch, element-wise clamped
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int q = 0; q < 8; q++) tp[r][q] = tw[kNlmPast - 4 + r][bc - 4 + q - c_origin];
    int dti = 0, ci = lane;
    while (ci >= ncc) { ci -= ncc; dti++; }
    while (dti < NDT) {
      const int dt = dti - kNlmPast;
      const int cc = cc_min + ci;
      float a[8];
#pragma unroll
      for (int q = 0; q < 8; q++) a[q] = 0.0f;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float* cr = tw[ring_row(dt + r - 4)] + (cc - 4 - c_origin);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float d = tp[r][q] - cr[q];
          a[q] = fmaf(d, d, a[q]);
        }
      }
      const float d = hsum8(a);
      float w = 0.0f;
      if (!(d > thr)) {
        w = schraudolph_exp(-d * inv);
        if (w < kNlmMinWeight) w = 0.0f;
      }
      wg[dti * 17 + ci] = w;
      ci += 32;
      while (ci >= ncc) { ci -= ncc; dti++; }
    }
  }
  __syncwarp();
  // compact the (few) non-zero weights, keeping the reference's flat (dt, df) order
  int n_live = 0;
  if (live) {
    for (int c0 = 0; c0 < NDT * 17; c0 += 32) {
      const int c = c0 + lane;
      float w = 0.0f;
      if (c < NDT * 17) {
        const int dti = c / 17;
        int cc = bc + (c - dti * 17) - kNlmRangeFreq;
        cc = cc < cc_min ? cc_min : (cc > cc_max ? cc_max : cc);
        w = wg[dti * 17 + cc - cc_min];
      }
      const bool on = w > 0.0f;
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int pos = n_live + __popc(m & ((1u << lane) - 1u));
        live_idx[warp][pos] = (unsigned short)c;
        live_w[warp][pos] = w;
      }
      n_live += __popc(m);
    }
  }
  __syncwarp();
  // Paste in the reference's order.  In this top-frequency block most candidates pass
  // (thr grows with (1 + 3 f^2)^2), so a lane walking the list alone spent its time in the
  // load -> FMA latency chain: the lanes fetch 32 (weight, value) pairs at once and the sums
  // are chained through shuffles (every lane carries the same running sums).
  for (int i = 0; i < lim; i++) {
    float acc = 0.0f, ws = 0.0f;
    for (int e0 = 0; e0 < n_live; e0 += 32) {
      const int e = e0 + lane;
      float w = 0.0f, v = 0.0f;
      if (e < n_live) {
        const int c = live_idx[warp][e];
        const int dti = c / 17;
        const int dfi = c - dti * 17;
        w = live_w[warp][e];
        int cb = bs + i + dfi - kNlmRangeFreq;
        cb = cb < 0 ? 0 : (cb > K - 1 ? K - 1 : cb);
        v = tw[dti][cb - c_origin];                 // frame t+dt, dt = dti-16, never wraps
      }
      const int cnt = n_live - e0 < 32 ? n_live - e0 : 32;
#pragma unroll 8
      for (int j = 0; j < cnt; j++) {
        const float wj = __shfl_sync(0xffffffffu, w, j);
        const float vj = __shfl_sync(0xffffffffu, v, j);
        acc = fmaf(wj, vj, acc);
        ws += wj;
      }
    }
    if (lane == 0) {
      const float tv = tw[kNlmPast][bs + i - c_origin];
      out[(size_t)target * pitchO + bs + i] = (ws > kNlmMinWeight) ? acc / ws : tv;
    }
  }
}

}  // namespace

bool launch_nlm(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                float* Shat, cudaStream_t st) {
  if (targets <= 0) return true;
  const int nfull = g.K / 8;
  if (g.K % 8 != 0) {
    // the tail block only reads S and writes its own output columns; it goes first because the
    // persistent main kernel occupies every SM until it ends (on a side stream it would wait)
    KTimer kt2(kcNlmTail, st);
    nlm_tail_kernel<<<(targets + TAILT - 1) / TAILT, TAILT * 32, 0, st>>>(
        S, g.Kp, g.K, nfull, first_target_row, targets, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  }
  if (nfull > 0) {
    KTimer kt(kcNlm, st);
    if (!launch_nlm3(g, S, first_target_row, targets, h, Shat, st)) return false;
  }
  return true;
}

}  // namespace sb
