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

#include <stdlib.h>

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


// ===========================================================================
// NLM v2: row-SSD sharing.
//
// d(t, dt, df) = sum_{r=0..7} R[t+r-4][partner][df],
//   R[a][p][df] = sum_{q<8} (S[a][bs+q] - S[p][bs+df+q])^2,
//   partner p = t + W(dt+r-4) = a + dt           when dt+r-4 lies in [-16, 4]
//                             = a + dt -/+ 21    otherwise (ring alias, quirk Q2)
// A row-pair term R[a][a+dt] belongs to 8 consecutive targets, so a thread that
// walks the rows a = t0-4, t0-3, ... of a tile of consecutive targets computes
// each term once and adds it to the (up to) 8 targets in flight: ~24 instead of
// 128 FP instructions per candidate.  The wrapped rows (7 of the 21 dt) need the
// second lag dt-/+21 for some rows of each target; those warps compute both.
//
//   CTA   = up to 64 consecutive targets x 6 paste blocks; 21 producer warps (one
//           per dt: control flow is warp-uniform) + 2 reducer warps
//   lane  = (paste block, df group); df groups = [-8,-5] [-4,-1] [0,3] [4,7] [8]:
//           three aligned LDS.128 (12 bins) per row serve 4 candidates
//   state = 8 targets in flight x 4 candidates partial sums in registers
//   paste = rare for noise-like maps: a thread that accepted candidates for the
//           target completing at this row writes its 8+1 partial sums to a
//           private slot of a 4-deep ring and sets a bit; the reducer warps sum
//           the flagged slots in (dt, df) order -> deterministic and independent
//           of how targets are tiled (streaming stays bit-exact).
//   sync  = no block barrier in the row loop: producers signal "target complete"
//           on an mbarrier per ring slot (one arrival per warp), reducers hand
//           the slot back on a second mbarrier, so slow (wrapped) warps and the
//           reducers never stall the other warps for more than the ring depth.
// The floating-point order of d differs from the reference (columns first, then
// rows) by ~1e-7 relative; see DESIGN.md.
// ===========================================================================
constexpr int B2 = 6;                       // paste blocks per CTA
constexpr int G2 = 5;                       // df groups per (block, dt)
constexpr int TT2 = 64;                     // targets per CTA
constexpr int ROWS2 = TT2 + 20;             // staged rows: t0-16 .. t0+TT2+3
constexpr int GRP2 = B2 + 4;                // column groups: 8 bins left, 24 right
constexpr int PITCH2 = 8 * GRP2;            // 80 floats, plain layout
constexpr int NRED2 = 2;                    // reducer warps
constexpr int NTH2 = (NDT + NRED2) * 32;    // 736 threads
constexpr int SLOTS2 = NDT * G2;            // 105 contributors per (target, block)
constexpr int MW2 = 4;                      // mask words per (ring slot, block)
constexpr int RING2 = 4;                    // ring depth (targets in hand-over)
constexpr size_t SMEM2_TILE = sizeof(float) * (size_t)ROWS2 * PITCH2;
constexpr size_t SMEM2_PART = sizeof(float) * RING2 * B2 * SLOTS2 * 9;
constexpr size_t SMEM2_MASK = sizeof(unsigned) * RING2 * B2 * MW2;
constexpr size_t SMEM2_GATE = sizeof(unsigned) * TT2;            // 6 live bits per target
constexpr size_t SMEM2_BAR = sizeof(unsigned long long) * 2 * RING2;
constexpr size_t SMEM2 = SMEM2_TILE + SMEM2_PART + SMEM2_MASK + SMEM2_GATE + SMEM2_BAR;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waiting warps sleep between polls: a spinning warp would keep winning the issue
// arbiter against the (slower) warps it is waiting for.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  if (mbar_try(bar, parity)) return;
  unsigned ns = 64;
  while (!mbar_try(bar, parity)) {
    __nanosleep(ns);
    if (ns < 512) ns <<= 1;
  }
}

// producer warp id -> dt index (dt + 16); slow (wrapped) dt indices 0,1,2,3,18,19,20
__constant__ int c_dt_of_warp[NDT] = {4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14,
                                      15, 16, 0,  1,  2,  17, 3,  18, 19, 20};

struct Nlm2Thread {
  const float* tile;
  int col_tg;       // tile column of bin bs
  int col_seg;      // tile column of bin bs + df0
  int df0;          // df of candidate 0 of this lane
  int nc;           // real candidates of this lane: 4, or 1 for the df = 8 lane
  int dt;
  int rlo, rhi;     // rows r < rlo or r >= rhi use the wrapped partner
  int wrap_shift;   // +21 / -21, 0 when this dt never wraps
  float inv, thr;
  int bc;           // centre bin of the paste block
  int hi_df;        // df whose centre is exactly bin K-1 (>= 9: no clamping)
  int lane_lo, c_lo;   // lane / candidate holding the distance for centres clamped to 0
  int lane_hi, c_hi;   // ... clamped to K-1
  int bl, slot_id;
  bool active, edge;
};

__device__ __forceinline__ void nlm2_rowterms(const Nlm2Thread& T, const float (&tg)[8], int row,
                                              float (&R)[4]) {
  row = row < 0 ? 0 : (row > ROWS2 - 1 ? ROWS2 - 1 : row);
  const float4* ps = reinterpret_cast<const float4*>(T.tile + row * PITCH2 + T.col_seg);
  const float4 s0 = ps[0], s1 = ps[1], s2 = ps[2];
  const float seg[12] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float r = 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const float d = tg[q] - seg[c + q];
      r = fmaf(d, d, r);
    }
    R[c] = r;
  }
}

template <int PHI>
__device__ __forceinline__ void nlm2_step(const Nlm2Thread& T, int i, int nT, float (&acc)[8][4],
                                          float* part, unsigned* pmask, const unsigned* gate,
                                          unsigned long long* bar_full,
                                          unsigned long long* bar_empty, int lane) {
  // ---- row a = t0-4+i: target-side bins, row terms against the partner rows ----
  const int row_a = 12 + i;
  const float4* pt = reinterpret_cast<const float4*>(T.tile + row_a * PITCH2 + T.col_tg);
  const float4 t0 = pt[0], t1 = pt[1];
  const float tg[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
  float Rm[4];
  nlm2_rowterms(T, tg, row_a + T.dt, Rm);
  // ---- add them to the targets in flight: slot (PHI - r) & 7 holds target i - r ----
  if (T.wrap_shift == 0) {                          // warp-uniform
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int s = (PHI - r) & 7;
#pragma unroll
      for (int c = 0; c < 4; c++) acc[s][c] = (r == 0) ? Rm[c] : acc[s][c] + Rm[c];
    }
  } else {
    float Rw[4];
    nlm2_rowterms(T, tg, row_a + T.dt + T.wrap_shift, Rw);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int s = (PHI - r) & 7;
      const bool w = (r < T.rlo) || (r >= T.rhi);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const float v = w ? Rw[c] : Rm[c];
        acc[s][c] = (r == 0) ? v : acc[s][c] + v;
      }
    }
  }
  // ---- target tau = i-7 completes with this row (r = 7) ----
  const int tau = i - 7;
  if (tau >= 0 && tau < nT) {                       // block-uniform
    const int s7 = (PHI - 7) & 7;
    float d[4];
#pragma unroll
    for (int c = 0; c < 4; c++) d[c] = acc[s7][c];
    if (T.edge) {                                   // block-uniform: first / last block tile
      // candidates whose centre clamps to bin 0 / K-1 take the distance of the
      // candidate that sits exactly on the edge (clamp_index BEFORE the patch)
      const float sl = T.c_lo == 0 ? d[0] : (T.c_lo == 1 ? d[1] : (T.c_lo == 2 ? d[2] : d[3]));
      const float sh = T.c_hi == 0 ? d[0] : (T.c_hi == 1 ? d[1] : (T.c_hi == 2 ? d[2] : d[3]));
      const float d_lo = __shfl_sync(0xffffffffu, sl, T.lane_lo);
      const float d_hi = __shfl_sync(0xffffffffu, sh, T.lane_hi);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int df = T.df0 + c;
        if (T.bc + df < 0) d[c] = d_lo;
        else if (df > T.hi_df) d[c] = d_hi;
      }
    }
    // weights, branch-free (nlm_filter_avx.c:214-221); dead lanes/blocks/candidates get 0
    const bool on = T.active && ((gate[tau] >> T.bl) & 1u);
    float w[4];
    float wany = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float wc = schraudolph_exp(-d[c] * T.inv);
      wc = (d[c] > T.thr) ? 0.f : wc;
      wc = (wc < kNlmMinWeight) ? 0.f : wc;
      wc = (on && c < T.nc) ? wc : 0.f;
      w[c] = wc;
      wany += wc;
    }
    const int ph = tau & (RING2 - 1);
    // the ring slot must have been handed back by the reducers (target tau - RING2)
    if (lane == 0) mbar_wait(&bar_empty[ph], ((unsigned)(tau / RING2) & 1u) ^ 1u);
    __syncwarp();
    if (wany != 0.f) {
      float pacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* crow = T.tile + (16 + tau + T.dt) * PITCH2 + T.col_tg + T.df0;   // frame t+dt
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (w[c] != 0.f) {
#pragma unroll
          for (int j = 0; j < 8; j++) pacc[j] = fmaf(w[c], crow[c + j], pacc[j]);
        }
      }
      float pw = 0.f;
#pragma unroll
      for (int c = 0; c < 4; c++) pw += w[c];
      float* p = part + ((size_t)(ph * B2 + T.bl) * SLOTS2 + T.slot_id) * 9;
#pragma unroll
      for (int j = 0; j < 8; j++) p[j] = pacc[j];
      p[8] = pw;
      atomicOr(&pmask[(ph * B2 + T.bl) * MW2 + (T.slot_id >> 5)], 1u << (T.slot_id & 31));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_full[ph]);
  }
}

__global__ void __launch_bounds__(NTH2, 1)
nlm2_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
            int targets, int last_row, float h, float* __restrict__ out, int pitchO) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  float* part = reinterpret_cast<float*>(smem_raw + SMEM2_TILE);
  unsigned* pmask = reinterpret_cast<unsigned*>(smem_raw + SMEM2_TILE + SMEM2_PART);
  unsigned* gate = reinterpret_cast<unsigned*>(smem_raw + SMEM2_TILE + SMEM2_PART + SMEM2_MASK);
  unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(
      smem_raw + SMEM2_TILE + SMEM2_PART + SMEM2_MASK + SMEM2_GATE);
  unsigned long long* bar_empty = bar_full + RING2;

  const int blk0 = blockIdx.x * B2;
  const int tau0 = blockIdx.y * TT2;
  const int nT = (targets - tau0 < TT2) ? targets - tau0 : TT2;
  const int row_t0 = first_target_row + tau0;        // S row of target 0 of this tile
  const int c_origin = blk0 * 8 - 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- stage rows t0-16 .. t0+nT+3 (clamped), edge-replicated columns ----
  for (int e = threadIdx.x; e < ROWS2 * PITCH2; e += NTH2) {
    const int r = e / PITCH2;
    const int c_rel = e - r * PITCH2;
    int c = c_origin + c_rel;
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    int gr = row_t0 - kNlmPast + r;
    gr = gr > last_row ? last_row : gr;
    tile[e] = S[(size_t)gr * pitchS + c];
  }
  if (threadIdx.x < RING2 * B2 * MW2) pmask[threadIdx.x] = 0u;
  if (threadIdx.x == 0) {
    for (int p = 0; p < RING2; p++) {
      mbar_init(&bar_full[p], NDT);        // one arrival per producer warp
      mbar_init(&bar_empty[p], NRED2);     // one arrival per reducer warp
    }
  }
  __syncthreads();
  // block energy gate per (target, paste block): sum of the 8 target bins >= 1e-6
  // (nlm_filter_avx.c:123-129)
  if (threadIdx.x < TT2) {
    unsigned bits = 0u;
    const float* tr = tile + (kNlmPast + threadIdx.x) * PITCH2 + 8;
    for (int bl = 0; bl < B2; bl++) {
      float ssum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; q++) ssum += tr[8 * bl + q];
      if (ssum >= 1e-6f) bits |= 1u << bl;
    }
    gate[threadIdx.x] = bits;
  }
  __syncthreads();

  if (warp < NDT) {
    // =========================== producers ===========================
    Nlm2Thread T;
    T.tile = tile;
    T.active = lane < B2 * G2;
    const int bl = T.active ? lane / G2 : 0;
    const int g = T.active ? lane - bl * G2 : 0;
    const int b = blk0 + bl;
    if (b >= nfull) T.active = false;
    T.edge = (blk0 == 0) || (blk0 + B2 >= nfull);
    T.bl = bl;
    // warp -> dt: the 7 dt whose patches wrap through the ring do twice the row
    // terms; they get the highest producer warp ids (issue priority) and are
    // spread 1/2/2/2 over the four SM sub-partitions (warp id mod 4)
    const int dti = c_dt_of_warp[warp];
    T.dt = dti - kNlmPast;
    T.slot_id = dti * G2 + g;
    T.df0 = -kNlmRangeFreq + 4 * g;
    T.nc = (g == G2 - 1) ? 1 : 4;
    T.col_tg = 8 + 8 * bl;
    T.col_seg = T.col_tg + T.df0;
    T.bc = 8 * b + 4;
    block_params(T.bc, K, h, &T.inv, &T.thr);
    T.rlo = 0;
    T.rhi = 8;
    T.wrap_shift = 0;
    if (T.dt > 1) { T.rhi = 9 - T.dt; T.wrap_shift = -NDT; }          // dt = 2,3,4
    if (T.dt < -12) { T.rlo = -12 - T.dt; T.wrap_shift = NDT; }       // dt = -13..-16
    {
      const int lane_base = bl * G2;
      int dlo = -T.bc;                                 // -4 for paste block 0, else below -8
      if (dlo < -kNlmRangeFreq) dlo = -kNlmRangeFreq;
      T.lane_lo = lane_base + (dlo + kNlmRangeFreq) / 4;
      T.c_lo = (dlo + kNlmRangeFreq) & 3;
      T.hi_df = K - 1 - T.bc;
      int dhi = T.hi_df > kNlmRangeFreq ? kNlmRangeFreq : T.hi_df;
      if (dhi < -kNlmRangeFreq) dhi = -kNlmRangeFreq;
      T.lane_hi = lane_base + (dhi + kNlmRangeFreq) / 4;
      T.c_hi = (dhi + kNlmRangeFreq) & 3;
    }
    float acc[8][4];
#pragma unroll
    for (int s = 0; s < 8; s++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[s][c] = 0.f;
    const int steps = nT + 7;
#define SB_NLM2_STEP(PHI)                                                                  \
  if (i0 + PHI < steps)                                                                    \
    nlm2_step<PHI>(T, i0 + PHI, nT, acc, part, pmask, gate, bar_full, bar_empty, lane);
    for (int i0 = 0; i0 < steps; i0 += 8) {
      SB_NLM2_STEP(0) SB_NLM2_STEP(1) SB_NLM2_STEP(2) SB_NLM2_STEP(3)
      SB_NLM2_STEP(4) SB_NLM2_STEP(5) SB_NLM2_STEP(6) SB_NLM2_STEP(7)
    }
#undef SB_NLM2_STEP
  } else {
    // =========================== reducers ============================
    const int rt = (warp - NDT) * 32 + lane;       // 0..63
    const int rbl = rt >> 3, rj = rt & 7;
    const bool valid = rbl < B2 && blk0 + rbl < nfull;
    for (int tau = 0; tau < nT; tau++) {
      const int ph = tau & (RING2 - 1);
      if (lane == 0) mbar_wait(&bar_full[ph], (unsigned)(tau / RING2) & 1u);
      __syncwarp();
      if (valid) {
        unsigned* mw = &pmask[(ph * B2 + rbl) * MW2];
        const float* pb = part + (size_t)(ph * B2 + rbl) * SLOTS2 * 9;
        float a = 0.f, ws = 0.f;
        for (int w = 0; w < MW2; w++) {
          unsigned m = mw[w];
          while (m) {                               // ascending slot = (dt, df) order
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const float* p = pb + (size_t)(w * 32 + bit) * 9;
            a += p[rj];
            ws += p[8];
          }
        }
        const float tv = tile[(kNlmPast + tau) * PITCH2 + 8 + 8 * rbl + rj];
        out[(size_t)(tau0 + tau) * pitchO + 8 * (blk0 + rbl) + rj] =
            (ws > kNlmMinWeight) ? a / ws : tv;     // nlm_filter_avx.c:237-243
      }
      __syncwarp();
      if (valid && rj < MW2) pmask[(ph * B2 + rbl) * MW2 + rj] = 0u;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[ph]);
    }
  }
}

}  // namespace

bool launch_nlm(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                float* Shat, cudaStream_t st) {
  if (targets <= 0) return true;
  const int nfull = g.K / 8;
  // SB_NLM_KERNEL=1: direct-form kernel (reference accumulation order); 2: row-SSD sharing
  // with producer/reducer warps; default 3: sb_nlm3.cu
  static const int variant = [] {
    const char* e = getenv("SB_NLM_KERNEL");
    return e ? atoi(e) : 3;
  }();
  KTimer* kt = g_profile ? new KTimer(kcNlm, st) : nullptr;
  if (nfull > 0 && variant == 3) {
    if (!launch_nlm3(g, S, first_target_row, targets, h, Shat, st)) {
      delete kt;
      return false;
    }
  } else if (nfull > 0 && variant == 2) {
    SB_CUDA(cudaFuncSetAttribute((const void*)nlm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)SMEM2));   // per device, hence on every call
    dim3 grid((nfull + B2 - 1) / B2, (targets + TT2 - 1) / TT2);
    // rows of S that exist: the last target's +4 look-ahead row
    const int last_row = first_target_row + targets - 1 + kNlmFuture;
    nlm2_kernel<<<grid, NTH2, SMEM2, st>>>(S, g.Kp, g.K, nfull, first_target_row, targets,
                                           last_row, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  } else if (nfull > 0) {
    dim3 grid((nfull + BT - 1) / BT, targets);
    nlm_kernel<<<grid, NTHREADS, 0, st>>>(S, g.Kp, g.K, nfull, first_target_row, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  }
  delete kt;
  if (g.K % 8 != 0) {
    KTimer kt2(kcNlmTail, st);
    nlm_tail_kernel<<<(targets + TAILT - 1) / TAILT, TAILT * 32, 0, st>>>(S, g.Kp, g.K, nfull, first_target_row,
                                                                          targets, h, Shat, g.Kp);
    SB_LAUNCH_CHECK();
  }
  return true;
}

}  // namespace sb
