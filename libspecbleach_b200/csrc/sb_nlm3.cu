// NLM v3: row-SSD sharing with free-running warps and a two-pass paste.
//
// Same mathematics as sb_nlm.cu (nlm_filter_avx.c:87-248): for target frame t,
// paste block b, candidate (dt, df)
//   d = sum_{r=0..7} R[t+r-4][lag(dt,r)][df],
//   R[a][lag][df] = sum_{q<8} (S[a][bs+q] - S[a+lag][bs+df+q])^2,
//   lag(dt,r) = dt        when dt+r-4 lies in [-16,4]
//             = dt -/+ 21 otherwise (21-slot ring alias, quirk Q2).
// A row term R[a][lag] serves the 8 consecutive targets whose patches contain
// row a, so it is computed once while a warp walks down the rows of a tile of 64
// consecutive targets (~19 FP instructions per candidate instead of 128).
//
// Pass 1 (distances): CTA = 64 targets x 16 paste blocks, 12 warps.  A lane owns
//   (paste block, df half): df in [-8,-1] or [0,8] -> 9 candidates from one 16-float
//   partner segment (4 aligned LDS.128) against the 8 target bins (2 LDS.128):
//   144 FP per 6 LDS.128.  A warp owns whole dt values ("units") and walks them
//   one after the other with NO inter-warp synchronisation:
//     * the 14 dt that never wrap add their row terms through a delay tree
//         A_i = R_{i-1} + R_i,  B_i = A_{i-2} + A_i,  d = B_{i-4} + B_i
//       i.e. ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)): 3 FADD per candidate and row;
//     * the 7 dt that wrap (-16..-13, 2..4) need two lags per row and keep the 8
//       targets in flight in registers (r0+r1+...+r7 in order).
//   Units come from a per-CTA queue (wrapping dt first, the last plain dt split in two target
//   halves), so a warp that finishes early takes the next one.
//   Pass 1 outputs, per (target, block), a 357-bit field of the candidates that pass the
//   hard threshold d <= 4 h_b^2 (a few percent, nlm_filter_avx.c:208-212) and, for each of
//   them, its weight at a fixed scratch address W[target][block][candidate] -- every
//   (target, block, candidate) has exactly one owner lane, so the scratch content does
//   not depend on scheduling.  Targets are processed in chunks of ~2k (whole waves of CTAs on
//   the device's SMs) so that most of the touched scratch (tens of MB) stays in L2.
// Pass 2 (nlm3_paste_kernel): 8 lanes per (target, block) walk the bit field and paste the
//   flagged candidates in the reference's (dt, df) order.  Deterministic and independent
//   of how the targets are tiled, so streaming stays bit-exact.
//
// Compiled with -ftz=true (the reference runs NLM under FTZ|DAZ).
//
// Round 2 tried and measured, then removed (commit 3e3f1e0 holds the code of
// the first three): a fused persistent form (one launch per call, paste as the tile's epilogue:
// -6 % NLM time, the two-lane step slower), a paste with lanes over candidates, two packed FFMA2 /
// FADD2 forms of the row walk (25 % fewer instructions, same time: the FP32 pipe is what this
// kernel saturates), and three 4-warp CTAs per SM with the bit fields set in global memory
// (device-resident +3 %, end to end -1 %).
#include "sb_common.h"

#include <map>
#include <utility>
#include <mutex>

namespace sb {
namespace {

constexpr int NDT = kNlmRing;               // 21
constexpr int NB3 = 16;                     // paste blocks per CTA
constexpr int TT3 = 64;                     // targets per CTA
constexpr int ROWS3 = TT3 + 20;             // staged rows t0-16 .. t0+TT3+3
constexpr int GRP3 = NB3 + 2;               // 8-bin column groups per row: bins [8*blk0-8, 8*blk0+8*NB3+8)
constexpr int SLOT3 = 12;                   // floats per column group: consecutive paste blocks are
                                            // 48 B apart, so a quarter-warp's LDS.128 hit 32 banks
constexpr int PITCH3 = GRP3 * SLOT3 + 4;    // 220 floats (== 28 mod 32: patch rows on distinct banks)
constexpr int WPITCH3 = 360;                // scratch floats per (target, block): 357 candidates
constexpr int GATED3 = 357;                 // bit 357 of the field: block skipped by the energy gate
constexpr int CHUNK3 = 2048;                // nominal targets per launch pair (scratch stays in L2)
constexpr int NW3 = 12;                     // warps
constexpr int NTH3 = NW3 * 32;              // 384 threads
constexpr int FW3 = 12;                     // flag words per (target, block): 21*17 = 357 bits
constexpr int NC3 = 9;                      // candidates per lane
constexpr size_t SMEM3_TILE = sizeof(float) * (size_t)ROWS3 * PITCH3;
constexpr size_t SMEM3_FLAGS = sizeof(unsigned) * (size_t)TT3 * NB3 * FW3;
constexpr size_t SMEM3_GATE = sizeof(unsigned) * TT3;
constexpr size_t SMEM3_PRM = sizeof(float) * 2 * NB3 + 16;               // + unit queue cursor
constexpr size_t SMEM3 = SMEM3_TILE + SMEM3_FLAGS + SMEM3_GATE + SMEM3_PRM;

// tile-relative bin column -> float offset inside a staged row
__device__ __forceinline__ int tcol3(int c_rel) { return SLOT3 * (c_rel >> 3) + (c_rel & 7); }

// Units in the order warps take them from the CTA's queue: {dt + 16, half}.  half 0 = all
// targets of the tile, 1 / 2 = first / second half.  The 7 wrapping dt (~2.5 x the work of a
// plain one) go first, the last plain dt are split in two so that the tail of the tile is
// made of short units.
constexpr int NUNIT3 = 27;
__constant__ signed char c_units3[NUNIT3][2] = {
    {0, 0},  {1, 0},  {2, 0},  {3, 0},  {18, 0}, {19, 0}, {20, 0}, {4, 0},  {5, 0},
    {6, 0},  {7, 0},  {8, 0},  {9, 0},  {10, 0}, {11, 0}, {12, 1}, {12, 2}, {13, 1},
    {13, 2}, {14, 1}, {14, 2}, {15, 1}, {15, 2}, {16, 1}, {16, 2}, {17, 1}, {17, 2}};

__device__ __forceinline__ float schraudolph_exp3(float x) {   // simd_utils.h:670-684
  if (x < -20.0f) return 0.0f;
  return __uint_as_float(__float2uint_rz(fmaf(12102203.0f, x, 1064866805.0f)));
}

__device__ __forceinline__ void block_params3(int bc, int K, float h, float* inv, float* thr) {
  float h2 = h * h;                                             // nlm_filter_avx.c:131-147
  if (K > 1) {
    const float nf = (float)bc / (float)(K - 1);
    const float fs = 1.0f + kNlmFreqScale * nf * nf;
    const float hv = h * fs;
    h2 = hv * hv;
  }
  *inv = 1.0f / h2;
  *thr = kNlmThrMul * h2;
}

struct Lane3 {
  const float* tile;
  unsigned* flags;        // flags of this lane's paste block: + tau * NB3 * FW3
  float* wout;            // weight scratch of this lane's paste block: + tau * wstride
  int wstride;            // floats between consecutive targets in the weight scratch
  unsigned long long live;   // bit tau: this lane's paste block passes the energy gate at target tau
  int col_tg;             // float offset of bin bs inside a staged row
  int col_seg;            // float offset of the first partner bin of this lane
  int g;                  // df half: 0 -> df = -8 + c (c < 8), 1 -> df = c (c < 9)
  int bl;
  int hi_df;              // df whose centre is bin K-1 (>= 9: none clamps)
  bool lo_clamp;          // paste block 0, lower half: df < -4 clamp to centre 0
  bool active;
  float thr, inv;
};

// row terms of the 9 candidates of this lane for target-side row `ra` against row `rp`
__device__ __forceinline__ void rowterms3(const Lane3& L, int ra, int rp, float (&R)[NC3]) {
  rp = rp < 0 ? 0 : (rp > ROWS3 - 1 ? ROWS3 - 1 : rp);
  const float4* pt = reinterpret_cast<const float4*>(L.tile + ra * PITCH3 + L.col_tg);
  const float4* ps = reinterpret_cast<const float4*>(L.tile + rp * PITCH3 + L.col_seg);
  const float4 t0 = pt[0], t1 = pt[1];
  const float4 s0 = ps[0], s1 = ps[1], s2 = ps[3], s3 = ps[4];   // two 12-float column groups
  const float tg[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
  const float seg[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w,
                         s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
  for (int c = 0; c < NC3; c++) {
    float r = 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const float d = tg[q] - seg[c + q];
      r = fmaf(d, d, r);
    }
    R[c] = r;
  }
}

// Rare path of pass 1: at least one of the lane's candidates passed.  Every one sets its bit in
// the (target, block) field and leaves its weight at W[target][block][candidate].  cmask
// drops c = 8 of the lower df half, which duplicates df = 0.
__device__ __forceinline__ void emit_slow3(unsigned* fl, float* w, int p0, unsigned cmask, float thr,
                                        float inv, float d0, float d1, float d2, float d3,
                                        float d4, float d5, float d6, float d7, float d8) {
  const float d[NC3] = {d0, d1, d2, d3, d4, d5, d6, d7, d8};
  unsigned bits = 0u;
#pragma unroll
  for (int c = 0; c < NC3; c++) {
    if (!(d[c] > thr) && ((cmask >> c) & 1u)) {
      bits |= 1u << c;
      w[p0 + c] = schraudolph_exp3(-d[c] * inv);   // nlm_filter_avx.c:214-221
    }
  }
  if (bits) {
    unsigned* f = fl + (p0 >> 5);
    const int sh = p0 & 31;
    atomicOr(f, bits << sh);
    if (sh > 32 - NC3 && (bits >> (32 - sh)) != 0u) atomicOr(f + 1, bits >> (32 - sh));
  }
}

// candidates that pass the hard threshold -> bit field + weights of (target tau, this paste block)
//
// Hot path: one min over the lane's 9 distances and one compare.  Everything that rewrites a
// distance lives in the rare branch, so the row walk keeps its registers untouched:
//  * candidates whose centre clamps to bin 0 / K-1 take the distance of the candidate that sits
//    on the edge (clamp_index is applied to the centre first).  Their fixed-up distances are
//    copies of other distances of the same lane, so min(original) <= min(fixed-up) and the
//    un-fixed min is a conservative test;
//  * the self candidate (dt 0, df 0: d == 0, always accepted) is pasted unconditionally by
//    pass 2 and never flagged here, so that quiet tiles never take the rare branch.
__device__ __forceinline__ void emit_flags3(const Lane3& L, int tau, int dti, const float (&d)[NC3],
                                            bool edge, bool valid) {
  const bool self = (dti == kNlmPast) && (L.g == 1);
  const float d0 = self ? 3.0e38f : d[0];
  const float dmin = fminf(fminf(fminf(d0, d[1]), fminf(d[2], d[3])),
                           fminf(fminf(d[4], d[5]), fminf(fminf(d[6], d[7]), d[8])));
  if (valid && !(dmin > L.thr) && ((L.live >> (tau & 63)) & 1ull)) {
    float e[NC3];
#pragma unroll
    for (int c = 0; c < NC3; c++) e[c] = d[c];
    if (edge) {   // block-uniform: first / last block tile
      if (L.lo_clamp) {
#pragma unroll
        for (int c = 0; c < 4; c++) e[c] = d[4];
      }
      if (L.g == 1 && L.hi_df < 8) {
        float dh = d[3];
#pragma unroll
        for (int c = 4; c < 8; c++) dh = (L.hi_df == c) ? d[c] : dh;
#pragma unroll
        for (int c = 4; c < NC3; c++) e[c] = (c > L.hi_df) ? dh : d[c];
      }
    }
    const unsigned cmask = (L.g == 0 ? 0xffu : 0x1ffu) & (self ? ~1u : ~0u);
    emit_slow3(L.flags + (size_t)tau * (NB3 * FW3), L.wout + (size_t)tau * L.wstride,
               dti * 17 + 8 * L.g, cmask, L.thr, L.inv, e[0], e[1], e[2], e[3], e[4], e[5], e[6],
               e[7], e[8]);
  }
}

// ---- a dt that never wraps: delay-tree accumulation ----
// targets [ta, tb) of the tile: rows ta .. tb+6 are walked, the first 7 only fill the tree
__device__ __forceinline__ void walk_plain3(const Lane3& L, int dti, int ta, int tb, bool edge) {
  const int dt = dti - kNlmPast;
  float R1[NC3], A[2][NC3], B[4][NC3];
#pragma unroll
  for (int c = 0; c < NC3; c++) {
    R1[c] = 0.f;
    A[0][c] = A[1][c] = 0.f;
    B[0][c] = B[1][c] = B[2][c] = B[3][c] = 0.f;
  }
  // rows ta .. tb+6, padded to a multiple of 4 steps (the padding rows are clamped into the
  // tile and their distances never emitted), so that the unrolled body has no step guard
  const int steps = tb + 7;
  for (int i0 = ta; i0 < steps; i0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u;
      int ra = 12 + i;                             // tile row of frame t0 - 4 + i
      ra = ra > ROWS3 - 1 ? ROWS3 - 1 : ra;
      float R0[NC3];
      rowterms3(L, ra, ra + dt, R0);
      float d[NC3];
#pragma unroll
      for (int c = 0; c < NC3; c++) {
        const float a0 = R1[c] + R0[c];
        const float b0 = A[u & 1][c] + a0;
        d[c] = B[u & 3][c] + b0;
        R1[c] = R0[c];
        A[u & 1][c] = a0;
        B[u & 3][c] = b0;
      }
      const int tau = i - 7;
      emit_flags3(L, tau, dti, d, edge, tau >= ta && tau < tb);
    }
  }
}

// ---- a dt whose patch rows wrap through the ring: 8 targets in flight ----
// age[r] = partial distance of the target whose patch row r is the current row.  Every step the
// partial sums move one age up THROUGH the add (age[r] = age[r-1] + term, evaluated from r = 7
// down), so the loop needs neither register rotation nor unrolling: r0+r1+...+r7 in order.
// The alias choice per patch row is a run-time select (72 FSEL per step, 7.5 % of the kernel's
// instructions).  Measured alternatives: dt as a template parameter folds the selects but the
// seven bodies (52 KB next to the 25 KB plain walker) thrash the 32 KB instruction cache, 3x
// slower; a warp-uniform branch per row is 4 % slower than the selects.
__device__ __forceinline__ void walk_wrap3(const Lane3& L, int dti, int ta, int tb, bool edge) {
  const int dt = dti - kNlmPast;
  int rlo = 0, rhi = 8, shift = 0;                   // rows r < rlo or r >= rhi use the alias
  if (dt > 1) { rhi = 9 - dt; shift = -NDT; }        // dt = 2,3,4
  else { rlo = -12 - dt; shift = NDT; }              // dt = -13..-16
  float age[8][NC3];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < NC3; c++) age[r][c] = 0.f;
  const int steps = tb + 7;
#pragma unroll 1
  for (int i = ta; i < steps; i++) {
    const int ra = 12 + i;
    float Rm[NC3], Rw[NC3];
    rowterms3(L, ra, ra + dt, Rm);
    rowterms3(L, ra, ra + dt + shift, Rw);
#pragma unroll
    for (int r = 7; r >= 1; r--) {
      const bool w = (r < rlo) || (r >= rhi);
#pragma unroll
      for (int c = 0; c < NC3; c++) age[r][c] = age[r - 1][c] + (w ? Rw[c] : Rm[c]);
    }
#pragma unroll
    for (int c = 0; c < NC3; c++) age[0][c] = (0 < rlo) ? Rw[c] : Rm[c];
    const int tau = i - 7;
    if (tau >= ta) {
      float d[NC3];
#pragma unroll
      for (int c = 0; c < NC3; c++) d[c] = age[7][c];
      emit_flags3(L, tau, dti, d, edge, true);
    }
  }
}

// S rows are addressed relative to `first_target_row` (+ chunk-local target index); Wg / Fg are
// the chunk's scratch: Wg[(target * nfull + block) * WPITCH3 + candidate], Fg[(target * nfull +
// block) * FW3 + word].
__global__ void __launch_bounds__(NTH3, 1)
nlm3_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
            int targets, int last_row, float h, float* __restrict__ Wg, unsigned* __restrict__ Fg) {
  constexpr int NTH = NTH3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  float* tile = reinterpret_cast<float*>(sp);            sp += SMEM3_TILE;
  unsigned* flags = reinterpret_cast<unsigned*>(sp);     sp += SMEM3_FLAGS;
  unsigned* gate = reinterpret_cast<unsigned*>(sp);      sp += SMEM3_GATE;
  float* prm = reinterpret_cast<float*>(sp);
  unsigned* next_unit = reinterpret_cast<unsigned*>(prm + 2 * NB3);

  // the highest block tiles hold the most candidates (h_b grows with frequency): schedule
  // them first so that they do not form the tail of the launch
  const int blk0 = ((int)gridDim.y - 1 - (int)blockIdx.y) * NB3;
  const int tau0 = blockIdx.x * TT3;
  const int nT = (targets - tau0 < TT3) ? targets - tau0 : TT3;
  const int row_t0 = first_target_row + tau0;          // S row of target 0 of this tile
  const int c_origin = blk0 * 8 - 8;
  const int lane = threadIdx.x & 31;

  // ---- stage rows t0-16 .. t0+nT+3 (clamped to the rows that exist), edge-replicated bins ----
  for (int e = threadIdx.x; e < ROWS3 * GRP3 * 8; e += NTH) {
    const int r = e / (GRP3 * 8);
    const int c_rel = e - r * (GRP3 * 8);
    int c = c_origin + c_rel;
    c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
    int gr = row_t0 - kNlmPast + r;
    gr = gr > last_row ? last_row : gr;
    tile[r * PITCH3 + tcol3(c_rel)] = S[(size_t)gr * pitchS + c];
  }
  for (int e = threadIdx.x; e < TT3 * NB3 * FW3; e += NTH) flags[e] = 0u;
  if (threadIdx.x == 0) *next_unit = 0u;
  if (threadIdx.x < NB3) {                              // 1/h_b^2 and 4 h_b^2 per paste block
    float inv, thr;
    block_params3(8 * (blk0 + threadIdx.x) + 4, K, h, &inv, &thr);
    prm[threadIdx.x] = inv;
    prm[NB3 + threadIdx.x] = thr;
  }
  __syncthreads();
  // block energy gate per (target, paste block): nlm_filter_avx.c:123-129
  if (threadIdx.x < TT3) {
    unsigned bits = 0u;
    const float* tr = tile + (kNlmPast + threadIdx.x) * PITCH3;
    for (int bl = 0; bl < NB3; bl++) {
      float ssum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; q++) ssum += tr[SLOT3 * (bl + 1) + q];
      if (ssum >= 1e-6f) bits |= 1u << bl;
    }
    gate[threadIdx.x] = bits;
  }
  __syncthreads();

  {
    Lane3 L;
    L.tile = tile;
    L.bl = lane >> 1;
    L.g = lane & 1;
    const int b = blk0 + L.bl;
    L.active = b < nfull;
    L.live = 0ull;
    if (L.active)
      for (int t = 0; t < TT3; t++) L.live |= (unsigned long long)((gate[t] >> L.bl) & 1u) << t;
    L.flags = flags + L.bl * FW3;
    L.wstride = nfull * WPITCH3;
    L.wout = Wg + ((size_t)tau0 * nfull + (L.active ? b : 0)) * WPITCH3;
    L.col_tg = SLOT3 * (L.bl + 1);
    L.col_seg = SLOT3 * (L.bl + L.g);
    const int bc = 8 * b + 4;
    L.inv = prm[L.bl];
    L.thr = prm[NB3 + L.bl];
    L.hi_df = K - 1 - bc;
    L.lo_clamp = (b == 0) && (L.g == 0);
    const bool edge = (blk0 == 0) || (blk0 + NB3 >= nfull);
#pragma unroll 1
    for (;;) {
      int u = 0;
      if (lane == 0) u = (int)atomicAdd(next_unit, 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= NUNIT3) break;
      const int dti = c_units3[u][0], half = c_units3[u][1];
      const int mid = (nT + 1) >> 1;
      const int ta = half == 2 ? mid : 0;
      const int tb = half == 1 ? mid : nT;
      if (ta >= tb) continue;
      if (dti < 4 || dti > 17) walk_wrap3(L, dti, ta, tb, edge);
      else walk_plain3(L, dti, ta, tb, edge);
    }
  }
  __syncthreads();

  // ---- bit fields (and the gate) of the tile -> scratch ----
  const int nbv = (nfull - blk0 < NB3) ? nfull - blk0 : NB3;      // valid paste blocks of this tile
  for (int e = threadIdx.x; e < nT * nbv * FW3; e += NTH) {
    const int tau = e / (nbv * FW3);
    const int rem = e - tau * (nbv * FW3);
    const int bl = rem / FW3;
    const int w = rem - bl * FW3;
    unsigned v = flags[(tau * NB3 + bl) * FW3 + w];
    if (w == (GATED3 >> 5) && !((gate[tau] >> bl) & 1u)) v |= 1u << (GATED3 & 31);
    Fg[((size_t)(tau0 + tau) * nfull + blk0 + bl) * FW3 + w] = v;
  }
}

// Pass 2: 8 lanes per (target, block); lane j owns bin 8*b + j.
__global__ void __launch_bounds__(128)
nlm3_paste_kernel(const float* __restrict__ S, int pitchS, int K, int nfull, int first_target_row,
                  int targets, float h, const float* __restrict__ Wg,
                  const unsigned* __restrict__ Fg, float* __restrict__ out, int pitchO) {
  const int pair = blockIdx.x * 16 + (threadIdx.x >> 3);
  const int j = threadIdx.x & 7;
  if (pair >= targets * nfull) return;
  const int tau = pair / nfull;
  const int b = pair - tau * nfull;
  const int bs = 8 * b;
  const float* trow = S + (size_t)(first_target_row + tau) * pitchS;      // frame t
  const float tv = trow[bs + j];
  const uint4* fp = reinterpret_cast<const uint4*>(Fg + (size_t)pair * FW3);
  const uint4 f0 = fp[0], f1 = fp[1], f2 = fp[2];
  const unsigned fw[FW3] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
  float res = tv;
  if (!((fw[GATED3 >> 5] >> (GATED3 & 31)) & 1u)) {
    const float w_self = schraudolph_exp3(-0.0f);     // self candidate: d == 0 for every h_b
    const float* wp = Wg + (size_t)pair * WPITCH3;
    float acc = 0.f, ws = 0.f;
    // dt-major walk: the 17 df bits of one dt are a bit field at a compile-time position, the
    // pasted row pointer is hoisted, ascending bit = ascending df (the reference's order)
#pragma unroll
    for (int dti = 0; dti < NDT; dti++) {
      constexpr int kDfBits = 2 * kNlmRangeFreq + 1;
      const int p0 = dti * kDfBits, w0 = p0 >> 5, sh = p0 & 31;
      unsigned m = fw[w0] >> sh;
      if (sh + kDfBits > 32) m |= fw[w0 + 1 < FW3 ? w0 + 1 : w0] << (32 - sh);
      m &= (1u << kDfBits) - 1u;
      if (dti == kNlmPast) m |= 1u << kNlmRangeFreq;   // the self candidate is always accepted
      const float* prow = trow + (ptrdiff_t)(dti - kNlmPast) * pitchS;   // frame t + dt
      const float* wrow = wp + p0;
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        // both loads are issued before either is used: one L2 round trip per candidate
        int c = bs + j + bit - kNlmRangeFreq;          // bin clamp(bs + j + df)
        c = c < 0 ? 0 : (c > K - 1 ? K - 1 : c);
        const float v = __ldg(prow + c);
        float wgt = __ldg(wrow + bit);
        if (dti == kNlmPast && bit == kNlmRangeFreq) wgt = w_self;
        if (!(wgt < kNlmMinWeight)) {                  // nlm_filter_avx.c:219-221
          acc = fmaf(wgt, v, acc);
          ws += wgt;
        }
      }
    }
    res = (ws > kNlmMinWeight) ? acc / ws : tv;        // nlm_filter_avx.c:237-243
  }
  out[(size_t)tau * pitchO + bs + j] = res;
}


}  // namespace

// Targets per chunk: close to CHUNK3, and such that the grid of the distance kernel (target tiles
// x block tiles, one CTA per SM) fills whole waves of the device's SMs.
int nlm3_chunk_targets(int K) {
  const int nfull = K / 8;
  const int btiles = (nfull + NB3 - 1) / NB3;
  if (btiles <= 0) return CHUNK3;
  static int sms_of[64] = {0};          // per device (handles on several GPUs of one process)
  int dev = 0;
  cudaGetDevice(&dev);
  int& sms = sms_of[dev & 63];
  if (sms == 0 &&
      (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0))
    sms = 148;
  int waves = (CHUNK3 / TT3 * btiles + sms / 2) / sms;      // nearest whole number of waves
  if (waves < 1) waves = 1;
  int ttiles = waves * sms / btiles;
  if (ttiles < 1) ttiles = 1;
  return ttiles * TT3;
}

// Scratch per stream (= per engine lane): weights and bit fields of one chunk of targets.
struct Scratch3 {
  float* W = nullptr;
  unsigned* F = nullptr;
  size_t pairs = 0;
};
static std::mutex g_scratch_mu;
// keyed by (device, stream): the legacy default stream is the same handle on every device
static std::map<std::pair<int, cudaStream_t>, Scratch3> g_scratch;

bool launch_nlm3(const GeoK& g, const float* S, int first_target_row, int targets, float h,
                 float* Shat, cudaStream_t st) {
  const int nfull = g.K / 8;
  if (nfull <= 0 || targets <= 0) return true;
  // per device and context, hence on every call rather than once per process
  SB_CUDA(cudaFuncSetAttribute((const void*)nlm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SMEM3));
  const int chunk = nlm3_chunk_targets(g.K);
  Scratch3 sc;
  {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    Scratch3& ref = g_scratch[std::make_pair(dev, st)];
    // sized for the largest chunk seen on this stream (small sub-calls keep it small)
    const size_t need = (size_t)(targets < chunk ? targets : chunk) * (size_t)nfull;
    if (ref.pairs < need) {
      SB_CUDA(cudaStreamSynchronize(st));
      if (ref.W) dev_free(ref.W);
      if (ref.F) dev_free(ref.F);
      ref = Scratch3();
      SB_CUDA(dev_malloc(&ref.W, need * WPITCH3 * sizeof(float)));
      SB_CUDA(dev_malloc(&ref.F, need * FW3 * sizeof(unsigned)));
      ref.pairs = need;
    }
    sc = ref;
  }
  // rows of S that exist: the last target's +4 look-ahead row
  const int last_row = first_target_row + targets - 1 + kNlmFuture;
  for (int t0 = 0; t0 < targets; t0 += chunk) {
    const int cnt = targets - t0 < chunk ? targets - t0 : chunk;
    dim3 grid((cnt + TT3 - 1) / TT3, (nfull + NB3 - 1) / NB3);
    nlm3_kernel<<<grid, NTH3, SMEM3, st>>>(S, g.Kp, g.K, nfull, first_target_row + t0, cnt, last_row, h,
                                           sc.W, sc.F);
    SB_LAUNCH_CHECK();
    const int pairs = cnt * nfull;
    // (paste on a side stream under the next chunk's distance kernel, two scratch sets: the NLM
    // launch alone 3.5 % faster, the two-lane step unchanged -- the other lane already fills
    // the ramps -- so the scratch stays single)
    nlm3_paste_kernel<<<(pairs + 15) / 16, 128, 0, st>>>(S, g.Kp, g.K, nfull, first_target_row + t0, cnt, h,
                                                         sc.W, sc.F, Shat + (size_t)t0 * g.Kp, g.Kp);
    SB_LAUNCH_CHECK();
  }
  return true;
}

}  // namespace sb
